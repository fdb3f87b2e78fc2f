/* mlbsim_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, no CUDA, no product sources) of the reference's Monte Carlo hot path:
 *   core/game_simulator.py:14-156        simulate_game / should_replace_pitcher
 *   core/half_inning_simulator.py:12-258 half-inning state machine, base-running handlers
 *   core/pa_simulator.py:23-143          beta_noise, resolve_base_outcome, check_home_run, resolve_contact
 *   core/bip_resolution.py:7-62          resolve_bip (probabilities arrive precomputed in mlbsim_params)
 *   core/fatigue_modeling.py:7-54        TTO / pitch-count multipliers
 *   assets/bullpen_utils.py:108-155      simulate_reliever_chain (pick index)
 *
 * Parity pin: `oracle_replay` is checked bit-exactly against tapes recorded from the UNMODIFIED
 * reference (tests/golden/replay_*.npz, produced by oracle/make_golden.py).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 *
 * Three entry points share one game engine:
 *   oracle_replay   draws come from a value-level tape (reference grammar, float64 compares)
 *   oracle_grammar  same engine, draws from a CPU RNG incl. Beta(30p,30(1-p)) noise — the CPU
 *                   port timed as cpu_baseline; can also emit the tape it consumed
 *   oracle_native   integer twin of the marginalised law the GPU kernel implements (one
 *                   Philox4x32-10 block per PA, 31-bit thresholds) for bit-exact histogram checks
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -pthread -shared -fPIC).
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "../include/mlbsim.h"


/* ------------------------------------------------------------------------------------------ */
/* pthread range splitter (no OpenMP dependency)                                              */
/* ------------------------------------------------------------------------------------------ */
static int g_threads = 0;
void oracle_set_threads(int n) { g_threads = n; }
int oracle_get_threads(void) {
  if (g_threads > 0) return g_threads;
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}
typedef void (*range_fn)(int64_t lo, int64_t hi, void* arg);
typedef struct { int64_t lo, hi; range_fn fn; void* arg; } range_job;
static void* range_tramp(void* p) { range_job* j = (range_job*)p; j->fn(j->lo, j->hi, j->arg); return NULL; }
static void parallel_ranges(int64_t lo, int64_t hi, range_fn fn, void* arg) {
  int nt = oracle_get_threads();
  int64_t n = hi - lo;
  if (n <= 0) return;
  if (nt > n) nt = (int)n;
  if (nt <= 1) { fn(lo, hi, arg); return; }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nt);
  range_job* jobs = (range_job*)malloc(sizeof(range_job) * nt);
  for (int t = 0; t < nt; ++t) {
    jobs[t].lo = lo + n * t / nt; jobs[t].hi = lo + n * (t + 1) / nt; jobs[t].fn = fn; jobs[t].arg = arg;
    if (pthread_create(&th[t], NULL, range_tramp, &jobs[t]) != 0) { th[t] = 0; fn(jobs[t].lo, jobs[t].hi, arg); }
  }
  for (int t = 0; t < nt; ++t) if (th[t]) pthread_join(th[t], NULL);
  free(th); free(jobs);
}
static pthread_mutex_t g_merge_lock = PTHREAD_MUTEX_INITIALIZER;

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al., SC'11), restated from the published algorithm.               */
/* ------------------------------------------------------------------------------------------ */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
    uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += PHILOX_W0; k1 += PHILOX_W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Philox2x32-10 (same paper): the generator of the native law */
#define PHILOX2_M 0xD256D193u
void oracle_philox2x32_10(const uint32_t ctr[2], uint32_t key, uint32_t out[2]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], k = key;
  for (int r = 0; r < 10; ++r) {
    uint64_t p = (uint64_t)PHILOX2_M * c0;
    c0 = (uint32_t)(p >> 32) ^ k ^ c1;
    c1 = (uint32_t)p;
    k += PHILOX_W0;
  }
  out[0] = c0; out[1] = c1;
}
/* the generator of the native law: Philox2x32 with NATIVE_PHILOX_ROUNDS rounds — 10 unless an experiment build of the
 * CUDA library says otherwise (csrc/philox.cuh) */
#ifndef NATIVE_PHILOX_ROUNDS
#define NATIVE_PHILOX_ROUNDS 10
#endif
static void philox2x32_native(const uint32_t ctr[2], uint32_t key, uint32_t out[2]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], k = key;
  for (int r = 0; r < NATIVE_PHILOX_ROUNDS; ++r) {
    uint64_t p = (uint64_t)PHILOX2_M * c0;
    c0 = (uint32_t)(p >> 32) ^ k ^ c1;
    c1 = (uint32_t)p;
    k += PHILOX_W0;
  }
  out[0] = c0; out[1] = c1;
}
/* 64-bit seed -> 32-bit Philox2x32 key: word 0 of Philox4x32-10 over a fixed domain-separation counter */
uint32_t oracle_native_key(uint64_t seed) {
  const uint32_t ctr[4] = {0x6d6c6273u, 0x696d3278u, 0u, 0u};
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t out[4];
  oracle_philox4x32_10(ctr, key, out);
  return out[0];
}
/* counter word 1: draw number [12:0] | kind [13] | matchup [25:14] | sim bits 37..32 [31:26] */
#define RNG_KIND_HALF_END (1u << 13)
#define RNG_MATCHUP_SHIFT 14
#define RNG_SIMHI_SHIFT 26

/* ------------------------------------------------------------------------------------------ */
/* draw sources for the float64 engine                                                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  /* tape read */
  const double* tape;
  int64_t len, pos;
  int underrun;
  /* rng (xoshiro256**) */
  int use_rng;
  uint64_t s[4];
  /* optional tape write */
  double* rec;
  int64_t rec_cap, rec_pos;
  int rec_overflow;
} source;

static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t splitmix64(uint64_t* x) {
  uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static uint64_t xoshiro_next(uint64_t s[4]) {
  uint64_t r = rotl64(s[1] * 5, 7) * 9, t = s[1] << 17;
  s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl64(s[3], 45);
  return r;
}
static inline double rng_u01(source* src) { return (double)(xoshiro_next(src->s) >> 11) * (1.0 / 9007199254740992.0); }

static inline void record(source* src, double v) {
  if (src->rec) {
    if (src->rec_pos < src->rec_cap) src->rec[src->rec_pos] = v; else src->rec_overflow = 1;
  }
  src->rec_pos++; /* counts draws even when nothing is stored */
}
static inline double tape_next(source* src) {
  if (src->pos >= src->len) { src->underrun = 1; src->pos++; return 0.5; }
  return src->tape[src->pos++];
}

static double rng_normal(source* src) { /* Marsaglia polar */
  for (;;) {
    double a = 2.0 * rng_u01(src) - 1.0, b = 2.0 * rng_u01(src) - 1.0, s = a * a + b * b;
    if (s > 0.0 && s < 1.0) return a * sqrt(-2.0 * log(s) / s);
  }
}
static double rng_gamma(source* src, double shape) { /* Marsaglia–Tsang; shape > 0 */
  if (shape < 1.0) {
    double u = rng_u01(src);
    while (u <= 0.0) u = rng_u01(src);
    return rng_gamma(src, shape + 1.0) * pow(u, 1.0 / shape);
  }
  double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (;;) {
    double x = rng_normal(src), v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    double u = rng_u01(src);
    if (u < 1.0 - 0.0331 * x * x * x * x) return d * v;
    if (u > 0.0 && log(u) < 0.5 * x * x + d * (1.0 - v + log(v))) return d * v;
  }
}

static inline double draw_uniform(source* src) {
  double v = src->use_rng ? rng_u01(src) : tape_next(src);
  record(src, v);
  return v;
}
static inline double draw_beta(source* src, double a, double b) { /* np.random.beta(a, b) */
  double v;
  if (src->use_rng) {
    double x = rng_gamma(src, a), y = rng_gamma(src, b);
    v = x / (x + y);
  } else {
    v = tape_next(src);
  }
  record(src, v);
  return v;
}
static inline int draw_pick(source* src, int n) { /* index into the FULL bullpen list */
  int idx;
  if (src->use_rng) {
    idx = (int)(rng_u01(src) * n);
    if (idx >= n) idx = n - 1;
  } else {
    idx = (int)tape_next(src);
    if (idx < 0) idx = 0;
    if (idx >= n) idx = n - 1;
  }
  record(src, (double)idx);
  return idx;
}

/* np.searchsorted(cdf, u, side='right') */
static inline int searchsorted_right(const double* cdf, int n, double u) {
  int i = 0;
  while (i < n && cdf[i] <= u) ++i;
  return i;
}

/* ------------------------------------------------------------------------------------------ */
/* float64 engine: one plate appearance (core/pa_simulator.py:91-143)                         */
/* ------------------------------------------------------------------------------------------ */
static int engine_pa(const mlbsim_params* P, source* src, int side, int slot, int pit, int tto, int pc) {
  /* core/fatigue_modeling.py:25-43 — two separate multiplies per rate, in this order */
  double pen = (tto == 2) ? 0.015 : (tto == 3) ? 0.035 : (tto >= 4) ? 0.060 : 0.0;
  double ak = P->pit_k[side][pit] * (1 - pen);
  double ab = P->pit_bb[side][pit] * (1 + pen);
  double fl = (double)(pc - 75) / 25.0;
  if (!(fl > 0.0)) fl = 0.0;
  double k_decay = 1.0 - 0.02 * fl;
  double bb_inflate = 1.0 + 0.03 * fl;
  ak = ak * k_decay;
  ab = ab * bb_inflate;
  /* core/pa_simulator.py:108-114 */
  double k = (P->bat_k[side][slot] + ak) / 2;
  double bb = (P->bat_bb[side][slot] + ab) / 2;
  k = k * P->k_mod;
  bb = bb * P->bb_mod;
  /* beta_noise, :23-32, :117-119 */
  double kp, bp;
  if (P->use_noise) {
    kp = (k <= 0.0) ? 0.0 : (k >= 1.0) ? 1.0 : draw_beta(src, k * 30, (1 - k) * 30);
    bp = (bb <= 0.0) ? 0.0 : (bb >= 1.0) ? 1.0 : draw_beta(src, bb * 30, (1 - bb) * 30);
  } else {
    kp = k; bp = bb;
  }
  bp = bp * 1.01;
  double u = draw_uniform(src);                     /* :122 */
  int base = (u < kp) ? 0 : (u < kp + bp) ? 1 : 2;  /* :35-41 */
  double uhr = draw_uniform(src);                   /* :126 — always consumed */
  int is_hr = uhr < P->pit_hr[side][pit];
  if (base == 0) return MLBSIM_K;
  if (base == 1) return MLBSIM_BB;
  if (is_hr) return MLBSIM_HR;
  /* resolve_contact, :51-88 */
  int type = searchsorted_right(P->cdf_bip, 4, draw_uniform(src));
  if (type > 3) type = 3;
  int hit = draw_uniform(src) < P->hit_prob[side][pit][slot][type]; /* stdlib stream, bip_resolution.py:62 */
  if (hit) return MLBSIM_1B + searchsorted_right(P->cdf_hit, 3, draw_uniform(src));
  if (draw_uniform(src) < 0.10) return MLBSIM_1B + searchsorted_right(P->cdf_fb, 2, draw_uniform(src));
  return MLBSIM_OUT;
}

static const uint8_t WALK_NS[8] = {1, 3, 5, 7, 5, 7, 5, 7}; /* half_inning_simulator.py:84-94 */

typedef struct { int pc, tto, pit; } staff_state;

/* one half inning (core/half_inning_simulator.py:138-258); returns runs */
static int engine_half(const mlbsim_params* P, source* src, int side, int* slot_io, staff_state* st, uint16_t* pa_counts) {
  int L = P->lineup_len[side];
  int outs = 0, runs = 0, n = 0, bases = 0, reached = 0;
  int bi = *slot_io;
  while (outs < 3 && n < 30) {
    if (bi > 0 && bi % L == 0) st->tto += 1;                                  /* :177-178 */
    int o = engine_pa(P, src, side, bi % L, st->pit, st->tto, st->pc);
    if (pa_counts) pa_counts[o]++;
    int r = 0;
    switch (o) {
      case MLBSIM_K:
      case MLBSIM_OUT:                                                          /* :73-82 */
        if ((bases & 1) && outs < 2 && draw_uniform(src) < 0.14) { bases &= ~1; outs += 2; }
        else outs += 1;
        break;
      case MLBSIM_BB:
        r = (bases == 7); bases = WALK_NS[bases];
        break;
      case MLBSIM_1B: {                                                         /* :97-109, :28-39 */
        int ns = 1, had2 = (bases >> 1) & 1;
        r = (bases >> 2) & 1;
        if (had2) { if (draw_uniform(src) < 0.4) r += 1; else ns |= 4; }
        if (bases & 1) { if (draw_uniform(src) < 0.8) ns |= 2; }
        if (outs == 2 && had2 && (ns & 4) && draw_uniform(src) < 0.10) { ns &= ~4; r += 1; }
        bases = ns;
        break;
      }
      case MLBSIM_2B: {                                                         /* :112-123 */
        int ns = 2;
        r = ((bases >> 2) & 1) + ((bases >> 1) & 1);
        if (bases & 1) { if (draw_uniform(src) < 0.4) r += 1; else ns |= 4; }
        bases = ns;
        break;
      }
      case MLBSIM_3B:
        r = __builtin_popcount(bases); bases = 4;
        break;
      case MLBSIM_HR:
        r = __builtin_popcount(bases) + 1; bases = 0;
        break;
    }
    runs += r;
    if (r > 0 || bases != 0) reached = 1;                                       /* :216-217 */
    st->pc += 1; bi += 1; n += 1;
  }
  if (reached && runs == 0 && draw_uniform(src) < 0.011) runs += 1;             /* :12-17 */
  if (reached && draw_uniform(src) < 0.01) runs += 1;                           /* :20-25 */
  *slot_io = bi % L;
  return runs;
}

typedef struct {
  int home_score, away_score, n_innings;
  int cap_a[4], cap_h[4];
  int n_used[2];
} game_result;

/* core/game_simulator.py:14-156 */
static void engine_game(const mlbsim_params* P, source* src, mlbsim_replay_game* out, game_result* res) {
  int score[2] = {0, 0};            /* [0] away, [1] home */
  int slot[2] = {0, 0};
  staff_state st[2] = {{0, 1, 0}, {0, 1, 0}};
  int inning = 1;
  int n_used[2] = {0, 0};
  if (out) memset(out, 0, sizeof(*out));
  for (;;) {
    int r[2] = {0, 0};
    for (int side = 0; side < 2; ++side) {
      if (side == 1 && inning == 9 && score[1] > score[0]) break;              /* :72 */
      if ((st[side].pc > 90 || st[side].tto >= 3) && P->bullpen_len[side] > 0) { /* :8-12, :48-54 */
        int idx = draw_pick(src, P->bullpen_len[side]);
        st[side].pit = 1 + idx; st[side].pc = 0; st[side].tto = 1;
        if (out) {
          uint8_t* used = side == 0 ? out->used_home : out->used_away;
          if (n_used[side] < MLBSIM_REPLAY_MAX_USED) used[n_used[side]] = (uint8_t)idx;
          else out->status |= MLBSIM_ST_USED_OVERFLOW;
        }
        n_used[side]++;
      }
      r[side] = engine_half(P, src, side, &slot[side], &st[side], out ? out->pa[side] : NULL);
      score[side] += r[side];
    }
    if (out) {
      if (inning <= MLBSIM_REPLAY_MAX_INNINGS) {
        out->away_runs[inning - 1] = (uint8_t)(r[0] > 255 ? 255 : r[0]);
        out->home_runs[inning - 1] = (uint8_t)(r[1] > 255 ? 255 : r[1]);
      } else out->status |= MLBSIM_ST_INNINGS_OVERFLOW;
    }
    if (res && inning <= 7 && (inning & 1)) { res->cap_a[inning >> 1] = score[0]; res->cap_h[inning >> 1] = score[1]; }
    if (inning >= 9 && score[0] != score[1]) break;                            /* :110-111 */
    if (src->underrun && inning >= 200) break; /* a dry tape (all 0.5) may never break a tie */
    inning += 1;
  }
  if (out) {
    out->home_score = score[1]; out->away_score = score[0]; out->n_innings = inning;
    out->n_used[0] = n_used[0]; out->n_used[1] = n_used[1];
  }
  if (res) {
    res->home_score = score[1]; res->away_score = score[0]; res->n_innings = inning;
    res->n_used[0] = n_used[0]; res->n_used[1] = n_used[1];
  }
}

/* ------------------------------------------------------------------------------------------ */
/* entry 1: replay                                                                            */
/* ------------------------------------------------------------------------------------------ */
typedef struct { const mlbsim_params* P; const double* tape; const uint64_t* offsets; mlbsim_replay_game* out; } replay_args;
static void replay_range(int64_t lo, int64_t hi, void* argp) {
  replay_args* a = (replay_args*)argp;
  for (int64_t g = lo; g < hi; ++g) {
    source src;
    memset(&src, 0, sizeof(src));
    src.tape = a->tape + a->offsets[g];
    src.len = (int64_t)(a->offsets[g + 1] - a->offsets[g]);
    engine_game(a->P, &src, &a->out[g], NULL);
    a->out[g].tape_used = (int32_t)src.pos;
    if (src.underrun) a->out[g].status |= MLBSIM_ST_TAPE_UNDERRUN;
  }
}
int oracle_replay(const mlbsim_params* P, const double* tape, const uint64_t* offsets, int64_t n_games,
                  mlbsim_replay_game* out) {
  replay_args a = {P, tape, offsets, out};
  parallel_ranges(0, n_games, replay_range, &a);
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* histogram helpers                                                                          */
/* ------------------------------------------------------------------------------------------ */
static inline int clampi(int v, int hi) { return v > hi ? hi : v; }

static void hist_add_game(mlbsim_hist* h, const int cap_a[4], const int cap_h[4], int a, int hm, int n_innings) {
  for (int c = 0; c < 4; ++c)
    h->joint[c][clampi(cap_a[c], MLBSIM_HIST_BINS - 1)][clampi(cap_h[c], MLBSIM_HIST_BINS - 1)]++;
  h->joint[4][clampi(a, MLBSIM_HIST_BINS - 1)][clampi(hm, MLBSIM_HIST_BINS - 1)]++;
  h->innings[clampi(n_innings, MLBSIM_INN_BINS - 1)]++;
  h->n_games++;
}

static void hist_merge(mlbsim_hist* dst, const mlbsim_hist* src) {
  uint64_t* d = (uint64_t*)dst;
  const uint64_t* s = (const uint64_t*)src;
  for (size_t i = 0; i < MLBSIM_HIST_WORDS; ++i) d[i] += s[i];
}

/* ------------------------------------------------------------------------------------------ */
/* entry 2: reference grammar with a CPU RNG (the cpu_baseline port)                          */
/* Games [game_lo, game_hi) are seeded by (seed, game index): results do not depend on the    */
/* thread count.  If tape_out != NULL the draws of game g are written at tape_out[g*stride..] */
/* and tape_len[g] receives the count (for large replay cross-checks).                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  const mlbsim_params* P; uint64_t seed; int64_t game_lo; mlbsim_hist* hist; mlbsim_replay_game* games_out;
  double* tape_out; int64_t tape_stride; int64_t* tape_len; int err;
} grammar_args;
static void grammar_range(int64_t lo, int64_t hi, void* argp) {
  grammar_args* a = (grammar_args*)argp;
  mlbsim_hist* local = (mlbsim_hist*)calloc(1, sizeof(mlbsim_hist));
  int err = 0;
  for (int64_t g = lo; g < hi; ++g) {
    source src;
    memset(&src, 0, sizeof(src));
    src.use_rng = 1;
    uint64_t sm = a->seed ^ (0xA0761D6478BD642Full * (uint64_t)(g + 1));
    for (int i = 0; i < 4; ++i) src.s[i] = splitmix64(&sm);
    if (a->tape_out) { src.rec = a->tape_out + (g - a->game_lo) * a->tape_stride; src.rec_cap = a->tape_stride; }
    mlbsim_replay_game rg;
    game_result res;
    memset(&res, 0, sizeof(res));
    engine_game(a->P, &src, &rg, &res);
    if (a->tape_len) a->tape_len[g - a->game_lo] = src.rec_pos;
    if (src.rec_overflow) err = 1;
    if (a->games_out) { rg.tape_used = (int32_t)src.rec_pos; a->games_out[g - a->game_lo] = rg; }
    hist_add_game(local, res.cap_a, res.cap_h, res.away_score, res.home_score, res.n_innings);
    for (int s = 0; s < 2; ++s) {
      for (int o = 0; o < 7; ++o) local->pa[s][o] += rg.pa[s][o];
      uint32_t mask = 0;
      const uint8_t* used = s == 0 ? rg.used_home : rg.used_away;
      int nu = rg.n_used[s] < MLBSIM_REPLAY_MAX_USED ? rg.n_used[s] : MLBSIM_REPLAY_MAX_USED;
      for (int i = 0; i < nu; ++i) { local->rel_picks[s][used[i]]++; mask |= 1u << used[i]; }
      for (int i = 0; i < 8; ++i) if (mask >> i & 1) local->rel_games[s][i]++;
    }
  }
  pthread_mutex_lock(&g_merge_lock);
  if (a->hist) hist_merge(a->hist, local);
  if (err) a->err = 1;
  pthread_mutex_unlock(&g_merge_lock);
  free(local);
}
int oracle_grammar(const mlbsim_params* P, uint64_t seed, int64_t game_lo, int64_t game_hi, mlbsim_hist* hist,
                   mlbsim_replay_game* games_out, double* tape_out, int64_t tape_stride, int64_t* tape_len) {
  grammar_args a = {P, seed, game_lo, hist, games_out, tape_out, tape_stride, tape_len, 0};
  parallel_ranges(game_lo, game_hi, grammar_range, &a);
  return a.err;
}

/* ------------------------------------------------------------------------------------------ */
/* entry 3: integer twin of the marginalised native law (what the CUDA kernel must reproduce   */
/* bit for bit).  Derivation of the law from the reference: DESIGN.md §"Native law".            */
/* ------------------------------------------------------------------------------------------ */
#define Q32(p) ((uint32_t)((p) * 4294967296.0 + 0.5))
static const uint32_t T_DP = 601295421u;      /* round(0.14 * 2^32)   half_inning_simulator.py:76 */
static const uint32_t T_040 = 1717986918u;    /* round(0.40 * 2^32)   :103, :120 */
static const uint32_t T_046 = 1975684956u;    /* round(0.46 * 2^32) = 0.4 + 0.6*0.10  :103 + :28-39 (two outs) */
static const uint32_t T_080_16 = 52429u;      /* round(0.80 * 2^16)   :105 (low half of w1) */
static const uint32_t T_END_BOTH = 472446u;   /* round(0.011*0.01 * 2^32)          misc AND ghost  :12-25 */
static const uint32_t T_END_ANY = 89721867u;  /* round((1-0.989*0.99) * 2^32)      misc OR ghost */
static const uint32_t T_END_GHOST = 42949673u;/* round(0.01 * 2^32)                ghost only (runs > 0) */

/* packed transition: new_bases | runs << 3 | outs_added << 6 */
static uint32_t native_transition(int o, int dA, int dB, int outs, int bases) {
  int nb = bases, r = 0, oa = 0;
  switch (o) {
    case MLBSIM_K: case MLBSIM_OUT:
      if ((bases & 1) && outs < 2 && dA) { nb = bases & ~1; oa = 2; } else oa = 1;
      break;
    case MLBSIM_BB: r = (bases == 7); nb = WALK_NS[bases]; break;
    case MLBSIM_1B: {
      int had2 = (bases >> 1) & 1;
      nb = 1; r = (bases >> 2) & 1;
      if (had2) { if (dA) r += 1; else nb |= 4; }
      if ((bases & 1) && dB) nb |= 2;
      break;
    }
    case MLBSIM_2B:
      nb = 2; r = ((bases >> 2) & 1) + ((bases >> 1) & 1);
      if (bases & 1) { if (dA) r += 1; else nb |= 4; }
      break;
    case MLBSIM_3B: r = __builtin_popcount(bases); nb = 4; break;
    case MLBSIM_HR: r = __builtin_popcount(bases) + 1; nb = 0; break;
  }
  return (uint32_t)nb | (uint32_t)r << 3 | (uint32_t)oa << 6;
}

/* one game of the native law; adds to `local` (if not NULL) and fills `rg` (if not NULL) */
static void native_game(const mlbsim_table* T, uint32_t matchup, uint32_t key, uint64_t sim, int count_pa,
                        mlbsim_hist* local, mlbsim_replay_game* rg) {
  const uint32_t c0 = (uint32_t)sim;
  const uint32_t c1base = ((uint32_t)(sim >> 32) << RNG_SIMHI_SHIFT) | (matchup << RNG_MATCHUP_SHIFT);
  uint32_t seq = 0;
  uint32_t w[2] = {0, 0}, w1_first = 0;
  int score[2] = {0, 0}, slot[2] = {0, 0};
  staff_state st[2] = {{0, 1, 0}, {0, 1, 0}};
  int cap_a[4] = {0, 0, 0, 0}, cap_h[4] = {0, 0, 0, 0};
  uint32_t used_mask[2] = {0, 0};
  int n_used[2] = {0, 0};
  int inning = 1;
  if (rg) memset(rg, 0, sizeof(*rg));
  for (;;) {
    int rr[2] = {0, 0};
    for (int side = 0; side < 2; ++side) {
      if (side == 1 && inning == 9 && score[1] > score[0]) break;
      int nbp = (int)T->bullpen_len[side];
      if ((st[side].pc > 90 || st[side].tto >= 3) && nbp > 0) {
        int idx = (int)(((w[1] & 0xFFFFu) * (uint32_t)nbp) >> 16); /* low half of the game's latest w1 (an out ignores it) */
        st[side].pit = 1 + idx; st[side].pc = 0; st[side].tto = 1;
        used_mask[side] |= 1u << idx;
        if (local) local->rel_picks[side][idx]++;
        if (rg) {
          uint8_t* used = side == 0 ? rg->used_home : rg->used_away;
          if (n_used[side] < MLBSIM_REPLAY_MAX_USED) used[n_used[side]] = (uint8_t)idx;
          else rg->status |= MLBSIM_ST_USED_OVERFLOW;
        }
        n_used[side]++;
      }
      int L = (int)T->lineup_len[side];
      int outs = 0, runs = 0, n = 0, bases = 0, reached = 0, sl = slot[side];
      while (outs < 3 && n < 30) {
        const uint32_t ctr[2] = {c0, c1base + seq};
        philox2x32_native(ctr, key, w);
        if (n == 0) w1_first = w[1]; /* bases empty: this w1 never reaches the transition -> end-of-half draw */
        if (sl == 0 && n > 0) st[side].tto += 1;
        int tier = (st[side].tto >= 4 ? 4 : st[side].tto) - 1;
        int o, dB;
        if ((T->flags & MLBSIM_FLAG_FATIGUE) && st[side].pit == 0 && st[side].pc > 75) {
          /* starter past 75 pitches: cumulative thresholds + per-pitch slope */
          const uint32_t* row = T->fthr[side][tier][sl];
          const int32_t* sp = T->fslope[side][tier][sl];
          uint32_t e = (uint32_t)(st[side].pc - 75);
          uint32_t u = w[0] >> 1;
          o = 0;
          for (int j = 0; j < 6; ++j) o += (u >= row[j] + e * (uint32_t)sp[j]);
          dB = (w[1] & 0xFFFFu) < T_080_16; /* runner on first takes second on a single: low half of w1 */
        } else {
          /* Walker alias draw over the eight variants: one table word per PA; the single arrives already split
             by the 0.8 draw (variant 7 = single on which the runner on first advances) */
          uint32_t col = w[0] >> 29;
          uint32_t e = T->alias[side][st[side].pit][tier][sl][col];
          o = ((uint32_t)(w[0] << 3) > e) ? (int)col : (int)(e & 7u); /* entry = (2^29 - 1 - cut29) << 3 | alias: exactly cut29 of the column's 2^29 draws keep it */
          dB = (o == MLBSIM_V_1B_ADV);
          if (dB) o = MLBSIM_1B;
        }
        uint32_t tA = (o == MLBSIM_1B) ? (outs == 2 ? T_046 : T_040) : (o == MLBSIM_2B) ? T_040 : T_DP;
        int dA = w[1] < tA;
        uint32_t t = native_transition(o, dA, dB, outs, bases);
        int r = (t >> 3) & 7;
        bases = t & 7; outs += t >> 6; runs += r;
        if (r > 0 || bases != 0) reached = 1;
        if (count_pa && local) local->pa[side][o]++;
        if (rg) rg->pa[side][o]++;
        st[side].pc += 1; n += 1; sl = (sl + 1 == L) ? 0 : sl + 1;
        seq += 1;
      }
      if (reached) runs += (runs == 0) ? (w1_first < T_END_BOTH) + (w1_first < T_END_ANY) : (w1_first < T_END_GHOST);
      slot[side] = sl;
      score[side] += runs;
      rr[side] = runs;
    }
    if (rg) {
      if (inning <= MLBSIM_REPLAY_MAX_INNINGS) {
        rg->away_runs[inning - 1] = (uint8_t)(rr[0] > 255 ? 255 : rr[0]);
        rg->home_runs[inning - 1] = (uint8_t)(rr[1] > 255 ? 255 : rr[1]);
      } else rg->status |= MLBSIM_ST_INNINGS_OVERFLOW;
    }
    if (inning <= 7 && (inning & 1)) { cap_a[inning >> 1] = score[0]; cap_h[inning >> 1] = score[1]; }
    if (inning >= 9 && score[0] != score[1]) break;
    if (inning >= MLBSIM_MAX_INNINGS) { /* termination guard (include/mlbsim.h): still tied, stop and count */
      if (local) local->n_truncated++;
      if (rg) rg->status |= MLBSIM_ST_TRUNCATED;
      break;
    }
    inning += 1;
  }
  if (local) {
    hist_add_game(local, cap_a, cap_h, score[0], score[1], inning);
    for (int s = 0; s < 2; ++s)
      for (int i = 0; i < 8; ++i) if (used_mask[s] >> i & 1) local->rel_games[s][i]++;
  }
  if (rg) {
    rg->home_score = score[1]; rg->away_score = score[0]; rg->n_innings = inning;
    rg->tape_used = (int32_t)seq; /* Philox PA blocks (= PAs) consumed */
    rg->n_used[0] = n_used[0]; rg->n_used[1] = n_used[1];
  }
}

typedef struct {
  const mlbsim_table* T; uint32_t matchup; uint64_t seed; mlbsim_hist* hist; uint32_t flags;
  mlbsim_replay_game* games; uint64_t sim_lo;
} native_args;
static void native_range(int64_t lo, int64_t hi, void* argp) {
  native_args* a = (native_args*)argp;
  const uint32_t key = oracle_native_key(a->seed);
  const int count_pa = !(a->flags & 1u);
  mlbsim_hist* local = a->hist ? (mlbsim_hist*)calloc(1, sizeof(mlbsim_hist)) : NULL;
  for (uint64_t sim = (uint64_t)lo; sim < (uint64_t)hi; ++sim)
    native_game(a->T, a->matchup, key, sim, count_pa, local, a->games ? &a->games[sim - a->sim_lo] : NULL);
  if (local) {
    pthread_mutex_lock(&g_merge_lock);
    hist_merge(a->hist, local);
    pthread_mutex_unlock(&g_merge_lock);
    free(local);
  }
}
/* sim indices must stay below 2^63 (they are split over threads as int64 ranges) */
int oracle_native(const mlbsim_table* T, uint32_t matchup, uint64_t sim_lo, uint64_t sim_hi, uint64_t seed,
                  mlbsim_hist* hist, uint32_t flags) {
  if (T->magic != MLBSIM_TABLE_MAGIC) return MLBSIM_ERR_TABLE;
  native_args a = {T, matchup, seed, hist, flags, NULL, sim_lo};
  parallel_ranges((int64_t)sim_lo, (int64_t)sim_hi, native_range, &a);
  return 0;
}
/* per-sim records of the native law (scores, per-inning runs, relievers, outcome counts) */
int oracle_native_persim(const mlbsim_table* T, uint32_t matchup, uint64_t sim_lo, uint64_t sim_hi, uint64_t seed,
                         mlbsim_replay_game* games) {
  if (T->magic != MLBSIM_TABLE_MAGIC) return MLBSIM_ERR_TABLE;
  native_args a = {T, matchup, seed, NULL, 0, games, sim_lo};
  parallel_ranges((int64_t)sim_lo, (int64_t)sim_hi, native_range, &a);
  return 0;
}

int oracle_abi_version(void) { return MLBSIM_ABI_VERSION; }
size_t oracle_sizeof_params(void) { return sizeof(mlbsim_params); }
size_t oracle_sizeof_table(void) { return sizeof(mlbsim_table); }
size_t oracle_sizeof_hist(void) { return sizeof(mlbsim_hist); }
size_t oracle_sizeof_replay_game(void) { return sizeof(mlbsim_replay_game); }
size_t oracle_sizeof_summary(void) { return sizeof(mlbsim_summary); }
