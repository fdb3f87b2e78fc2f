// replay_kernel.cu — tape-replay mode for sm_100a: one thread = one game, every random value is
// read from a recorded float64 tape in the reference's own draw order, every
// probability is the same IEEE double the reference compares against.  Integer results (runs per
// half-inning, final score, innings played, relievers used, per-game outcome counts) must be
// bit-exact with core/game_simulator.py:14-156 on the same draws.
//
// Floating point: this file is compiled with -fmad=false and all arithmetic that feeds a
// comparison goes through __dmul_rn/__dadd_rn/__ddiv_rn, so nvcc can never contract
// `u < k_prob + bb_prob * 1.01`-style expressions into an FMA (core/pa_simulator.py:119-124).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mlbsim.h"
#include "native_kernel.h"

#ifndef REPLAY_L2_PREFETCH
#define REPLAY_L2_PREFETCH 256 /* L2 prefetch-size hint of the tape loads: 0 (none), 128 or 256 bytes; 256 = +5 % measured */
#endif

namespace {

// Sequential reader of one game's float64 tape: one 8-byte read-only load per draw.  Consecutive
// draws of a lane fall in the same 32-byte sector / 128-byte line, which L1 keeps between them as
// long as the per-thread footprint stays small — hence no per-thread record in local memory
// (results go straight to global memory) and a modest number of resident threads per SM.
// (Measured on B200, 400 k games: an explicit 4-double register buffer with prefetch was 4.5x
// slower than this — 86 registers and a divergent refill at every call site.)
struct Tape {
  const double* p;
  uint32_t len;
  uint32_t pos;
  __device__ __forceinline__ void init(const double* tape, uint64_t, uint64_t lo, uint64_t hi) {
    p = tape + lo; len = (uint32_t)(hi - lo); pos = 0;
  }
  __device__ __forceinline__ double next() {
    // past-the-end reads return 0.5 and are flagged by the caller through pos > len
    double v = 0.5;
    if (pos < len) {
#if REPLAY_L2_PREFETCH == 256
      asm volatile("ld.global.nc.L2::256B.f64 %0, [%1];" : "=d"(v) : "l"(p + pos));
#elif REPLAY_L2_PREFETCH == 128
      asm volatile("ld.global.nc.L2::128B.f64 %0, [%1];" : "=d"(v) : "l"(p + pos));
#else
      v = __ldg(p + pos);
#endif
    }
    ++pos;
    return v;
  }
};

// np.searchsorted(cdf, u, side='right')
template <int N>
__device__ __forceinline__ int searchsorted_right(const double* cdf, double u) {
  int i = 0;
#pragma unroll
  for (int j = 0; j < N; ++j) i += (cdf[j] <= u) ? 1 : 0;
  return i;
}

struct Staff {
  int pc, tto, pit;
};

// core/pa_simulator.py:91-143 with core/fatigue_modeling.py:25-43 in front
__device__ int replay_pa(const mlbsim_params& P, Tape& tp, int side, int slot, const Staff& st) {
  const int tto = st.tto;
  const double pen = (tto == 2) ? 0.015 : (tto == 3) ? 0.035 : (tto >= 4) ? 0.060 : 0.0;
  double ak = __dmul_rn(P.pit_k[side][st.pit], __dsub_rn(1.0, pen));
  double ab = __dmul_rn(P.pit_bb[side][st.pit], __dadd_rn(1.0, pen));
  double fl = __ddiv_rn((double)(st.pc - 75), 25.0);
  if (!(fl > 0.0)) fl = 0.0;
  const double k_decay = __dsub_rn(1.0, __dmul_rn(0.02, fl));
  const double bb_inflate = __dadd_rn(1.0, __dmul_rn(0.03, fl));
  ak = __dmul_rn(ak, k_decay);
  ab = __dmul_rn(ab, bb_inflate);
  double k = __ddiv_rn(__dadd_rn(P.bat_k[side][slot], ak), 2.0);
  double bb = __ddiv_rn(__dadd_rn(P.bat_bb[side][slot], ab), 2.0);
  k = __dmul_rn(k, P.k_mod);
  bb = __dmul_rn(bb, P.bb_mod);
  double kp, bp;
  if (P.use_noise) {  // beta_noise: no draw when p <= 0 or p >= 1
    kp = (k <= 0.0) ? 0.0 : (k >= 1.0) ? 1.0 : tp.next();
    bp = (bb <= 0.0) ? 0.0 : (bb >= 1.0) ? 1.0 : tp.next();
  } else {
    kp = k;
    bp = bb;
  }
  bp = __dmul_rn(bp, 1.01);
  const double u = tp.next();
  const int base = (u < kp) ? 0 : (u < __dadd_rn(kp, bp)) ? 1 : 2;
  const double uhr = tp.next();  // always consumed (pa_simulator.py:126)
  const bool is_hr = uhr < P.pit_hr[side][st.pit];
  if (base == 0) return MLBSIM_K;
  if (base == 1) return MLBSIM_BB;
  if (is_hr) return MLBSIM_HR;
  int type = searchsorted_right<4>(P.cdf_bip, tp.next());
  if (type > 3) type = 3;
  const bool hit = tp.next() < P.hit_prob[side][st.pit][slot][type];
  if (hit) return MLBSIM_1B + searchsorted_right<3>(P.cdf_hit, tp.next());
  if (tp.next() < 0.10) return MLBSIM_1B + searchsorted_right<2>(P.cdf_fb, tp.next());
  return MLBSIM_OUT;
}

// core/half_inning_simulator.py:138-258
__device__ __forceinline__ int replay_half(const mlbsim_params& P, Tape& tp, int side, int& slot_io, Staff& st,
                                            unsigned long long& pa_counts) {
  const int L = P.lineup_len[side];
  int outs = 0, runs = 0, n = 0, bases = 0;
  bool reached = false;
  int bi = slot_io;
  while (outs < 3 && n < 30) {
    if (bi > 0 && bi % L == 0) st.tto += 1;
    const int o = replay_pa(P, tp, side, bi % L, st);
    pa_counts += 1ull << (9 * o);   // 9-bit fields: up to 511 of one outcome per side per game
    int r = 0;
    if (o == MLBSIM_K || o == MLBSIM_OUT) {
      if ((bases & 1) && outs < 2 && tp.next() < 0.14) { bases &= ~1; outs += 2; }
      else outs += 1;
    } else if (o == MLBSIM_BB) {
      r = (bases == 7);
      bases = 1 | ((bases & 1) << 1) | ((((bases >> 1) | (bases >> 2)) & 1) << 2);
    } else if (o == MLBSIM_1B) {
      int ns = 1;
      const int had2 = (bases >> 1) & 1;
      r = (bases >> 2) & 1;
      if (had2) { if (tp.next() < 0.4) r += 1; else ns |= 4; }
      if (bases & 1) { if (tp.next() < 0.8) ns |= 2; }
      if (outs == 2 && had2 && (ns & 4) && tp.next() < 0.10) { ns &= ~4; r += 1; }
      bases = ns;
    } else if (o == MLBSIM_2B) {
      int ns = 2;
      r = ((bases >> 2) & 1) + ((bases >> 1) & 1);
      if (bases & 1) { if (tp.next() < 0.4) r += 1; else ns |= 4; }
      bases = ns;
    } else if (o == MLBSIM_3B) {
      r = __popc(bases); bases = 4;
    } else {
      r = __popc(bases) + 1; bases = 0;
    }
    runs += r;
    if (r > 0 || bases != 0) reached = true;
    st.pc += 1; bi += 1; n += 1;
  }
  if (reached && runs == 0 && tp.next() < 0.011) runs += 1;
  if (reached && tp.next() < 0.01) runs += 1;
  slot_io = bi % L;
  return runs;
}

}  // namespace

#ifndef REPLAY_MIN_BLOCKS
#define REPLAY_MIN_BLOCKS 8
#endif
extern "C" __global__ void __launch_bounds__(128, REPLAY_MIN_BLOCKS)
mlbsim_replay_kernel(const ReplayArgs args) {
  __shared__ mlbsim_params P;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(args.params);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&P);
    for (int i = threadIdx.x; i < (int)(sizeof(mlbsim_params) / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const uint64_t total = args.offsets[args.n_games];

  // `out` was zeroed by the launcher: only non-zero fields are written, straight to global memory
  // (no per-thread record in local memory).
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < args.n_games;
       g += (long long)gridDim.x * blockDim.x) {
    Tape tp;
    tp.init(args.tape, total, args.offsets[g], args.offsets[g + 1]);
    mlbsim_replay_game* const R = args.out + g;
    int status = 0;
    int score0 = 0, score1 = 0, slot0 = 0, slot1 = 0, used0 = 0, used1 = 0;
    Staff st0 = {0, 1, 0}, st1 = {0, 1, 0};
    unsigned long long pa0 = 0, pa1 = 0;
    int inning = 1;
    for (;;) {
      int r0 = 0, r1 = 0;
      // ---- top: away bats (side 0) against the home staff ----
      if ((st0.pc > 90 || st0.tto >= 3) && P.bullpen_len[0] > 0) {
        int idx = (int)tp.next();
        const int nb = P.bullpen_len[0];
        idx = idx < 0 ? 0 : (idx >= nb ? nb - 1 : idx);
        st0.pit = 1 + idx; st0.pc = 0; st0.tto = 1;
        if (used0 < MLBSIM_REPLAY_MAX_USED) R->used_home[used0] = (uint8_t)idx; else status |= MLBSIM_ST_USED_OVERFLOW;
        used0++;
      }
      r0 = replay_half(P, tp, 0, slot0, st0, pa0);
      score0 += r0;
      // ---- bottom: home bats (side 1) unless it leads after the top of the 9th ----
      if (!(inning == 9 && score1 > score0)) {
        if ((st1.pc > 90 || st1.tto >= 3) && P.bullpen_len[1] > 0) {
          int idx = (int)tp.next();
          const int nb = P.bullpen_len[1];
          idx = idx < 0 ? 0 : (idx >= nb ? nb - 1 : idx);
          st1.pit = 1 + idx; st1.pc = 0; st1.tto = 1;
          if (used1 < MLBSIM_REPLAY_MAX_USED) R->used_away[used1] = (uint8_t)idx; else status |= MLBSIM_ST_USED_OVERFLOW;
          used1++;
        }
        r1 = replay_half(P, tp, 1, slot1, st1, pa1);
        score1 += r1;
      }
      if (inning <= MLBSIM_REPLAY_MAX_INNINGS) {
        if (r0) R->away_runs[inning - 1] = (uint8_t)(r0 > 255 ? 255 : r0);
        if (r1) R->home_runs[inning - 1] = (uint8_t)(r1 > 255 ? 255 : r1);
      } else {
        status |= MLBSIM_ST_INNINGS_OVERFLOW;
      }
      if (inning >= 9 && score0 != score1) break;
      // a tape that ran dry cannot end the game by itself (0.5 draws): stop after a bound
      if (tp.pos > tp.len && inning >= 200) break;
      inning += 1;
    }
    if (tp.pos > tp.len) status |= MLBSIM_ST_TAPE_UNDERRUN;
    int4 head0 = make_int4(score1, score0, inning, (int)tp.pos);
    int4 head1 = make_int4(used0, used1, status, 0);
    reinterpret_cast<int4*>(R)[0] = head0;
    reinterpret_cast<int4*>(R)[1] = head1;
#pragma unroll
    for (int o = 0; o < MLBSIM_N_OUTCOMES; ++o) {
      const uint16_t c0 = (uint16_t)((pa0 >> (9 * o)) & 511ull), c1 = (uint16_t)((pa1 >> (9 * o)) & 511ull);
      if (c0) R->pa[0][o] = c0;
      if (c1) R->pa[1][o] = c1;
    }
  }
}
