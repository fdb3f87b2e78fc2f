// replay_flat_kernel.cu — tape-replay mode (bit-exact verification against the reference's recorded draws).
//
// Contract: bit-exact integers on the reference's recorded draws (the fields of core/game_simulator.py:134-144),
// float64 compares, no FMA contraction.  Organised like the native kernel: the warp runs in rounds of REPLAY_UNROLL
// plate-appearance steps followed by one boundary pass, lanes refill by ballot prefix.
//
// The tape is the only HBM traffic (~535 float64 per game).  The reference's draw grammar
//     [beta_k] [beta_bb] u u_hr [u_bip u_hit (u_hit3 | u_fb [u_fb2])] [h1] [h2] [h3]
// makes WHICH value a draw is depend on the values before it, but not WHERE the next few values are: a plate
// appearance consumes at most eleven consecutive tape entries.
//
// Memory path (MLBSIM_REPLAY_RING = 1, the product): every lane streams its game's tape through a 256-byte ring
// of its own in shared memory — eight 32-byte granules, filled with cp.async (LDGSTS, L1 bypassed) up to 256 bytes ahead
// of the read position, so every tape byte comes from HBM exactly once and its latency is paid a step ahead of its
// use.  A plate appearance reads its entries from the ring with LDS.64: the eight leading entries at fixed offsets
// (independent loads), the three base-running draws at offsets the grammar computes.  The last entries of a game are
// patched to 0.5 in the ring (what a read past the end of the tape must yield), so no read in the step looks at the
// tape length.  800 threads per SM (one block), 200 KB of rings.  Measured: 5.7e8 games/s against 5.1e8 for the window
// loads it replaces (profiles/r02_rp5 .. rp9.log); the kernel is now bound by instruction issue and dependent latencies
// (issue active 61 %, long-scoreboard stalls 0.7 per issue, was 5.9), no longer by L1 wavefronts.
//
// MLBSIM_REPLAY_RING = 0 (kept for A/B): every step issues eleven INDEPENDENT 8-byte loads of the window
// tape[pos .. pos+10] straight from global memory — as aligned 16-byte units, because a lane's load costs one L1
// wavefront whatever its width — and the grammar then only selects among registers.  Degenerate rates (k or bb outside
// (0, 1): no Beta draw, core/pa_simulator.py:23-32) fall back to draw-by-draw reads in both variants.
//
// The per-PA K / BB rates of every (side, pitcher, TTO tier, batter) up to 75 pitches are tabulated in shared memory at
// block start with the reference's operation order (core/fatigue_modeling.py:25-43, core/pa_simulator.py:108-114); past
// 75 pitches the pitcher's side comes from two factored tables (base rate x TTO penalty, fatigue ramp by pitch count),
// again the reference's own operations in its order, so no step divides.
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/mlbsim.h"
#include "game_rules.cuh"
#include "native_kernel.h"

#ifndef REPLAY_UNROLL
#define REPLAY_UNROLL 3   /* plate-appearance steps per round (ring kernel: 2 / 3 / 4 / 5 measured, profiles/r02_rp6.log, r02_rp7.log) */
#endif
#ifndef REPLAY_STRAIGHT
#define REPLAY_STRAIGHT MLBSIM_REPLAY_RING   /* 1: the plate appearance as straight-line code (every compare evaluated, bit operations, split points from the constant bank) */
#endif
#ifndef REPLAY_RATE_TABLES
#define REPLAY_RATE_TABLES 3   /* 1: final rates tabulated up to 75 pitches, computed beyond; 2: factored tables (six float64 operations per plate
                                  appearance) at every pitch count: -2.6 % on a matchup with bullpens, where nobody passes 75 pitches
                                  (profiles/r02_rp8.log); 3: final rates up to 75 pitches, factored tables beyond */
#endif
#ifndef REPLAY_CONST_BANK
#define REPLAY_CONST_BANK 0  /* 1: the step's float64 literals that have no immediate form (1.01, 0.1, 0.4, 0.14, 0.8) come from the kernel-argument
                                constant bank as compare operands instead of being built in registers every step */
#endif
#if REPLAY_CONST_BANK
#define RC_101 args.consts[0]
#define RC_010 args.consts[1]
#define RC_040 args.consts[2]
#define RC_014 args.consts[3]
#define RC_080 args.consts[4]
#else
#define RC_101 1.01
#define RC_010 0.10
#define RC_040 0.4
#define RC_014 0.14
#define RC_080 0.8
#endif
#ifndef REPLAY_OFFSET_CACHE
#define REPLAY_OFFSET_CACHE 0   /* 1: the tape offsets of the warp's next 32 games sit in registers (one game per lane, handed out by shuffle)
                                and the first lines of their tapes are asked into L2 when the window is loaded: a lane that starts a game no longer
                                pays two dependent trips to HBM (offsets, then the first granules) with the whole warp waiting behind it.
                                Measured: -0.5 % (6.48 vs 6.51e8, profiles/r02_rp15.log) — the offsets are L2 hits already */
#endif
#ifndef REPLAY_FILL_FIRST
#define REPLAY_FILL_FIRST 0   /* 1: a step first asks for its new granules (their own cp.async group), then waits for everything but that group:
                                 the requests leave before the lane stalls on the older ones instead of after */
#endif
#ifndef REPLAY_INLINE_BOUNDARY
#define REPLAY_INLINE_BOUNDARY 0   /* 1 (ring kernel): a half inning that ends inside a round goes on into the next half right there when nothing rare
                                      happens (no game over, no pitcher change, entries already in the ring, not at the end of the tape); the rare
                                      cases wait for the round-end pass as before.  Lanes no longer idle for the rest of the round.  Bit-exact; +0.9 / +1.0 / -1.5 % at
                                      3 / 4 / 5 steps per round (profiles/r02_rp18.log): the inline block costs ~96 instructions per step — off */
#endif
#if REPLAY_INLINE_BOUNDARY && !defined(REPLAY_BOUNDARY_SKIP_WAIT)
#define REPLAY_BOUNDARY_SKIP_WAIT 1   /* (the inline boundary needs the landed-granule bookkeeping) */
#endif
#ifndef REPLAY_BOUNDARY_SKIP_WAIT
#define REPLAY_BOUNDARY_SKIP_WAIT 0   /* 1: the boundary pass waits for outstanding cp.async copies only when one of its (at most three) entries lies in a
                                         granule asked for at the lane's latest fill; everything older landed at that step's own wait.  +0.7 % (6.50 vs 6.45e8,
                                         profiles/r02_rp17.log): the copies are tracked per warp, so the wait goes only when no lane needs it — off */
#endif
#ifndef REPLAY_STEP_LOOP
#define REPLAY_STEP_LOOP 0   /* 1: the steps of a round as a real loop (a quarter of the code) instead of unrolled copies */
#endif
#ifndef REPLAY_FILL_NESTED
#define REPLAY_FILL_NESTED 1 /* ring fill as nested early exits with immediate offsets instead of a loop with pointer increments */
#endif
#ifndef REPLAY_FILL_UNROLL
#define REPLAY_FILL_UNROLL 0 /* ring fill: granule requests laid out straight-line per step before the general loop (0: the loop alone, +4.5 %: less code, profiles/r02_rp6.log) */
#endif
#define REPLAY_FLAT_THREADS MLBSIM_REPLAY_FLAT_THREADS
#define REPLAY_FLAT_BLOCKS MLBSIM_REPLAY_FLAT_BLOCKS

namespace {

constexpr int LUT_WORDS = 24 * 8 * 4;  // [outs*8+bases][outcome][dA][dB], same layout as the native kernel

// context word: slot[3:0] | tier[5:4] (min(tto,4)-1) | pit[8:6] | pitch_count[31:9]
constexpr uint32_t CX_SLOT = 0xFu;
constexpr uint32_t CX_TIER_ONE = 1u << 4;
constexpr uint32_t CX_TIER_MASK = 3u << 4;
constexpr uint32_t CX_TIER_GE2 = 2u << 4;
constexpr int CX_PIT_SHIFT = 6;
constexpr int CX_PC_SHIFT = 9;
constexpr uint32_t CX_NEXT_PA = (1u << CX_PC_SHIFT) + 1u;
// half state: ob[4:0] | n_pa + 34 [14:8] | runs[21:16] | reached[29:24] | three outs [30]
constexpr uint32_t HS_OB = 31u;
constexpr uint32_t HS_NEW_HALF = (64u - 30u) << 8;
constexpr uint32_t HS_PA_ONE = 1u << 8;
constexpr uint32_t HS_OVER = (1u << 30) | (1u << 14);
constexpr uint32_t HS_REACHED = 0x3Fu << 24;
constexpr uint32_t HS_DEAD = 1u << 15;
constexpr uint32_t HS_PARKED = 1u << 31;   // (inline boundary) half over, left to the round-end pass

constexpr int KB_ENTRIES = 2 * MLBSIM_MAX_PITCHERS * MLBSIM_N_TIERS * MLBSIM_MAX_LINEUP;   // 504

constexpr int BASE_ENTRIES = 2 * MLBSIM_MAX_PITCHERS * MLBSIM_N_TIERS;   // 56
constexpr int FAT_ENTRIES = 64;                                           // pitch counts 75 .. 138 (beyond: computed on the spot)

struct FlatSmem {
  mlbsim_params P;
  uint32_t lut[LUT_WORDS];
#if REPLAY_RATE_TABLES >= 2
  // The pitcher's side of the K / BB rates, factored the way core/fatigue_modeling.py:25-43 multiplies it:
  // base[side][pit][tier] = {pit_k * (1 - tto_penalty), pit_bb * (1 + tto_penalty)}, fat[e] = {1 - 0.02 * fl, 1 + 0.03 * fl} with
  // fl = max(pitch_count - 75, 0) / 25 — every entry the reference's own operations in the reference's order, so that a
  // plate appearance at any pitch count is two table reads and the remaining six float64 operations (a starter spends
  // 15 of his ~90 plate appearances past 75 pitches: computing fl there took a float64 division in most steps).
  double2 base[BASE_ENTRIES];
  double2 fat[FAT_ENTRIES];
#endif
#if REPLAY_RATE_TABLES != 2
  double2 kb[KB_ENTRIES];   // {k, bb} before noise at pitch_count <= 75: [side][pit][tier][slot]
#endif
};

// effective K / BB rates of one plate appearance before noise: fatigue-adjusted pitcher rate averaged with the batter's,
// times the umpire modifiers — core/fatigue_modeling.py:25-43 (two separate multiplies per rate, in this order), then
// core/pa_simulator.py:108-114
__device__ __forceinline__ double2 effective_rates(const mlbsim_params& P, uint32_t side, uint32_t pit, uint32_t tier, uint32_t slot, int pc) {
  const double pen = tier == 0 ? 0.0 : tier == 1 ? 0.015 : tier == 2 ? 0.035 : 0.060;
  double ak = __dmul_rn(P.pit_k[side][pit], __dsub_rn(1.0, pen));
  double ab = __dmul_rn(P.pit_bb[side][pit], __dadd_rn(1.0, pen));
  double fl = __ddiv_rn((double)(pc - 75), 25.0);
  if (!(fl > 0.0)) fl = 0.0;
  ak = __dmul_rn(ak, __dsub_rn(1.0, __dmul_rn(0.02, fl)));
  ab = __dmul_rn(ab, __dadd_rn(1.0, __dmul_rn(0.03, fl)));
  double2 r;
  r.x = __dmul_rn(__ddiv_rn(__dadd_rn(P.bat_k[side][slot], ak), 2.0), P.k_mod);
  r.y = __dmul_rn(__ddiv_rn(__dadd_rn(P.bat_bb[side][slot], ab), 2.0), P.bb_mod);
  return r;
}

__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// window read: entry pos + i of the lane's tape, 0.5 past the end (flagged later through pos > len)
#ifndef REPLAY_PREFETCH
#define REPLAY_PREFETCH 32  /* > 0: the plate-appearance window also prefetches the line this many entries ahead of it (REPLAY_PREFETCH_LEVEL
                               1 = into L1, 2 = into L2); measured +3 % at 32 entries into L1, nothing at 16, worse into L2
                               (profiles/r02_rp4_prefetch.log) */
#endif
#ifndef REPLAY_PREFETCH_LEVEL
#define REPLAY_PREFETCH_LEVEL 1
#endif
#ifndef REPLAY_VEC
#define REPLAY_VEC 2   /* tape entries per window load: 1 = LDG.64, 2 = LDG.128, 4 = LDG.256 (aligned units) */
#endif

// N consecutive entries from `pos` as ALIGNED vector loads: every lane's load costs one L1 wavefront whatever its width,
// so 16- or 32-byte units cut the wavefronts of a window by that factor; the lane then shifts the window into place by
// its misalignment (selects in registers).  Unguarded: the tape buffer is readable to the end of the 128-byte chunk
// holding its last draw (mlbsim_run_replay pads its copy; mlbsim_run_replay_dev asks the caller to); entries at or
// past the game's end read as 0.5 exactly like `take` — only a lane within N entries of its game's end pays for that.
template <int N>
__device__ __forceinline__ void peek_window(const double* p, uint32_t pos, uint32_t len, double (&w)[N]) {
  const double* q = p + pos;
#if REPLAY_PREFETCH > 0
  if (N >= 9) {
#if REPLAY_PREFETCH_LEVEL == 1
    asm volatile("prefetch.global.L1 [%0];" ::"l"(pos + REPLAY_PREFETCH < len ? q + REPLAY_PREFETCH : q));   // (stays inside this game's tape)
#else
    asm volatile("prefetch.global.L2 [%0];" ::"l"(pos + REPLAY_PREFETCH < len ? q + REPLAY_PREFETCH : q));
#endif
  }
#endif
#if REPLAY_VEC == 1
#pragma unroll
  for (int i = 0; i < N; ++i) asm volatile("ld.global.nc.L2::256B.f64 %0, [%1];" : "=d"(w[i]) : "l"(q + i));
#elif REPLAY_VEC == 2
  const uint32_t off = (uint32_t)(((uintptr_t)q >> 3) & 1u);      // 0 or 1 entries past a 16-byte boundary
  const double* a = q - off;
  constexpr int U = (N + 2) / 2;                                   // units covering [q - 1, q + N)
  double buf[2 * U];
#pragma unroll
  for (int u = 0; u < U; ++u)
    asm volatile("ld.global.nc.L2::256B.v2.f64 {%0, %1}, [%2];" : "=d"(buf[2 * u]), "=d"(buf[2 * u + 1]) : "l"(a + 2 * u));
#pragma unroll
  for (int i = 0; i < N; ++i) w[i] = off ? buf[i + 1] : buf[i];
#else
  const uint32_t off = (uint32_t)(((uintptr_t)q >> 3) & 3u);      // 0..3 entries past a 32-byte boundary
  const double* a = q - off;
  constexpr int U = (N + 3 + 3) / 4;                               // units covering [q - 3, q + N)
  double buf[4 * U];
#pragma unroll
  for (int u = 0; u < U; ++u)
    asm volatile("ld.global.nc.L2::256B.v4.f64 {%0, %1, %2, %3}, [%4];"
                 : "=d"(buf[4 * u]), "=d"(buf[4 * u + 1]), "=d"(buf[4 * u + 2]), "=d"(buf[4 * u + 3]) : "l"(a + 4 * u));
  double t[N + 1];
#pragma unroll
  for (int i = 0; i < N + 1; ++i) t[i] = (off & 2u) ? buf[i + 2] : buf[i];
#pragma unroll
  for (int i = 0; i < N; ++i) w[i] = (off & 1u) ? t[i + 1] : t[i];
#endif
  if (pos + (uint32_t)N > len) {
#pragma unroll
    for (int i = 0; i < N; ++i)
      if (pos + (uint32_t)i >= len) w[i] = 0.5;
  }
}

// predicated tape read: past-the-end reads yield 0.5 (flagged later through pos > len)
__device__ __forceinline__ double take(const double* p, uint32_t& pos, uint32_t len, bool want) {
  double v = 0.5;
  if (want) {
    if (pos < len) asm volatile("ld.global.nc.L2::256B.f64 %0, [%1];" : "=d"(v) : "l"(p + pos));
    ++pos;
  }
  return v;
}

#if MLBSIM_REPLAY_RING
// ---- per-lane tape ring -------------------------------------------------------------------------
// Every lane streams its game's tape through RING_BYTES of shared memory of its own: eight 32-byte granules, granule g
// of the lane's (32-byte aligned) tape in slot g & 7.  cp.async (LDGSTS, 16 bytes each, L1 bypassed) moves every byte
// from HBM exactly once, asked for up to 256 bytes ahead of the read position; the plate-appearance window is then
// read with LDS.128 — no L1 wavefront per 8 useful bytes, no window re-reads through a thrashing L1, and the global
// latency is paid a step ahead instead of inside the step.
constexpr uint32_t RING_BYTES = MLBSIM_REPLAY_RING_LANE_BYTES;   // 256
constexpr uint32_t RING_GRANULES = RING_BYTES / 32u;              // 8
static_assert(RING_BYTES == 256, "ring addressing below assumes 16 units of 16 bytes");

#ifndef REPLAY_RING_L2_HINT
#define REPLAY_RING_L2_HINT 256   /* L2 prefetch size carried by the ring's cp.async copies: 64 / 128 / 256 bytes -> 5.03 / 6.32 / 6.51e8 games/s (profiles/r02_rp14.log) */
#endif
#ifndef REPLAY_RING_PREFETCH
#define REPLAY_RING_PREFETCH 0    /* > 0: every fill also asks L2 for the line this many bytes past the last granule it requested */
#endif
#define RP_STR2(x) #x
#define RP_STR(x) RP_STR2(x)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global.L2::" RP_STR(REPLAY_RING_L2_HINT) "B [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all_but_last() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void lds128_f64(uint32_t addr, double& a, double& b) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}

// granule `g` of the lane's tape -> its ring slot (two 16-byte copies)
__device__ __forceinline__ void ring_fetch(uint32_t ring, const double* p, uint32_t g) {
  const uint32_t dst = ring | ((g & (RING_GRANULES - 1u)) << 5);
  const double* src = p + ((size_t)g << 2);
  cp_async16(dst, src);
  cp_async16(dst + 16u, src + 2);
}

// the same with the granule number as an immediate offset from a base (`so` = byte offset of the base granule, unwrapped)
__device__ __forceinline__ void ring_fetch_at(uint32_t ring, uint32_t so, const double* src, int k) {
  const uint32_t dst = ring | ((so + 32u * (uint32_t)k) & (RING_BYTES - 32u));
  cp_async16(dst, src + 4 * k);
  cp_async16(dst + 16u, src + 4 * k + 2);
}

// tape entry `idx` of the lane, out of its ring (resident: see the fill rule in the step)
__device__ __forceinline__ double ring_at(uint32_t ring, uint32_t idx) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(ring | ((idx << 3) & (RING_BYTES - 8u))));
  return v;
}

// A window that reaches past the game's last draw (the last plate appearance or two of a game): the entries past the end
// must read 0.5 (flagged later through pos > len).  Their ring slots hold whatever follows the game in the tape buffer,
// or nothing yet — put 0.5 there, so that no read in the step has to look at `len`.  (Everything asked for has landed
// before this is called and nothing past the game's last granule is ever asked for: the patch stays.)
template <int N>
__device__ __forceinline__ void ring_patch_tail(uint32_t ring, uint32_t pos, uint32_t len) {
#pragma unroll
  for (int i = 0; i < N; ++i)
    if (pos + (uint32_t)i >= len) asm volatile("st.shared.f64 [%0], %1;" ::"r"(ring | (((pos + (uint32_t)i) << 3) & (RING_BYTES - 8u))), "d"(0.5) : "memory");
}

__device__ __forceinline__ double ring_take(uint32_t ring, uint32_t& pos, uint32_t len, bool want) {
  double v = 0.5;
  if (want) {
    if (pos < len) v = ring_at(ring, pos);
    ++pos;
  }
  return v;
}
#endif

// REGULAR (host-verified, mlbsim_api.cu `replay_params_regular`): noise on, both staffs have a bullpen — so no pitcher ever
// passes 120 pitches (game_simulator.py:8-12 takes him out at the first half-inning boundary past 90) — and every K / BB rate
// up to there lies strictly inside (0, 1): every plate appearance draws both Beta values, the draw-by-draw path and the
// on-the-spot rate computation are not instantiated (2 096 instead of 3 096 SASS instructions, 64 instead of 72 registers).
template <bool NOISE, bool REGULAR>
__device__ __forceinline__ void replay_body(const ReplayArgs& args, unsigned long long* counter, uint32_t chunk) {
  __shared__ FlatSmem S;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(args.params);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&S.P);
    for (int i = threadIdx.x; i < (int)(sizeof(mlbsim_params) / 4); i += blockDim.x) dst[i] = src[i];
    for (int i = threadIdx.x; i < LUT_WORDS; i += blockDim.x) {
      const int ob = i >> 5, o = (i >> 2) & 7, dA = (i >> 1) & 1, dB = i & 1;
      S.lut[i] = transition_entry(o < 7 ? o : MLBSIM_OUT, dA, dB, ob >> 3, ob & 7);
    }
  }
  __syncthreads();
#if REPLAY_RATE_TABLES >= 2
  for (int i = threadIdx.x; i < BASE_ENTRIES; i += blockDim.x) {
    const uint32_t tier = (uint32_t)i % MLBSIM_N_TIERS, sp = (uint32_t)i / MLBSIM_N_TIERS;
    const double pen = tier == 0 ? 0.0 : tier == 1 ? 0.015 : tier == 2 ? 0.035 : 0.060;
    S.base[i] = make_double2(__dmul_rn(S.P.pit_k[sp / MLBSIM_MAX_PITCHERS][sp % MLBSIM_MAX_PITCHERS], __dsub_rn(1.0, pen)),
                             __dmul_rn(S.P.pit_bb[sp / MLBSIM_MAX_PITCHERS][sp % MLBSIM_MAX_PITCHERS], __dadd_rn(1.0, pen)));
  }
  for (int e = threadIdx.x; e < FAT_ENTRIES; e += blockDim.x) {
    const double fl = __ddiv_rn((double)e, 25.0);   // e = max(pitch_count - 75, 0)
    S.fat[e] = make_double2(__dsub_rn(1.0, __dmul_rn(0.02, fl)), __dadd_rn(1.0, __dmul_rn(0.03, fl)));
  }
#endif
#if REPLAY_RATE_TABLES != 2
  for (int i = threadIdx.x; i < KB_ENTRIES; i += blockDim.x) {
    const uint32_t slot = (uint32_t)i % MLBSIM_MAX_LINEUP, g = (uint32_t)i / MLBSIM_MAX_LINEUP;
    S.kb[i] = effective_rates(S.P, g / (MLBSIM_MAX_PITCHERS * MLBSIM_N_TIERS), (g / MLBSIM_N_TIERS) % MLBSIM_MAX_PITCHERS, g % MLBSIM_N_TIERS, slot, 0);
  }
#endif
  __syncthreads();
  const mlbsim_params& P = S.P;
  const int lane = threadIdx.x & 31;
  const uint32_t n_games = (uint32_t)args.n_games;
  constexpr bool noise = NOISE;
  const uint32_t L0 = (uint32_t)P.lineup_len[0], L1 = (uint32_t)P.lineup_len[1];
  const uint32_t nb0 = (uint32_t)P.bullpen_len[0], nb1 = (uint32_t)P.bullpen_len[1];

  uint32_t chunk_next = 0, chunk_end = 0;
  bool exhausted = false;
#if REPLAY_OFFSET_CACHE
  // offsets window: lane l holds the tape bounds of game oc_base + l
  uint32_t oc_base = 0, oc_len = 0;
  bool oc_valid = false;
  uint64_t oc_lo = 0;
#endif
#if MLBSIM_REPLAY_RING
  extern __shared__ __align__(256) unsigned char replay_ring_smem[];
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(replay_ring_smem) + threadIdx.x * RING_BYTES;
  // Ring coordinates: `tp` is moved back to the 32-byte granule holding the game's first draw and `pos` / `len` count
  // from there (the 0..3 entries of shift ride in the top bits of `status` and come off again when the game is written).
  uint32_t req_g = 0, end_g = 0;   // next granule to ask for; granules of this game's tape
#if REPLAY_BOUNDARY_SKIP_WAIT
  uint32_t landed_g = 0;           // granules below this one were asked for before the lane's latest fill: they have landed
#endif
#endif

  // lane state
  const double* tp = nullptr;
  uint32_t pos = 0, len = 0, game = 0;
  uint32_t cur = 0, oth = 0, sc = 0, hidx = 0, hs = HS_DEAD;
  uint32_t nused = 0;                          // picks so far: home staff [7:0], away staff [15:8]
  uint32_t status = 0;
  unsigned long long pa_cur = 0, pa_oth = 0;   // 9-bit outcome counters of the side at bat / the other side

  for (;;) {
    // ---------------- refill / termination ----------------
    if (__any_sync(0xffffffffu, hs == HS_DEAD)) {
      const uint32_t need = __ballot_sync(0xffffffffu, hs == HS_DEAD);
      if (chunk_next >= chunk_end && !exhausted) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counter, (unsigned long long)chunk);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= (unsigned long long)n_games) {
          exhausted = true;
        } else {
          chunk_next = (uint32_t)base;
          const unsigned long long e = base + chunk;
          chunk_end = e < (unsigned long long)n_games ? (uint32_t)e : n_games;
        }
      }
      const uint32_t avail = chunk_end - chunk_next;
      const uint32_t rank = __popc(need & lanemask_lt());
#if REPLAY_OFFSET_CACHE
      // games chunk_next .. chunk_next + min(want, avail) - 1 are about to be handed out: slide the window when they do not all
      // lie inside it (once per <= 32 games; warp-uniform)
      {
        const uint32_t want0 = __popc(need);
        const uint32_t take = want0 < avail ? want0 : avail;
        if (take != 0u && (!oc_valid || chunk_next < oc_base || chunk_next + take > oc_base + 32u)) {
          oc_base = chunk_next;
          oc_valid = true;
          const uint32_t g0 = oc_base + (uint32_t)lane, g = g0 < n_games ? g0 : n_games;
          const uint32_t g1 = g < n_games ? g + 1u : n_games;
          oc_lo = args.offsets[g];
          oc_len = (uint32_t)(args.offsets[g1] - oc_lo);
          if (g0 < chunk_end) {   // this game will be played by this warp: its tape's first lines into L2 now
            const double* t0 = args.tape + (oc_lo & ~(uint64_t)15);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(t0));
            if (oc_len > 16u) asm volatile("prefetch.global.L2 [%0];" ::"l"(t0 + 16));
          }
        }
      }
      const uint32_t oc_idx = (hs == HS_DEAD && rank < avail) ? chunk_next + rank - oc_base : 0u;
      const uint64_t sh_lo = __shfl_sync(0xffffffffu, oc_lo, (int)oc_idx);
      const uint32_t sh_len = __shfl_sync(0xffffffffu, oc_len, (int)oc_idx);
#endif
      if (hs == HS_DEAD && rank < avail) {
        game = chunk_next + rank;
#if REPLAY_OFFSET_CACHE
        const uint64_t lo = sh_lo, hi = sh_lo + sh_len;
#else
        const uint64_t lo = args.offsets[game], hi = args.offsets[game + 1];
#endif
#if MLBSIM_REPLAY_RING
        const uint32_t shift = (uint32_t)lo & 3u;
        tp = args.tape + (lo - shift); len = (uint32_t)(hi - lo) + shift; pos = shift;
        status = shift << 30;
        end_g = (len + 3u) >> 2;
        // the first window (at most entries 3 .. 14 of the aligned tape) needs granules 0 .. 3 before anything else
#pragma unroll
        for (uint32_t g = 0; g < 4u; ++g)
          if (g < end_g) ring_fetch(ring, tp, g);
        req_g = 4u;
#if REPLAY_BOUNDARY_SKIP_WAIT
        landed_g = 0u;
#endif
#if REPLAY_FILL_FIRST
        cp_async_commit();   // the first window's granules are a group of their own: the first step's wait must cover them
#endif
        cur = 0; oth = 0; sc = 0; hidx = 0; hs = HS_NEW_HALF; nused = 0; pa_cur = 0; pa_oth = 0;
#else
        tp = args.tape + lo; len = (uint32_t)(hi - lo); pos = 0;
        cur = 0; oth = 0; sc = 0; hidx = 0; hs = HS_NEW_HALF; nused = 0; status = 0; pa_cur = 0; pa_oth = 0;
#endif
      }
      const uint32_t want = __popc(need);
      chunk_next += want < avail ? want : avail;
      if (exhausted && !__any_sync(0xffffffffu, hs != HS_DEAD)) break;
    }

#if REPLAY_STEP_LOOP
#pragma unroll 1
#else
#pragma unroll
#endif
    for (int rep = 0; rep < REPLAY_UNROLL; ++rep) {
    if ((hs & (HS_OVER | HS_DEAD)) == 0u) {
      // ---------------- one plate appearance (core/pa_simulator.py:91-143) ----------------
      const uint32_t side = hidx & 1u;
      const uint32_t slot = cur & CX_SLOT;
      const uint32_t tier = (cur >> 4) & 3u;
      const uint32_t pit = (cur >> CX_PIT_SHIFT) & 7u;
      const int pc = (int)(cur >> CX_PC_SHIFT);
#if REPLAY_RATE_TABLES >= 2
      double2 rates;
#if REPLAY_RATE_TABLES == 3
      if (pc <= 75) {
        rates = S.kb[((side * MLBSIM_MAX_PITCHERS + pit) * MLBSIM_N_TIERS + tier) * MLBSIM_MAX_LINEUP + slot];
      } else
#endif
      if (REGULAR || pc < 75 + FAT_ENTRIES) {
        const double2 b = S.base[(side * MLBSIM_MAX_PITCHERS + pit) * MLBSIM_N_TIERS + tier];
        const int fe = pc > 75 ? pc - 75 : 0;
#ifdef MLBSIM_DEBUG_BOUNDS
        if (REGULAR && fe >= FAT_ENTRIES) __trap();   // the dispatch rule promised pc <= 120
#endif
        const double2 f = S.fat[REGULAR && fe >= FAT_ENTRIES ? FAT_ENTRIES - 1 : fe];   // (REGULAR: pc <= 120 by construction; the clamp only keeps the read inside the table)
        rates.x = __dmul_rn(__ddiv_rn(__dadd_rn(P.bat_k[side][slot], __dmul_rn(b.x, f.x)), 2.0), P.k_mod);
        rates.y = __dmul_rn(__ddiv_rn(__dadd_rn(P.bat_bb[side][slot], __dmul_rn(b.y, f.y)), 2.0), P.bb_mod);
      } else {
        rates = effective_rates(P, side, pit, tier, slot, pc);
      }
#else
      const double2 rates = pc <= 75 ? S.kb[((side * MLBSIM_MAX_PITCHERS + pit) * MLBSIM_N_TIERS + tier) * MLBSIM_MAX_LINEUP + slot]
                                     : effective_rates(P, side, pit, tier, slot, pc);
#endif
      const double k = rates.x, bb = rates.y;
      // beta_noise (core/pa_simulator.py:23-32): a draw only when 0 < p < 1
      const bool has_k = REGULAR || (noise && k > 0.0 && k < 1.0), has_b = REGULAR || (noise && bb > 0.0 && bb < 1.0);
      const uint32_t ob = hs & HS_OB;
      const uint32_t outs = ob >> 3;
      const bool on1 = (ob & 1u) != 0, on2 = (ob & 2u) != 0;
      uint32_t o;
      double h1, h2, h3;
      bool is_out, is_1b, held;
#if MLBSIM_REPLAY_RING
      {
        // Fill rule.  Everything asked for before this step has landed after the wait; then ask for the granules up to
        // 256 bytes past the read position.  A step (plus a boundary pass) moves the position by at most 14 entries =
        // 4 granules and its window ends at most 3 granules past the position's own, so asking up to position + 8 at
        // every step keeps the next step's window inside what was asked for one step earlier.
#if !REPLAY_FILL_FIRST
        cp_async_wait_all();
#endif
        const uint32_t want = (pos >> 2) + RING_GRANULES;
        const uint32_t lim = want < end_g ? want : end_g;
#if REPLAY_BOUNDARY_SKIP_WAIT
        landed_g = req_g;
#endif
#if REPLAY_FILL_NESTED
        // up to four granules per step (the first step of a game asks for four, a step after a maximal step and a boundary
        // pass likewise): nested early exits, addresses as immediates off one base
        if (req_g < lim) {
          const uint32_t n = lim - req_g;
#ifdef MLBSIM_DEBUG_BOUNDS
          if (n > 4u) __trap();   // a step plus a boundary pass moves the position by at most 14 entries
#endif
          const double* src = tp + ((size_t)req_g << 2);
          const uint32_t so = req_g << 5;
#if REPLAY_RING_PREFETCH > 0
          if (lim + REPLAY_RING_PREFETCH / 32 < end_g) asm volatile("prefetch.global.L2 [%0];" ::"l"(tp + ((size_t)lim << 2) + REPLAY_RING_PREFETCH / 8));
#endif
          req_g = lim;
          ring_fetch_at(ring, so, src, 0);
          if (n > 1u) {
            ring_fetch_at(ring, so, src, 1);
            if (n > 2u) {
              ring_fetch_at(ring, so, src, 2);
              if (n > 3u) ring_fetch_at(ring, so, src, 3);
            }
          }
        }
#else
#pragma unroll
        for (int k = 0; k < REPLAY_FILL_UNROLL; ++k)
          if (req_g < lim) { ring_fetch(ring, tp, req_g); ++req_g; }
#pragma unroll 1
        while (req_g < lim) { ring_fetch(ring, tp, req_g); ++req_g; }   // (the first step of a game asks for four)
#endif
#if REPLAY_FILL_FIRST
        // this step's window lies in granules asked for at an earlier step (the fill rule): only those must have landed
        cp_async_commit();
        cp_async_wait_all_but_last();
#endif
      }
#endif
      if (!noise || REGULAR || (has_k && has_b)) {
        // ---- window path: every draw of this plate appearance sits at a known slot of tape[pos .. pos + B + 8] ----
        constexpr int B = NOISE ? 2 : 0;
#if MLBSIM_REPLAY_RING
        // entry by entry out of the ring: the eight leading entries sit at fixed offsets (independent loads), the three
        // base-running draws at offsets computed below
        if (pos + (uint32_t)(B + 9) > len) ring_patch_tail<B + 9>(ring, pos, len);
        const double kp = NOISE ? ring_at(ring, pos) : k;
        const double bp = __dmul_rn(NOISE ? ring_at(ring, pos + 1u) : bb, RC_101);
        const double u = ring_at(ring, pos + B), uhr = ring_at(ring, pos + B + 1);   // u_hr is always consumed (:126)
#else
        double w[B + 9];
        peek_window<B + 9>(tp, pos, len, w);
        const double kp = NOISE ? w[0] : k;
        const double bp = __dmul_rn(NOISE ? w[1] : bb, 1.01);
        const double u = w[B], uhr = w[B + 1];                          // u_hr is always consumed (:126)
#endif
#if REPLAY_STRAIGHT
        // every compare evaluated, outcomes combined with bit operations: straight-line code instead of a branch (and a
        // reconvergence point) per alternative; the split points come from the constant bank
        const bool is_k = u < kp;
        const bool lt_kb = u < __dadd_rn(kp, bp);
        const bool is_bb = !is_k & lt_kb;
        const bool contact = !is_k & !lt_kb;
        const bool lt_hr = uhr < P.pit_hr[side][pit];
        const bool is_hr = contact & lt_hr;
        const bool bip = contact & !lt_hr;
        // resolve_contact (:51-88)
        const double ubip = ring_at(ring, pos + B + 2), uhit = ring_at(ring, pos + B + 3), u7 = ring_at(ring, pos + B + 4), u8 = ring_at(ring, pos + B + 5);
        uint32_t type = (uint32_t)(args.cdf_bip[0] <= ubip) + (uint32_t)(args.cdf_bip[1] <= ubip) + (uint32_t)(args.cdf_bip[2] <= ubip) +
                        (uint32_t)(args.cdf_bip[3] <= ubip);
        type = type > 3u ? 3u : type;
        const double hp = P.hit_prob[side][pit][slot][type];
        const bool hit = bip & (uhit < hp);                                   // stdlib stream (core/bip_resolution.py:62)
        const bool fb = bip & !hit & (u7 < RC_010);                           // u7: 1B/2B/3B split of a hit, else the 10 % fallback draw
        const uint32_t o_hit = MLBSIM_1B + (uint32_t)(args.cdf_hit[0] <= u7) + (uint32_t)(args.cdf_hit[1] <= u7) + (uint32_t)(args.cdf_hit[2] <= u7);
        const uint32_t o_fb = MLBSIM_1B + (uint32_t)(args.cdf_fb[0] <= u8) + (uint32_t)(args.cdf_fb[1] <= u8);
        o = hit ? o_hit : (uint32_t)MLBSIM_OUT;
        o = fb ? o_fb : o;
        o = is_hr ? (uint32_t)MLBSIM_HR : o;
        o = is_bb ? (uint32_t)MLBSIM_BB : o;
        o = is_k ? (uint32_t)MLBSIM_K : o;
#else
        const bool is_k = u < kp;
        const bool is_bb = !is_k && (u < __dadd_rn(kp, bp));
        const bool contact = !is_k && !is_bb;
        const bool is_hr = contact && (uhr < P.pit_hr[side][pit]);
        const bool bip = contact && !is_hr;
        // resolve_contact (:51-88)
#if MLBSIM_REPLAY_RING
        const double ubip = ring_at(ring, pos + B + 2), uhit = ring_at(ring, pos + B + 3), u7 = ring_at(ring, pos + B + 4), u8 = ring_at(ring, pos + B + 5);
#else
        const double ubip = w[B + 2], uhit = w[B + 3], u7 = w[B + 4], u8 = w[B + 5];
#endif
        uint32_t type = (uint32_t)(P.cdf_bip[0] <= ubip) + (uint32_t)(P.cdf_bip[1] <= ubip) + (uint32_t)(P.cdf_bip[2] <= ubip) +
                        (uint32_t)(P.cdf_bip[3] <= ubip);
        if (type > 3u) type = 3u;
        const bool hit = bip && (uhit < P.hit_prob[side][pit][slot][type]);   // stdlib stream (core/bip_resolution.py:62)
        const bool fb = bip && !hit && (u7 < 0.10);                           // u7: 1B/2B/3B split of a hit, else the 10 % fallback draw
        o = MLBSIM_OUT;
        if (hit) o = MLBSIM_1B + (uint32_t)(P.cdf_hit[0] <= u7) + (uint32_t)(P.cdf_hit[1] <= u7) + (uint32_t)(P.cdf_hit[2] <= u7);
        if (fb) o = MLBSIM_1B + (uint32_t)(P.cdf_fb[0] <= u8) + (uint32_t)(P.cdf_fb[1] <= u8);
        if (is_hr) o = MLBSIM_HR;
        if (is_bb) o = MLBSIM_BB;
        if (is_k) o = MLBSIM_K;
#endif
        // ---- handler draws (core/half_inning_simulator.py:73-136, :28-39): they follow the B + 2 (+3 if in play, +1 if
        //      the fallback drew) entries above ----
        is_out = (o == MLBSIM_K) | (o == MLBSIM_OUT);
        is_1b = o == MLBSIM_1B;
        const bool is_2b = o == MLBSIM_2B;
        const bool need1 = (is_out & on1 & (outs < 2u)) | (is_1b & on2) | (is_2b & on1);       // double play | runner on 2B on a single | runner on 1B on a double
        const bool need2 = is_1b & on1;                                                          // runner on 1B on a single
#if MLBSIM_REPLAY_RING
        const uint32_t i1 = pos + (uint32_t)B + 2u + (bip ? 3u + (uint32_t)fb : 0u);
        const uint32_t i2 = i1 + (uint32_t)need1, i3 = i2 + (uint32_t)need2;
        h1 = ring_at(ring, i1);
        h2 = ring_at(ring, i2);   // (read whether or not it is a draw of this plate appearance: the transition ignores what it does not need)
        h3 = ring_at(ring, i3);
        held = is_1b & on2 & !(h1 < RC_040);
        const bool need3 = held & (outs == 2u);                                                  // two outs, runner from 2B held at 3B
        pos = i3 + (uint32_t)need3;
#else
        h1 = bip ? (fb ? w[B + 6] : w[B + 5]) : w[B + 2];
        h2 = fb ? (need1 ? w[B + 7] : w[B + 6]) : (need1 ? w[B + 6] : w[B + 5]);                 // (a single is always a ball in play)
        held = is_1b && on2 && !(h1 < 0.4);
        const bool need3 = held && outs == 2u;                                                   // two outs, runner from 2B held at 3B
        h3 = need2 ? (fb ? w[B + 8] : w[B + 7]) : (fb ? w[B + 7] : w[B + 6]);                    // (held => need1)
        pos += (uint32_t)B + 2u + (bip ? 3u + (uint32_t)fb : 0u) + (uint32_t)need1 + (uint32_t)need2 + (uint32_t)need3;
#endif
      } else {
        // ---- degenerate rate (k or bb outside (0, 1)): draw by draw ----
#if MLBSIM_REPLAY_RING
#define TAKE(want) ring_take(ring, pos, len, (want))
#else
#define TAKE(want) take(tp, pos, len, (want))
#endif
        const double vk = TAKE(has_k);
        const double vb = TAKE(has_b);
        const double kp = k <= 0.0 ? 0.0 : k >= 1.0 ? 1.0 : vk;
        double bp = bb <= 0.0 ? 0.0 : bb >= 1.0 ? 1.0 : vb;
        bp = __dmul_rn(bp, 1.01);
        const double u = TAKE(true);
        const double uhr = TAKE(true);
        const bool is_k = u < kp;
        const bool is_bb = !is_k && (u < __dadd_rn(kp, bp));
        const bool contact = !is_k && !is_bb;
        const bool is_hr = contact && (uhr < P.pit_hr[side][pit]);
        const bool bip = contact && !is_hr;
        const double ubip = TAKE(bip);
        uint32_t type = (uint32_t)(P.cdf_bip[0] <= ubip) + (uint32_t)(P.cdf_bip[1] <= ubip) + (uint32_t)(P.cdf_bip[2] <= ubip) +
                        (uint32_t)(P.cdf_bip[3] <= ubip);
        if (type > 3u) type = 3u;
        const double uhit = TAKE(bip);
        const bool hit = bip && (uhit < P.hit_prob[side][pit][slot][type]);
        const double u7 = TAKE(bip);
        const bool fb = bip && !hit && (u7 < 0.10);
        const double u8 = TAKE(fb);
        o = MLBSIM_OUT;
        if (hit) o = MLBSIM_1B + (uint32_t)(P.cdf_hit[0] <= u7) + (uint32_t)(P.cdf_hit[1] <= u7) + (uint32_t)(P.cdf_hit[2] <= u7);
        if (fb) o = MLBSIM_1B + (uint32_t)(P.cdf_fb[0] <= u8) + (uint32_t)(P.cdf_fb[1] <= u8);
        if (is_hr) o = MLBSIM_HR;
        if (is_bb) o = MLBSIM_BB;
        if (is_k) o = MLBSIM_K;
        is_out = (o == MLBSIM_K) || (o == MLBSIM_OUT);
        is_1b = o == MLBSIM_1B;
        const bool is_2b = o == MLBSIM_2B;
        h1 = TAKE((is_out && on1 && outs < 2u) || (is_1b && on2) || (is_2b && on1));
        h2 = TAKE(is_1b && on1);
        held = is_1b && on2 && !(h1 < 0.4);
        h3 = TAKE(held && outs == 2u);
      }
#undef TAKE
#if REPLAY_STRAIGHT
      const bool lt14 = h1 < RC_014, lt40 = h1 < RC_040, lt10 = h3 < RC_010;
      const bool dA = is_out ? lt14 : (lt40 | (held & (outs == 2u) & lt10));   // (held implies a single)
#else
      bool dA;
      if (is_out) dA = h1 < 0.14;
      else if (is_1b) dA = (h1 < 0.4) || (held && outs == 2u && h3 < 0.10);
      else dA = h1 < 0.4;
#endif
      const bool dB = h2 < RC_080;
      const uint32_t ent = S.lut[(ob << 5) + (o << 2) + (dA ? 2u : 0u) + (dB ? 1u : 0u)];
      hs = (hs & ~HS_OB) + HS_PA_ONE + ent;
      pa_cur += 1ull << (9u * o);

      cur += CX_NEXT_PA;
      const uint32_t L = side ? L1 : L0;
      const bool wrapped = (cur & CX_SLOT) == L;
      if (wrapped) cur -= L;
      if ((hs & HS_OVER) == 0u) {
        if (wrapped && (cur & CX_TIER_MASK) != CX_TIER_MASK) cur += CX_TIER_ONE;   // :177-178
      }
    }
#if REPLAY_INLINE_BOUNDARY && MLBSIM_REPLAY_RING
    // ---------------- end of half inning, common case, inside the round ----------------
    // Same arithmetic as the round-end pass below, evaluated into temporaries and committed only when nothing rare happens;
    // otherwise the lane is parked untouched and the round-end pass does the whole boundary.
    if (rep + 1 < REPLAY_UNROLL && (hs & HS_OVER) != 0u && (hs & HS_PARKED) == 0u) {
      const uint32_t side = hidx & 1u;
      uint32_t runs = (hs >> 16) & 63u;
      const bool reached = (hs & HS_REACHED) != 0u;
      // the (at most two) entries read here landed at this step's own wait, and none of them lies past the end of the tape
      bool fast = (((pos + 2u) >> 2) < landed_g) & (pos + 3u <= len);
      const double b0 = ring_at(ring, pos), b1 = ring_at(ring, pos + 1u);
      const bool need_m = reached & (runs == 0u);
      const double ug = need_m ? b1 : b0;
      const uint32_t npos = pos + (uint32_t)need_m + (uint32_t)reached;
      runs += (uint32_t)(need_m & (b0 < 0.011)) + (uint32_t)(reached & (ug < 0.01));
      const uint32_t nsc = sc + (runs << (side * 16u));
      const uint32_t inning = (hidx >> 1) + 1u;
      const uint32_t a = nsc & 0xFFFFu, h = nsc >> 16;
      const bool over = ((side != 0u) & (hidx >= 17u) & (a != h)) | ((hidx == 16u) & (h > a));
      const uint32_t nb = side ? nb0 : nb1;     // staff that takes the mound next: its context is `oth`
      const bool hook = (nb != 0u) & (((oth & CX_TIER_GE2) != 0u) | (oth >= ((uint32_t)91 << CX_PC_SHIFT)));
      fast = fast & !over & !hook & (inning <= (uint32_t)MLBSIM_REPLAY_MAX_INNINGS);
      if (fast) {
        pos = npos;
        sc = nsc;
        if (runs) {
          mlbsim_replay_game* const R = args.out + game;
          (side ? R->home_runs : R->away_runs)[inning - 1] = (uint8_t)(runs > 255u ? 255u : runs);
        }
        hs = HS_NEW_HALF;
        ++hidx;
        { const uint32_t t = cur; cur = oth; oth = t; }
        { const unsigned long long t = pa_cur; pa_cur = pa_oth; pa_oth = t; }
      } else {
        hs |= HS_PARKED;
      }
    }
#endif
    }

    if ((hs & HS_OVER) != 0u) {
      // ---------------- end of half inning (:242-250) ----------------
      const uint32_t side = hidx & 1u;
      uint32_t runs = (hs >> 16) & 63u;
      const bool reached = (hs & HS_REACHED) != 0u;
      // misc run, ghost run, reliever pick: at most three consecutive entries, loaded together
#if MLBSIM_REPLAY_RING
#if REPLAY_BOUNDARY_SKIP_WAIT
      if (((pos + 2u) >> 2) >= landed_g) cp_async_wait_all();   // (rare: the last step moved the position by almost four granules)
#else
      cp_async_wait_all();   // (the granules of these entries were asked for at the last step's start at the latest)
#endif
      if (pos + 3u > len) ring_patch_tail<3>(ring, pos, len);
      const double b0 = ring_at(ring, pos), b1 = ring_at(ring, pos + 1u), b2 = ring_at(ring, pos + 2u);
#else
      double bw[3];
      peek_window<3>(tp, pos, len, bw);
      const double b0 = bw[0], b1 = bw[1], b2 = bw[2];
#endif
      const bool need_m = reached && runs == 0u;
      const double um = b0, ug = need_m ? b1 : b0;
      const double upick = reached ? (need_m ? b2 : b1) : b0;
      pos += (uint32_t)need_m + (uint32_t)reached;
      if (need_m && um < 0.011) runs += 1u;
      if (reached && ug < 0.01) runs += 1u;
      sc += runs << (side * 16u);
      const uint32_t inning = (hidx >> 1) + 1u;
      mlbsim_replay_game* const R = args.out + game;
      if (runs) {
        if (inning <= MLBSIM_REPLAY_MAX_INNINGS) {
          uint8_t* dst = side ? R->home_runs : R->away_runs;
          dst[inning - 1] = (uint8_t)(runs > 255u ? 255u : runs);
        }
      }
      if (inning > MLBSIM_REPLAY_MAX_INNINGS) status |= MLBSIM_ST_INNINGS_OVERFLOW;
      const uint32_t a = sc & 0xFFFFu, h = sc >> 16;
      bool over = (side != 0u && hidx >= 17u && a != h) || (hidx == 16u && h > a);
      // a tape that ran dry cannot end the game by itself (0.5 draws): stop after a bound (as the oracle does)
      if (side != 0u && pos > len && inning >= 200u) over = true;
      if (!over) {
        hs = HS_NEW_HALF;
        ++hidx;
        { const uint32_t t = cur; cur = oth; oth = t; }
        { const unsigned long long t = pa_cur; pa_cur = pa_oth; pa_oth = t; }
        const uint32_t nb = side ? nb0 : nb1;     // staff that now takes the mound
        if (nb != 0u && ((cur & CX_TIER_GE2) != 0u || cur >= ((uint32_t)91 << CX_PC_SHIFT))) {   // game_simulator.py:8-12
          int idx = (int)upick;
          ++pos;
          idx = idx < 0 ? 0 : (idx >= (int)nb ? (int)nb - 1 : idx);
          cur = (cur & CX_SLOT) | ((uint32_t)(idx + 1) << CX_PIT_SHIFT);
          const uint32_t sh = side ? 0u : 8u;     // new side 0 -> home staff
          const uint32_t cnt = (nused >> sh) & 255u;
          if (cnt < MLBSIM_REPLAY_MAX_USED) (side ? R->used_home : R->used_away)[cnt] = (uint8_t)idx;
          else status |= MLBSIM_ST_USED_OVERFLOW;
          nused += 1u << sh;
        }
      } else {
        if (pos > len) status |= MLBSIM_ST_TAPE_UNDERRUN;
        // side at bat when the game ended: pa_cur belongs to it
        const unsigned long long pa0 = side ? pa_oth : pa_cur, pa1 = side ? pa_cur : pa_oth;
#if MLBSIM_REPLAY_RING
        reinterpret_cast<int4*>(R)[0] = make_int4((int)h, (int)a, (int)inning, (int)(pos - (status >> 30)));
        reinterpret_cast<int4*>(R)[1] = make_int4((int)(nused & 255u), (int)((nused >> 8) & 255u), (int)(status & 0x3FFFFFFFu), 0);
#else
        reinterpret_cast<int4*>(R)[0] = make_int4((int)h, (int)a, (int)inning, (int)pos);
        reinterpret_cast<int4*>(R)[1] = make_int4((int)(nused & 255u), (int)((nused >> 8) & 255u), (int)status, 0);
#endif
#pragma unroll
        for (int o = 0; o < MLBSIM_N_OUTCOMES; ++o) {
          const uint16_t c0 = (uint16_t)((pa0 >> (9 * o)) & 511ull), c1 = (uint16_t)((pa1 >> (9 * o)) & 511ull);
          if (c0) R->pa[0][o] = c0;
          if (c1) R->pa[1][o] = c1;
        }
        hs = HS_DEAD;
      }
    }
    __syncwarp();
  }
}

}  // namespace

extern "C" __global__ void __launch_bounds__(REPLAY_FLAT_THREADS, REPLAY_FLAT_BLOCKS)
mlbsim_replay_flat_kernel(const ReplayArgs args, unsigned long long* counter, uint32_t chunk) {
  replay_body<true, false>(args, counter, chunk);     // simulate_game(use_noise=True), any rates
}

extern "C" __global__ void __launch_bounds__(REPLAY_FLAT_THREADS, REPLAY_FLAT_BLOCKS)
mlbsim_replay_flat_regular_kernel(const ReplayArgs args, unsigned long long* counter, uint32_t chunk) {
  replay_body<true, true>(args, counter, chunk);      // use_noise=True, bullpens on both sides, all rates inside (0, 1)
}

extern "C" __global__ void __launch_bounds__(REPLAY_FLAT_THREADS, REPLAY_FLAT_BLOCKS)
mlbsim_replay_flat_nonoise_kernel(const ReplayArgs args, unsigned long long* counter, uint32_t chunk) {
  replay_body<false, false>(args, counter, chunk);
}
