// native_kernel.cu — native-RNG Monte Carlo game sampler for sm_100a.
//
// Replaces the serial loop of cli/run_distribution_simulator.py:368-382 over
// core/game_simulator.py:14-156 -> core/half_inning_simulator.py:138-258 -> core/pa_simulator.py:91-143
// (+ bip_resolution, fatigue_modeling, bullpen_utils.simulate_reliever_chain).
//
// Execution model
//   * one thread = one game at a time, a register-resident state machine.  The warp runs in
//     ROUNDS: OPT_UNROLL plate-appearance steps (every live lane whose half inning is still
//     running plays one PA per step: straight-line code, ~31 of 32 lanes active), then ONE
//     boundary pass for all lanes whose half ended during the round (misc/ghost run, score,
//     first-1/3/5/7 commits, game-over test, side swap, pitcher hook).  Batching the boundaries
//     costs a lane at most OPT_UNROLL-1 idle steps per half inning but runs the boundary code
//     once per round with ~4x more active lanes (measured +11 % over per-PA boundaries);
//   * a lane that finishes a game immediately starts the next unassigned sim of its warp's
//     chunk (warp-ballot prefix, no atomics); warps pull chunks of the matchup's sim range
//     from a global counter.  Sim i draws Philox2x32-10 blocks with counter (i, draw number,
//     kind, matchup) and a key derived from the seed (philox.cuh), so which lane / block / GPU
//     runs it is irrelevant to its result;
//   * the per-matchup alias table (32 KB, rows at [side][pitcher 8][tier][16 slots], one word read per PA; it
//     arrives as ONE cp.async.bulk copy of the image mlbsim_stage_tables_kernel laid out at upload time, completion on
//     an mbarrier), the transition LUT (4.6 KB), the lineup-wrap table (2.5 KB) and the five joint (away, home)
//     histograms (20 KB of u32 bins) live in shared memory; PA outcomes are counted per warp (ATOMS.POPC.INC merges
//     the lanes that hit the same counter); everything is flushed to the uint64 device totals once per (block,
//     matchup).  The state words are laid out so that table addresses are bit fields of the state and a plate
//     appearance is a handful of adds (see the "side context" and "half-inning state" comments below); a lineup wrap
//     costs nothing inside a round (repeated table rows) and one table lookup at its end;
//   * no tensor cores: there is no dense contraction in this path.
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/mlbsim.h"
#include "game_rules.cuh"
#include "native_kernel.h"
#include "philox.cuh"

// tuning switches (scripts/build_variants.py builds one library per set of -D switches, scripts/variant_bench.sh benches them; results under profiles/)
#ifndef OPT_UNROLL
#define OPT_UNROLL 5   /* plate appearances per round: one refill check and one boundary pass per round (3/4/5/6 measured: profiles/r02_v15_variants.log) */
#endif
#ifndef OPT_PA_WARP
#define OPT_PA_WARP 1  /* 1: PA outcome counters per warp (16 words, same-address lanes merged by ATOMS.POPC.INC) instead of per thread:
                          same speed, 40 KB less shared memory per block (profiles/r02_v15_variants.log) */
#endif
#ifndef OPT_GATE_LAST
#define OPT_GATE_LAST 0  /* G > 0: skip the last PA step of a round when fewer than G lanes are still inside their half */
#endif
#ifndef OPT_REFILL_EVERY
#define OPT_REFILL_EVERY 1  /* look for finished lanes every k-th round only */
#endif
#ifndef OPT_WRAP_ROUND
#define OPT_WRAP_ROUND 1  /* 1: lineup wrap fixed up once per round (repeated table rows), 0: after every plate appearance */
#endif
#ifndef OPT_WRAP_LUT
#define OPT_WRAP_LUT 1  /* 1: the once-per-round lineup-wrap fix-up is one shared-memory lookup instead of ~15 ALU instructions (+2.5 %, profiles/r02_v19_variants.log) */
#endif
#ifndef OPT_PIPELINE_PHILOX
#define OPT_PIPELINE_PHILOX 1  /* 1: the Philox block of a lane's next plate appearance is drawn one step ahead, inside the current step's basic block (+1.2 %, profiles/r02_v22_variants.log) */
#endif
#ifndef OPT_REFILL_INLINE
#define OPT_REFILL_INLINE 0  /* 1: a lane whose game ends takes its next sim right there in the boundary pass (one shared-memory fetch-and-add
                                on the warp's chunk cursor); the warp-wide ballot refill at the top of a round then runs only when the chunk is used
                                up.  Bit-exact, but -0.7 %: the returning ATOMS inside the divergent game-over branch costs more than the ballot
                                refill it saves (profiles/r02_v25_variants.log) */
#endif
#ifndef OPT_ALIAS_OR
#define OPT_ALIAS_OR 0   /* 1: the staged alias table sits at a 32 KB-aligned shared-memory address (in dynamic shared memory, behind the histograms),
                            so the row address is (ctx & row bits) | base — one LOP3 — instead of a mask and an add: 43 instead of 44 instructions per step.
                            Bit-exact; 3.195 vs 3.187e9 games/s (+0.1 ... +0.4 %, inside the run-to-run spread; profiles/r02_v30_variants.log): off */
#endif
#ifndef OPT_TMA
#define OPT_TMA 1      /* 1: tables reach shared memory as cp.async.bulk copies of the pre-laid-out device image (mbarrier completion) */
#endif
#ifndef NATIVE_LB_THREADS
#define NATIVE_LB_THREADS MLBSIM_NATIVE_MAX_THREADS
#endif
#ifndef NATIVE_LB_BLOCKS
#define NATIVE_LB_BLOCKS 2   /* 2 x 640 threads = 40 warps per SM at <= 48 registers, no spills */
#endif

namespace {

// ---- model constants (reference file:line) ----------------------------------------------------
constexpr uint32_t T_DP = 601295421u;        // 0.14 * 2^32   half_inning_simulator.py:76
constexpr uint32_t T_040 = 1717986918u;      // 0.40 * 2^32   :103 (1B: runner on 2B scores), :120 (2B: runner on 1B scores)
constexpr uint32_t T_046 = 1975684956u;      // 0.40 + 0.60*0.10: :103 combined with the two-out rule :28-39
constexpr uint32_t T_080_16 = 52429u;         // 0.80 * 2^16   :105 (1B: runner on 1B takes second; low half of w1)
constexpr uint32_t T_END_BOTH = 472446u;     // 0.011*0.01        misc run AND ghost run  :12-25
constexpr uint32_t T_END_ANY = 89721867u;    // 1-0.989*0.99      misc run OR ghost run
constexpr uint32_t T_END_GHOST = 42949673u;  // 0.01              ghost run only (inning already scored)
constexpr int MAX_PA_PER_HALF = 30;          // half_inning_simulator.py:160
constexpr int PITCH_LIMIT = 90;              // game_simulator.py:8
constexpr int FATIGUE_START = 75;            // fatigue_modeling.py:38
constexpr uint32_t LAST_HIDX = 2u * MLBSIM_MAX_INNINGS - 1u;   // termination guard: bottom half of inning MLBSIM_MAX_INNINGS

// ---- shared-memory layout ---------------------------------------------------------------------
constexpr int SB = 32;                                    // smem joint-histogram bins per team
constexpr int FAT_WORDS = MLBSIM_STAGED_FAT_WORDS;        // 576
// transition LUT: one row per (outs*8 + bases), eight 16-byte groups per row, one per draw variant:
// {sub-draw A threshold at this number of outs, state delta if A fails, state delta if A succeeds, 0} — one LDS.128 per PA.
// Row stride 36 words: consecutive rows start one 16-byte bank group apart, so lanes in different
// (outs, bases) states that look up the same variant do not collide.
constexpr int LUT_ROW_WORDS = 36;
constexpr int LUT_WORDS = 32 * LUT_ROW_WORDS;                                                                  // 1152
constexpr int JOINT_WORDS = MLBSIM_N_CAPS * SB * SB;                                                          // 5120
constexpr int ALIAS_WORDS = MLBSIM_STAGED_ALIAS_WORDS;                                                        // 7680 (30 KB)

struct __align__(16) Smem {
  uint32_t lut[LUT_WORDS];   // first: its rows are addressed by absolute 14-bit shared addresses kept in the state word
#if !OPT_ALIAS_OR
  uint32_t alias[ALIAS_WORDS];   // static, so that the hot loop's LDS carries its address as an immediate
#endif
  uint32_t fthr[FAT_WORDS];  // (fthr and fslope are contiguous: one bulk copy)
  int32_t fslope[FAT_WORDS];
#if OPT_WRAP_LUT
  uint32_t wlut[640];        // lineup-wrap fix-up: byte offset = side * 2048 + half_over * 256 + (slot' | tier << 4) * 4
#endif
  uint32_t innings[MLBSIM_INN_BINS];
  uint32_t rel_games[16];
  uint32_t rel_picks[16];
  uint32_t truncated;
  uint32_t go;  // block-uniform "this matchup still has work"
#if OPT_REFILL_INLINE
  uint32_t wnext[MLBSIM_NATIVE_MAX_THREADS / 32];   // per warp: next unassigned sim of the warp's chunk (offset from sim_lo)
#endif
  unsigned long long mbar;   // completion barrier of the table copies
};
constexpr int PA_FIELDS = 16;  // [side][draw variant]: the two single variants are added up at the flush
// Dynamic shared memory: [five 32x32 joint histograms] [PA outcome counters].  Per-thread counters: pa[16][PA_STRIDE], the
// stride being the compile-time maximum block size, not blockDim.x, so that `variant * stride` is an immediate multiply;
// per-warp counters (OPT_PA_WARP): pa[warp][16].
constexpr uint32_t PA_STRIDE = MLBSIM_NATIVE_MAX_THREADS;
constexpr uint32_t ALIAS_BYTES = (uint32_t)ALIAS_WORDS * 4u;
constexpr uint32_t DYN_JOINT_BYTES = (uint32_t)JOINT_WORDS * 4u;
constexpr uint32_t DYN_PA_WORD0 = DYN_JOINT_BYTES / 4u;
constexpr uint32_t DYN_PA_BYTES = OPT_PA_WARP ? (uint32_t)(MLBSIM_NATIVE_MAX_THREADS / 32) * PA_FIELDS * 4u : (uint32_t)PA_FIELDS * 4u * PA_STRIDE;

}  // namespace

// Statically allocated (42 KB) at global scope so that its PTX symbol keeps this plain name: the
// hot loop takes `mov.u32 r, mlbsim_S` (a link-time constant in the .shared space) as its base and
// every access becomes LDS [Rindex + imm] with no generic->shared conversion in the loop.
// It holds the alias table laid out [side][pitcher 8][tier 4][16 slots][8 words]: ten bits of a side's context word
// ARE the row number, and rows sit at slot + 16 - L so that running off the end of the lineup is a carry out of the
// slot field into the tier field.  Dynamic shared memory: the joint histograms, then the PA outcome counters.
__shared__ Smem mlbsim_S;
extern __shared__ __align__(16) uint32_t mlbsim_S_pa[];

namespace {
#define S mlbsim_S
#define S_pa mlbsim_S_pa

__device__ __forceinline__ uint32_t smem_static_base() {
  uint32_t a;
  asm("mov.u32 %0, mlbsim_S;" : "=r"(a));
  return a;
}
__device__ __forceinline__ uint32_t smem_dynamic_base() {
  uint32_t a;
  asm("mov.u32 %0, mlbsim_S_pa;" : "=r"(a));
  return a;
}

#if OPT_ALIAS_OR
// shared-window address of the staged alias table: the first 32 KB boundary behind the histograms and PA counters
// (rounded down inside an asm statement: a compiler that knows the low bits are zero turns the OR of the lookup back into mask + add)
__device__ __forceinline__ uint32_t alias_base() {
  uint32_t a = smem_dynamic_base() + DYN_JOINT_BYTES + DYN_PA_BYTES + 0x7FFFu;
  asm("and.b32 %0, %0, 0xFFFF8000;" : "+r"(a));
  return a;
}
#else
__device__ __forceinline__ uint32_t alias_base() { return smem_static_base() + (uint32_t)offsetof(Smem, alias); }
#endif

// All hot shared-memory traffic uses 32-bit shared-window addresses (ld/st/red.shared), so the
// loop carries one base register instead of 64-bit generic pointers.
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void red_inc_shared(uint32_t addr) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr)); }
__device__ __forceinline__ uint32_t fetch_inc_shared(uint32_t addr) {
  uint32_t v;
  asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// ---- bulk-copy engine (TMA, 1-D form) + mbarrier ------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(mbar),
      "r"(parity)
      : "memory");
}

// ---- per-side context word (one per batting side; `cur` = side at bat, `oth` = the other) -------
// slot'[8:5] | tier[10:9] | pit[13:11] | side[14] | pitch_count[27:15] | bias[31:28]
//   * ctx & CX_ROW_BYTES IS the byte offset of the alias row in the staged table of both sides (a row is 32 bytes; the
//     side bit never changes): one LOP3, no per-lane base register;
//   * slot' = slot + 16 - L ("biased"; L = lineup length, bias = 16 - L rides in the top nibble): stepping past the last
//     batter carries out of the slot field into the tier field — the TTO increment of half_inning_simulator.py:177-178.
//     The first three rows of every (pitcher, tier) group — padding, since real rows start at `bias` >= 7 — repeat the
//     first three batters of the group the carry comes from (stage kernel below), so that a round of four plate
//     appearances never has to look at the wrap: one fix-up per ROUND puts the slot back to `bias` (WRAP_STEP = false);
//   * pitch_count is the top field below the bias, so "pitch_count > n" is an unsigned compare of
//     (ctx & CX_PC_BITS) against ((n + 1) << CX_PC_SHIFT).  13 bits: MLBSIM_MAX_INNINGS keeps a game below 7 680 PAs.
constexpr int CX_SLOT_SHIFT = 5;
constexpr uint32_t CX_SLOT = 0xFu << CX_SLOT_SHIFT;
constexpr uint32_t CX_TIER_ONE = 1u << 9;
constexpr uint32_t CX_TIER_MASK = 3u << 9;
constexpr uint32_t CX_TIER_GE2 = 2u << 9;
constexpr int CX_T_SHIFT = 9;  // (ctx >> 9) & 31 = pit*4 + tier
constexpr int CX_PIT_SHIFT = 11;
constexpr int CX_SIDE_SHIFT = 14;
constexpr uint32_t CX_SIDE = 1u << CX_SIDE_SHIFT;
constexpr int CX_PC_SHIFT = 15;
constexpr uint32_t CX_NEXT_PA = (1u << CX_PC_SHIFT) + (1u << CX_SLOT_SHIFT);  // one more pitch, next batter
constexpr int CX_BIAS_SHIFT = 28;
constexpr uint32_t CX_PC_BITS = (1u << CX_BIAS_SHIFT) - 1u;   // mask that drops the bias nibble
constexpr uint32_t CX_ROW_BYTES = 0x7FE0u;                    // side | pit | tier | slot', already scaled by the 32-byte row
constexpr int WRAP_DUP_ROWS = OPT_UNROLL - 1;                 // plate appearances that can follow a wrap inside one round
static_assert(WRAP_DUP_ROWS >= 1 && WRAP_DUP_ROWS <= 16 - MLBSIM_MAX_LINEUP, "repeated rows must fit below the smallest bias");

// ---- half-inning state word -------------------------------------------------------------------
// shared-memory address of the LUT row of (outs*8 + bases) [13:2] | three outs [14] | lane has no game [15] |
// runs [20:16] | reached_count [25:21] | n_pa + 2 [31:26]: the count starts at 2, so "no PA yet in this
// half" is `hs < 3 << 26` and the 30th PA carries into bit 31 (the 30-PA cap).  A plate appearance is ONE ADD:
// the LUT entry of (row, variant, sub-draw A) holds new row address - this row's address + runs + reached + three-outs
// flag + one PA, so hs' = hs + entry.
constexpr uint32_t HS_ROW = 0x3FFCu;   // absolute .shared address: the LUT sits in the first 16 KB of the window (checked at kernel start)
constexpr uint32_t HS_3OUT = 1u << 14;
constexpr uint32_t HS_DEAD = 1u << 15;                   // (a dead lane holds exactly this value)
constexpr int HS_RUNS_SHIFT = 16;
constexpr int HS_REACHED_SHIFT = 21;
constexpr uint32_t HS_REACHED = 31u << HS_REACHED_SHIFT;
constexpr int HS_PA_SHIFT = 26;
constexpr uint32_t HS_PA_ONE = 1u << HS_PA_SHIFT;
constexpr uint32_t HS_NEW_HALF = (32u - MAX_PA_PER_HALF) << HS_PA_SHIFT;
constexpr uint32_t HS_FIRST_PA_BELOW = HS_NEW_HALF + HS_PA_ONE;   // hs < this  <=>  no PA played yet in this half
constexpr uint32_t HS_OVER = (1u << 31) | HS_3OUT;      // 30 PAs, or three outs

// threshold of sub-draw A: double play (K/OUT), runner on 2B scores on a single (0.46 with two
// outs: 0.4 + 0.6*0.1), runner on 1B scores on a double
__device__ uint32_t subdraw_a_threshold(int o, int two_outs) {
  if (o == MLBSIM_1B) return two_outs ? T_046 : T_040;
  if (o == MLBSIM_2B) return T_040;
  return T_DP;
}

__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// joint-histogram commit: shared u32 bin when both scores fit the 32x32 window, else the global
// u64 bin directly (a team scoring >= 32 runs: ~never, but exact)
__device__ __forceinline__ void commit_joint(uint32_t joint_addr, unsigned long long* H, uint32_t c, uint32_t a, uint32_t h) {
  if ((a | h) < (uint32_t)SB) {
    red_inc_shared(joint_addr + ((c * (SB * SB) + a * SB + h) << 2));
  } else {
    const uint32_t ac = a < MLBSIM_HIST_BINS - 1 ? a : MLBSIM_HIST_BINS - 1;
    const uint32_t hc = h < MLBSIM_HIST_BINS - 1 ? h : MLBSIM_HIST_BINS - 1;
    atomicAdd(&H[((size_t)c * MLBSIM_HIST_BINS + ac) * MLBSIM_HIST_BINS + hc], 1ull);
  }
}

// word offsets inside mlbsim_hist
constexpr size_t H_INN = (size_t)MLBSIM_N_CAPS * MLBSIM_HIST_BINS * MLBSIM_HIST_BINS;
constexpr size_t H_PA = H_INN + MLBSIM_INN_BINS;
constexpr size_t H_RELG = H_PA + 16;
constexpr size_t H_RELP = H_RELG + 16;
constexpr size_t H_N = H_RELP + 16;
constexpr size_t H_TRUNC = H_N + 1;

// One matchup, all the sims this warp can grab.  FATIGUE: some staff has no bullpen (pitch counts
// pass 75); COUNT_PA: keep the PA_RECAP_LOG outcome counters.
// Lineup wrap.  The biased slot runs from 16 - L to 15; one step further the add has carried into the tier field
// (that carry IS the TTO increment: the next PA of this half starts a new pass through the lineup) and the slot field
// counts 0, 1, 2 ... from the first batter again.  This puts the slot back to bias + that count and takes the tier
// increment back when the tier was already 4+ (tier bits zero after the carry: it ran on into the pitcher field, and
// -TIER_ONE borrows it back), or when the wrap was the very last PA of the half: the next half then starts at index 0,
// where the reference's `batter_idx > 0` is false.  Lanes that did not wrap pass through unchanged.
__device__ __forceinline__ uint32_t wrap_fix(uint32_t cur, uint32_t hs) {
  const uint32_t low = cur & CX_SLOT;
  const uint32_t bias = (cur >> (CX_BIAS_SHIFT - CX_SLOT_SHIFT)) & CX_SLOT;   // bias << CX_SLOT_SHIFT
  uint32_t fix = bias;
  if ((cur & CX_TIER_MASK) == 0u || ((hs & HS_OVER) != 0u && low == 0u)) fix -= CX_TIER_ONE;
  return low < bias ? cur + fix : cur;
}

// WRAP_STEP: put the lineup slot back after every plate appearance instead of once per round (needed when a lineup is
// too short for the repeated rows: more than one wrap could fall into one round).
template <bool FATIGUE, bool COUNT_PA, bool WRAP_STEP>
__device__ __forceinline__ void play_matchup(const NativeArgs& args, unsigned long long* const H,
                                             unsigned long long* const counter, const uint32_t rng_matchup, const uint32_t L0,
                                             const uint32_t L1, const uint32_t nb0, const uint32_t nb1) {
  const int lane = threadIdx.x & 31;
  const uint32_t n_sims = args.n_sims;
  const uint32_t sb = smem_static_base();
  const uint32_t A_ALIAS = alias_base();
  const uint32_t init0 = (16u - L0) * ((1u << CX_BIAS_SHIFT) + (1u << CX_SLOT_SHIFT));   // bias nibble + biased slot 0, side 0
  const uint32_t init1 = (16u - L1) * ((1u << CX_BIAS_SHIFT) + (1u << CX_SLOT_SHIFT)) | CX_SIDE;
  const uint32_t A_FTHR = sb + (uint32_t)offsetof(Smem, fthr);
  const uint32_t A_FSLOPE = sb + (uint32_t)offsetof(Smem, fslope);
  const uint32_t A_LUT = sb + (uint32_t)offsetof(Smem, lut);
  const uint32_t hs_new_half = HS_NEW_HALF | A_LUT;   // row of (0 outs, bases empty)
  const uint32_t A_JOINT = smem_dynamic_base();
  const uint32_t A_INN = sb + (uint32_t)offsetof(Smem, innings);
  const uint32_t A_RELG = sb + (uint32_t)offsetof(Smem, rel_games);
  const uint32_t A_RELP = sb + (uint32_t)offsetof(Smem, rel_picks);
  const uint32_t A_TRUNC = sb + (uint32_t)offsetof(Smem, truncated);
#if OPT_WRAP_LUT
  const uint32_t A_WLUT = sb + (uint32_t)offsetof(Smem, wlut);
#endif
#if OPT_PA_WARP
  // per-warp PA counters: pa[warp][side*8 + variant]
  constexpr uint32_t pa_stride = 4u;    // bytes between variants
  constexpr uint32_t pa_side = 32u;     // bytes between sides
  const uint32_t pa_mine = smem_dynamic_base() + DYN_JOINT_BYTES + (threadIdx.x >> 5) * (PA_FIELDS * 4u);
#else
  // per-thread PA counters: pa[f][tid], f = side*8 + variant (column = thread: bank = lane, conflict-free)
  constexpr uint32_t pa_stride = PA_STRIDE * 4u;  // bytes between variants
  constexpr uint32_t pa_side = 8u * pa_stride;    // bytes between sides
  const uint32_t pa_mine = smem_dynamic_base() + DYN_JOINT_BYTES + threadIdx.x * 4u;
#endif

#if OPT_REFILL_INLINE
  const uint32_t A_WNEXT = sb + (uint32_t)offsetof(Smem, wnext) + (threadIdx.x >> 5) * 4u;
  if (lane == 0) sts32(A_WNEXT, 0u);
  __syncwarp();
#endif

  // ---- warp-uniform work state ----
  uint32_t chunk_next = 0, chunk_end = 0;
  bool exhausted = false;
#if OPT_REFILL_EVERY > 1
  uint32_t round_no = 0;
#endif

  // ---- lane state ----
  uint32_t c0 = 0, c1 = 0;   // Philox counter: c0 = sim_lo, c1 = PA number | matchup | sim_hi
  uint32_t cur = 0, oth = 0; // side contexts: cur = side at bat
  uint32_t sc = 0;           // score: side at bat | other side << 16
  uint32_t hidx = 0;         // half-inning index: side = hidx & 1, inning = (hidx >> 1) + 1
  uint32_t hs = HS_DEAD;     // half-inning state; HS_DEAD = this lane has no game
  uint32_t used = 0;         // reliever bitmasks: home staff bits 0-7, away staff 8-15
  uint32_t w1_first = 0;     // w1 of the current half's first PA (end-of-half draw)
  uint32_t pa_addr = pa_mine;  // counters of the side at bat

  for (;;) {
    // ---------------- refill / termination (warp-synchronous) ----------------
#if OPT_REFILL_EVERY > 1
    // a finished lane may sit out up to OPT_REFILL_EVERY - 1 rounds; the bookkeeping below then runs less often.
    // A warp with no live lane at all looks every round (it would spin otherwise).
    const bool look = (round_no++ % OPT_REFILL_EVERY) == 0u || !__any_sync(0xffffffffu, hs != HS_DEAD);
    if (look && __any_sync(0xffffffffu, hs == HS_DEAD)) {
#else
    if (__any_sync(0xffffffffu, hs == HS_DEAD)) {
#endif
      const uint32_t need = __ballot_sync(0xffffffffu, hs == HS_DEAD);
#if OPT_REFILL_INLINE
      // the chunk cursor lives in shared memory (lanes advance it on their own when their game ends; it may have run
      // past the end of the chunk by the number of lanes that found it empty)
      chunk_next = lds32(A_WNEXT);
      if (chunk_next > chunk_end) chunk_next = chunk_end;
      __syncwarp();
#endif
      if (chunk_next >= chunk_end && !exhausted) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(counter, (unsigned long long)args.chunk);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= (unsigned long long)n_sims) {
          exhausted = true;
        } else {
          chunk_next = (uint32_t)base;
          const unsigned long long e = base + args.chunk;
          chunk_end = e < (unsigned long long)n_sims ? (uint32_t)e : n_sims;
        }
      }
      const uint32_t avail = chunk_end - chunk_next;
      const uint32_t rank = __popc(need & lanemask_lt());
      if (hs == HS_DEAD && rank < avail) {
        const unsigned long long sim = (unsigned long long)args.sim_lo + (chunk_next + rank);
        c0 = (uint32_t)sim;
        c1 = ((uint32_t)(sim >> 32) << RNG_SIMHI_SHIFT) | rng_matchup;
        cur = init0; oth = init1;
        sc = 0; hidx = 0; hs = hs_new_half; used = 0;   // (pa_addr already points at the away side's counters)
      }
      const uint32_t want = __popc(need);
      chunk_next += want < avail ? want : avail;
#if OPT_REFILL_INLINE
      if (lane == 0) sts32(A_WNEXT, chunk_next);
      __syncwarp();
#endif
      if (exhausted && !__any_sync(0xffffffffu, hs != HS_DEAD)) break;
    }

    uint32_t w1 = 0;
#if OPT_PIPELINE_PHILOX
    uint32_t nw0 = 0, nw1 = 0;   // the block of this lane's NEXT plate appearance, drawn one step ahead (see below)
#endif
#pragma unroll
    for (int rep = 0; rep < OPT_UNROLL; ++rep) {
#if OPT_GATE_LAST > 0
      if (rep == OPT_UNROLL - 1 && __popc(__ballot_sync(0xffffffffu, (hs & (HS_OVER | HS_DEAD)) == 0u)) < OPT_GATE_LAST) break;
#endif
    if ((hs & (HS_OVER | HS_DEAD)) == 0u) {   // live lane whose half inning is still running
      // ---------------- one plate appearance ----------------
      uint32_t w0;
#if OPT_PIPELINE_PHILOX
      // Software pipeline over the steps of a round: a lane that is live at step s + 1 was live at step s, and its
      // counter there is this one plus one, whatever this plate appearance turns out to be.  So the next block is drawn
      // HERE, in the same basic block as this step's table look-ups: two independent dependency chains per warp (ten
      // multiply-xor rounds against two shared-memory round trips) instead of one after the other.  Counter-based
      // generator: when a block is computed changes nothing about its value.
      if (rep == 0) philox2x32_native(c0, c1, args.keys, nw0, nw1);
      w0 = nw0;
      w1 = nw1;
      ++c1;
      if (rep + 1 < OPT_UNROLL) philox2x32_native(c0, c1, args.keys, nw0, nw1);
#else
      philox2x32_native(c0, c1, args.keys, w0, w1);
      ++c1;
#endif

      // The first PA of a half has empty bases, so its w1 never reaches the transition: keep it
      // as this half's end-of-inning (misc / ghost run) draw.
      // (a half always starts at step 0 of a round: new halves come out of the boundary pass, new games out of the refill)
      if (rep == 0 && hs < HS_FIRST_PA_BELOW) w1_first = w1;

      // Walker alias draw: one shared-memory word per PA.  Entry = (2^29 - 1 - cut29) << 3 | alias; the column is kept iff
      // (w0 << 3) > entry — for exactly cut29 of the column's 2^29 draws: the tie with the entry's alias bits falls to
      // the alias, so an outcome of probability 0 never occurs.
      const uint32_t col = w0 >> 29;
      const uint32_t row = cur & CX_ROW_BYTES;
#ifdef MLBSIM_DEBUG_BOUNDS   // (compute-sanitizer stand-in: scripts/r02_cycle.sh runs the GPU suite on a build with these traps)
      {
        const uint32_t sd = hidx & 1u, Lc = sd ? L1 : L0, nbp = sd ? nb1 : nb0, bias = 16u - Lc;
        const uint32_t slotp = (cur & CX_SLOT) >> CX_SLOT_SHIFT, pit = (cur >> CX_PIT_SHIFT) & 7u;
        if (row + (col << 2) >= ALIAS_BYTES) __trap();
        if ((cur >> CX_BIAS_SHIFT) != bias || ((cur >> CX_SIDE_SHIFT) & 1u) != sd) __trap();
        if (WRAP_STEP ? (slotp < bias || pit > nbp) : (slotp < bias && (slotp >= (uint32_t)WRAP_DUP_ROWS || rep == 0))) __trap();
        if ((hs & HS_ROW) < A_LUT || (hs & HS_ROW) > A_LUT + 23u * LUT_ROW_WORDS * 4u || ((hs & HS_ROW) - A_LUT) % (LUT_ROW_WORDS * 4u)) __trap();
        if ((c1 & 0x1FFFu) == 0u) __trap();   // the PA field of the Philox counter wrapped
      }
#endif
#if OPT_ALIAS_OR
      const uint32_t e = lds32(((cur & CX_ROW_BYTES) | A_ALIAS) + (col << 2));   // (A_ALIAS has no bit in common with the row field)
#else
      const uint32_t e = lds32(A_ALIAS + row + (col << 2));
#endif
      uint32_t o = ((w0 << 3) > e) ? col : (e & 7u);   // draw variant 0..7 (7 = single, runner on first advances)
      if (FATIGUE) {
        if ((cur & CX_PC_BITS) >= ((uint32_t)(FATIGUE_START + 1) << CX_PC_SHIFT)) {   // past 75 pitches (rare)
          uint32_t T = (cur >> CX_T_SHIFT) & 31u;  // pit*4 + tier
          const uint32_t slotp = (cur & CX_SLOT) >> CX_SLOT_SHIFT, bias = cur >> CX_BIAS_SHIFT;
          uint32_t slot = slotp - bias;
          if (!WRAP_STEP && slotp < bias) {
            // after a lineup wrap inside this round (repeated rows): the batter is number slot', and a carry that left
            // tier bits 0 came out of a saturated tier — it is still (pit - 1, tier 4+)
            slot = slotp;
            if ((T & 3u) == 0u) T -= 1u;
          }
          if (T < 4u) {  // the starter
            const uint32_t side = (cur >> CX_SIDE_SHIFT) & 1u;
            const uint32_t ex = ((cur & CX_PC_BITS) >> CX_PC_SHIFT) - FATIGUE_START;
            const uint32_t srow = (((side * MLBSIM_N_TIERS + T) * MLBSIM_MAX_LINEUP + slot) * MLBSIM_ROW_WORDS) << 2;
            const uint32_t u = w0 >> 1;
            o = 0;
#pragma unroll
            for (int j = 0; j < 6; ++j) o += (u >= lds32(A_FTHR + srow + 4 * j) + ex * lds32(A_FSLOPE + srow + 4 * j)) ? 1u : 0u;
            // thresholds know seven outcomes: the single's 0.8 draw comes from the low half of w1 here
            if (o == (uint32_t)MLBSIM_1B && (w1 & 0xFFFFu) < T_080_16) o = MLBSIM_V_1B_ADV;
          }
        }
      }

      // base-running sub-draw A from w1 (full 32-bit compare against the group's threshold); sub-draw B
      // arrived inside the variant.  The entry is a delta of the whole state word.
      const uint4 g = lds128((hs & HS_ROW) + (o << 4));
      hs += w1 < g.x ? g.z : g.y;

      if (COUNT_PA) red_inc_shared(pa_addr + o * pa_stride);

      // next batter, one more pitch (pitch_count counts PAs: half_inning_simulator.py:218-220)
      cur += CX_NEXT_PA;
      if (WRAP_STEP) cur = wrap_fix(cur, hs);
    }
    }
    // Stepping past the last batter carried into the tier field and left the slot in the repeated rows 0..2 of the next
    // group: put it back to `bias` once per round.
    if (!WRAP_STEP) {
#if OPT_WRAP_LUT
      uint32_t off = (cur >> 3) & 0x8FCu;                 // side[14] -> 2048, slot' | tier [10:5] -> words
      if ((hs & HS_OVER) != 0u) off += 256u;
      cur += lds32(A_WLUT + off);
#else
      cur = wrap_fix(cur, hs);
#endif
    }
    if ((hs & HS_OVER) != 0u) {   // (a dead lane has only HS_DEAD set)
      {
        // ---------------- end of half inning ----------------
        // `sc` keeps the score of the side AT BAT in its low half and the other side's in its high half (the halves
        // trade places when the sides do), so adding this half's runs needs no shift by the side.
        const uint32_t side = hidx & 1u;
        uint32_t runs = (hs >> HS_RUNS_SHIFT) & 31u;
        // misc / ghost run (half_inning_simulator.py:12-25): rare (2 %), only if a runner reached.
        // T_END_BOTH < T_END_GHOST < T_END_ANY, so one compare screens all three.
        if (w1_first < T_END_ANY && (hs & HS_REACHED) != 0u) {
          runs += (runs == 0u) ? 1u + (uint32_t)(w1_first < T_END_BOTH) : (uint32_t)(w1_first < T_END_GHOST);
        }
        sc += runs;
        // first 1/3/5/7 innings: committed when the bottom half of that inning ends (hidx = 1, 5, 9, 13): the home team
        // just batted (low half), joint bin [away][home].  Both below 32 (all but never false): shared u32 bin at
        // away * 32 + home = (sc >> 11) + (sc & 31); else the global u64 bin directly.
        if ((hidx & 0xFFFFFFF3u) == 1u) {
          if ((sc & 0xFFE0FFE0u) == 0u) red_inc_shared(A_JOINT + ((hidx & 0xCu) << 10) + (((sc >> 11) + (sc & 31u)) << 2));
          else commit_joint(A_JOINT, H, hidx >> 2, sc >> 16, sc & 0xFFFFu);
        }
        // game_simulator.py:72 (bottom of the 9th skipped when the home team leads) and :110-111; termination guard:
        // the bottom of inning MLBSIM_MAX_INNINGS ends the game whatever the score.
        const uint32_t bat = sc & 0xFFFFu, fld = sc >> 16;   // score of the side that just batted / of the side in the field
        const bool over = ((side != 0u) & (hidx >= 17u) & ((bat != fld) | (hidx >= LAST_HIDX))) | ((hidx == 16u) & (fld > bat));
        if (!over) {
          hs = hs_new_half;
          ++hidx;
          sc = __funnelshift_l(sc, sc, 16);   // the other side bats next
          const uint32_t t = cur; cur = oth; oth = t;
          if (COUNT_PA) pa_addr += side ? 0u - pa_side : pa_side;   // counters of the side that bats next
          // should_replace_pitcher (game_simulator.py:8-12) for the staff that now takes the mound
          const uint32_t nb = side ? nb0 : nb1;  // `side` is the half that just ended
          if (nb != 0u && ((cur & CX_TIER_GE2) != 0u || (cur & CX_PC_BITS) >= ((uint32_t)(PITCH_LIMIT + 1) << CX_PC_SHIFT))) {
            // stationary law of simulate_reliever_chain: uniform, with replacement.  The half ended on
            // an out, whose transition ignores the low half of w1 (only singles read it).
            const uint32_t idx = ((w1 & 0xFFFFu) * nb) >> 16;
            cur = (cur & (CX_SLOT | CX_SIDE | ~CX_PC_BITS)) | ((idx + 1u) << CX_PIT_SHIFT);
            const uint32_t bit = idx + (side ? 0u : 8u);
            // "sims in which reliever i appeared" (run_distribution_simulator.py:391-397): count his first
            // appearance in this game here rather than walking the mask when the game ends
            if (((used >> bit) & 1u) == 0u) red_inc_shared(A_RELG + (bit << 2));
            used |= 1u << bit;
            red_inc_shared(A_RELP + (bit << 2));
          }
        } else {
          // ---------------- game over: commit ----------------
          const uint32_t a_ = side ? fld : bat, h_ = side ? bat : fld;
          if (COUNT_PA && side) pa_addr -= pa_side;   // back to the away side's counters: where the lane's next game starts
          commit_joint(A_JOINT, H, 4u, a_, h_);
          const uint32_t inn = (hidx >> 1) + 1u;
          red_inc_shared(A_INN + ((inn < MLBSIM_INN_BINS - 1 ? inn : MLBSIM_INN_BINS - 1) << 2));
          if (a_ == h_) red_inc_shared(A_TRUNC);
          hs = HS_DEAD;
#if OPT_REFILL_INLINE
          // next game of this lane: the next unassigned sim of the warp's chunk, if there is one (which lane plays a sim
          // does not matter to its result); if not, the lane stays dead until the ballot refill has fetched a chunk
          const uint32_t k = fetch_inc_shared(A_WNEXT);
          if (k < chunk_end) {
            const unsigned long long sim = (unsigned long long)args.sim_lo + k;
            c0 = (uint32_t)sim;
            c1 = ((uint32_t)(sim >> 32) << RNG_SIMHI_SHIFT) | rng_matchup;
            cur = init0; oth = init1;
            sc = 0; hidx = 0; hs = hs_new_half; used = 0;
          }
#endif
        }
      }
    }
    __syncwarp();  // reconverge before the next iteration's ballot: one loop trip = one PA for all lanes
  }
}

}  // namespace

// Lays one mlbsim_table out the way the native kernel's shared memory wants it (run once per upload):
// alias rows [side][pit 8][tier][16 slots][8 words] with the row of batter s at slot' = s + 16 - L(side); the first
// WRAP_DUP_ROWS rows of a group repeat the first batters of the group whose lineup wrap carries into it (see the context
// word): the group's own batters when tier > 0 (the carry was a TTO increment), the batters of (pit - 1, tier 4+) when
// tier == 0 (the carry ran out of a saturated tier into the pitcher field).  Everything else is zero.  Fatigue
// thresholds and slopes follow, then the header words.  The native kernel stages a matchup with two bulk copies.
extern "C" __global__ void __launch_bounds__(256)
mlbsim_stage_tables_kernel(const mlbsim_table* __restrict__ tables, mlbsim_staged_table* __restrict__ staged, int n) {
  const int m = blockIdx.x;
  if (m >= n) return;
  const mlbsim_table* T = tables + m;
  mlbsim_staged_table* D = staged + m;
  const uint32_t* src = &T->alias[0][0][0][0][0];
  for (int i = threadIdx.x; i < MLBSIM_STAGED_ALIAS_WORDS; i += blockDim.x) {
    // destination word i: row = i / 8 = ((side*8 + pit)*4 + tier)*16 + slot'
    const uint32_t w = (uint32_t)i & 7u, row = (uint32_t)i >> 3, pslot = row & 15u, tier = (row >> 4) & 3u, pit = (row >> 6) & 7u, side = row >> 9;
    const uint32_t L = T->lineup_len[side], bias = 16u - L;
    uint32_t sp = pit, st = tier, sb = 0xFFFFFFFFu;   // source (pitcher, tier, batter); batter 0xFFFFFFFF = none
    if (pslot >= bias) {
      sb = pslot - bias;
    } else if (pslot < (uint32_t)WRAP_DUP_ROWS && pslot < L && (pit | tier) != 0u) {
      sb = pslot;
      if (tier == 0u) { sp = pit - 1u; st = MLBSIM_N_TIERS - 1; }
    }
    uint32_t v = 0;
    if (sb != 0xFFFFFFFFu && sp < (uint32_t)MLBSIM_MAX_PITCHERS)
      v = src[((((side * MLBSIM_MAX_PITCHERS + sp) * MLBSIM_N_TIERS + st) * MLBSIM_MAX_LINEUP) + sb) * MLBSIM_ROW_WORDS + w];
    D->alias[i] = v;
  }
  const uint32_t* fs = &T->fthr[0][0][0][0];
  const int32_t* ss = &T->fslope[0][0][0][0];
  for (int i = threadIdx.x; i < MLBSIM_STAGED_FAT_WORDS; i += blockDim.x) {
    D->fthr[i] = fs[i];
    D->fslope[i] = ss[i];
  }
  if (threadIdx.x == 0) {
    D->hdr[0] = T->flags; D->hdr[1] = T->lineup_len[0]; D->hdr[2] = T->lineup_len[1];
    D->hdr[3] = T->bullpen_len[0]; D->hdr[4] = T->bullpen_len[1];
    D->hdr[5] = D->hdr[6] = D->hdr[7] = 0;
  }
}

extern "C" __global__ void __launch_bounds__(NATIVE_LB_THREADS, NATIVE_LB_BLOCKS)
mlbsim_native_kernel(const NativeArgs args) {
  const int tid = threadIdx.x;
  const uint32_t n_sims = args.n_sims;
  const bool count_pa = (args.flags & 1u) == 0;

  // transition LUT + sub-draw thresholds: derived once per block from the branchy semantics above
  const uint32_t lut_base = smem_static_base() + (uint32_t)offsetof(Smem, lut);
  if (lut_base + LUT_WORDS * 4u > HS_ROW + 4u) __trap();   // the state word keeps LUT row addresses in 14 bits
  for (int i = tid; i < LUT_WORDS; i += blockDim.x) {
    const int ob = i / LUT_ROW_WORDS, w = i % LUT_ROW_WORDS;
    uint32_t v = 0;
    if (w < 32) {
      const int var = w >> 2, k = w & 3;
      const int o = var == MLBSIM_V_1B_ADV ? (int)MLBSIM_1B : var, dB = var == MLBSIM_V_1B_ADV;
      if (k == 0) {
        v = subdraw_a_threshold(o, (ob >> 3) == 2);
      } else if (k < 3) {
        const uint32_t e = transition_entry(o, k - 1, dB, ob >> 3, ob & 7);
        // delta of the state word: (address of the next (outs*8 + bases) row - address of this row) + runs + reached
        // + three-outs flag + one plate appearance; hs += v is the whole transition
        v = ((e & 31u) - (uint32_t)ob) * (uint32_t)(LUT_ROW_WORDS * 4) + (((e >> 16) & 7u) << HS_RUNS_SHIFT) + (((e >> 24) & 1u) << HS_REACHED_SHIFT) +
            (((e >> 30) & 1u) ? HS_3OUT : 0u) + HS_PA_ONE;
      }
    }
    S.lut[i] = v;
  }
#if OPT_TMA
  const uint32_t mbar = smem_static_base() + (uint32_t)offsetof(Smem, mbar);
  uint32_t mbar_phase = 0;
  if (tid == 0) mbar_init(mbar, 1);
#endif

  // Matchup schedule of a block: first the matchups it owns (blockIdx.x, + gridDim.x, ...: a 1 024-point sweep
  // spreads over the grid without two blocks staging the same table), then whatever is still unfinished, found by
  // a block-wide parallel look at the work counters (one load per thread, not one barrier pair per matchup).
  uint32_t next_owned = blockIdx.x;
  for (;;) {
    uint32_t m;
    if (next_owned < args.n_matchups) {
      m = next_owned;
      next_owned += gridDim.x;
    } else {
      const uint32_t start = blockIdx.x % args.n_matchups;
      uint32_t found = 0xFFFFFFFFu;
      for (uint32_t base = 0; base < args.n_matchups; base += blockDim.x) {
        const uint32_t k = base + tid;
        uint32_t c = start + k;
        if (c >= args.n_matchups) c -= args.n_matchups;
        const int open = k < args.n_matchups && *(volatile unsigned long long*)&args.counters[c] < (unsigned long long)n_sims;
        if (__syncthreads_or(open)) {
          if (tid == 0) S.go = 0xFFFFFFFFu;
          __syncthreads();
          if (open) atomicMin(&S.go, k);
          __syncthreads();
          found = start + S.go;
          if (found >= args.n_matchups) found -= args.n_matchups;
          break;
        }
      }
      if (found == 0xFFFFFFFFu) break;   // every matchup's sims are handed out: the blocks holding them finish them
      m = found;
    }
    __syncthreads();  // previous matchup fully flushed (shared memory free for the next table); S.go free for reuse
    if (tid == 0) S.go = (*(volatile unsigned long long*)&args.counters[m] < (unsigned long long)n_sims) ? 1u : 0u;
    __syncthreads();
    if (!S.go) continue;

    // ---- stage this matchup's table; clear block-local accumulators ----
    const mlbsim_staged_table* T = args.tables + m;
    const uint32_t hdr_flags = __ldg(&T->hdr[0]);
    const uint32_t L0 = __ldg(&T->hdr[1]), L1 = __ldg(&T->hdr[2]), nb0 = __ldg(&T->hdr[3]), nb1 = __ldg(&T->hdr[4]);
    const bool fatigue = (hdr_flags & MLBSIM_FLAG_FATIGUE) != 0;
    {
#if OPT_TMA
      // One elected thread arms the barrier with the byte count and hands the copies to the bulk-copy engine; every
      // thread clears accumulators meanwhile and then waits on the barrier's phase.
      if (tid == 0) {
        const uint32_t bytes = ALIAS_BYTES + (fatigue ? 2u * FAT_WORDS * 4u : 0u);
        fence_proxy_async();   // the previous matchup's generic-proxy accesses to these buffers are ordered before the async writes
        mbar_arrive_expect_tx(mbar, bytes);
        bulk_copy_g2s(alias_base(), T->alias, ALIAS_BYTES, mbar);
        if (fatigue) bulk_copy_g2s(smem_static_base() + (uint32_t)offsetof(Smem, fthr), T->fthr, 2u * FAT_WORDS * 4u, mbar);
      }
#else
      const uint4* src = reinterpret_cast<const uint4*>(T->alias);
      uint4* dst = reinterpret_cast<uint4*>(S_pa + ((alias_base() - smem_dynamic_base()) >> 2));   // (OPT_ALIAS_OR; the static layout is `S.alias`)
#if !OPT_ALIAS_OR
      dst = reinterpret_cast<uint4*>(S.alias);
#endif
      for (int i = tid; i < ALIAS_WORDS / 4; i += blockDim.x) dst[i] = __ldg(src + i);
      if (fatigue) {
        const uint4* fsrc = reinterpret_cast<const uint4*>(T->fthr);
        uint4* fdst = reinterpret_cast<uint4*>(S.fthr);
        for (int i = tid; i < 2 * FAT_WORDS / 4; i += blockDim.x) fdst[i] = __ldg(fsrc + i);  // fthr then fslope
      }
#endif
      for (int i = tid; i < JOINT_WORDS; i += blockDim.x) S_pa[i] = 0;   // joint histograms: start of dynamic smem
      if (tid < MLBSIM_INN_BINS) S.innings[tid] = 0;
      if (tid < 16) { S.rel_games[tid] = 0; S.rel_picks[tid] = 0; }
      if (tid == 0) S.truncated = 0;
#if OPT_WRAP_LUT
      if (tid < 256) {
        // what wrap_fix would add for (side, half over?, slot', tier): see its comment
        const uint32_t side = (uint32_t)tid >> 7, over = ((uint32_t)tid >> 6) & 1u, slotp = (uint32_t)tid & 15u, tier = ((uint32_t)tid >> 4) & 3u;
        const uint32_t bias = 16u - (side ? L1 : L0);
        uint32_t fix = 0;
        if (slotp < bias) {
          fix = bias << CX_SLOT_SHIFT;
          if (tier == 0u || (over && slotp == 0u)) fix -= CX_TIER_ONE;
        }
        S.wlut[side * 512u + over * 64u + ((uint32_t)tid & 63u)] = fix;
      }
#endif
      if (count_pa) {
#if OPT_PA_WARP
        if (tid < (int)(DYN_PA_BYTES / 4u)) S_pa[DYN_PA_WORD0 + tid] = 0u;
#else
        for (int f = 0; f < PA_FIELDS; ++f) S_pa[DYN_PA_WORD0 + f * PA_STRIDE + tid] = 0u;
#endif
      }
#if OPT_TMA
      mbar_wait(mbar, mbar_phase);
      mbar_phase ^= 1u;
#endif
    }
    __syncthreads();

    const uint32_t rng_matchup = (args.matchup_base + m) << RNG_MATCHUP_SHIFT;
    unsigned long long* const H = args.hist + (size_t)m * MLBSIM_HIST_WORDS;
    unsigned long long* const counter = args.counters + m;

    const bool short_lineup = !OPT_WRAP_ROUND || (L0 < L1 ? L0 : L1) <= (uint32_t)WRAP_DUP_ROWS;   // more than one wrap per round possible
    if (short_lineup) {   // (rare: tests and the reference's 3-batter demo)
      if (fatigue) {
        if (count_pa) play_matchup<true, true, true>(args, H, counter, rng_matchup, L0, L1, nb0, nb1);
        else play_matchup<true, false, true>(args, H, counter, rng_matchup, L0, L1, nb0, nb1);
      } else {
        if (count_pa) play_matchup<false, true, true>(args, H, counter, rng_matchup, L0, L1, nb0, nb1);
        else play_matchup<false, false, true>(args, H, counter, rng_matchup, L0, L1, nb0, nb1);
      }
    } else if (fatigue) {
      if (count_pa) play_matchup<true, true, false>(args, H, counter, rng_matchup, L0, L1, nb0, nb1);
      else play_matchup<true, false, false>(args, H, counter, rng_matchup, L0, L1, nb0, nb1);
    } else {
      if (count_pa) play_matchup<false, true, false>(args, H, counter, rng_matchup, L0, L1, nb0, nb1);
      else play_matchup<false, false, false>(args, H, counter, rng_matchup, L0, L1, nb0, nb1);
    }

    // ---- block flush: shared u32 bins / PA counters -> device u64 totals ----
#if OPT_PA_WARP
    __syncthreads();
    if (count_pa && tid < PA_FIELDS) {
      uint32_t v = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += S_pa[DYN_PA_WORD0 + w * PA_FIELDS + tid];
      const int var = tid & 7;
      if (v) atomicAdd(&H[H_PA + (tid >> 3) * 8 + (var == MLBSIM_V_1B_ADV ? (int)MLBSIM_1B : var)], (unsigned long long)v);
    }
#else
    if (count_pa) {
      for (int f = 0; f < PA_FIELDS; ++f) {
        const uint32_t v = __reduce_add_sync(0xffffffffu, S_pa[DYN_PA_WORD0 + f * PA_STRIDE + tid]);
        const int var = f & 7;
        if (lane == 0 && v) atomicAdd(&H[H_PA + (f >> 3) * 8 + (var == MLBSIM_V_1B_ADV ? (int)MLBSIM_1B : var)], (unsigned long long)v);
      }
    }
    __syncthreads();
#endif
    for (int i = tid; i < JOINT_WORDS; i += blockDim.x) {
      const uint32_t v = S_pa[i];
      if (v) {
        const int c = i / (SB * SB), a = (i / SB) % SB, h = i % SB;
        atomicAdd(&H[((size_t)c * MLBSIM_HIST_BINS + a) * MLBSIM_HIST_BINS + h], (unsigned long long)v);
      }
    }
    if (tid < MLBSIM_INN_BINS) {
      const uint32_t v = S.innings[tid];
      if (v) atomicAdd(&H[H_INN + tid], (unsigned long long)v);
      const uint32_t games = __reduce_add_sync(0xffffffffu, v);   // every finished game sits in exactly one innings bin
      if (tid == 0 && games) atomicAdd(&H[H_N], (unsigned long long)games);
      if (tid == 0 && S.truncated) atomicAdd(&H[H_TRUNC], (unsigned long long)S.truncated);
    }
    if (tid < 16) {
      if (S.rel_games[tid]) atomicAdd(&H[H_RELG + tid], (unsigned long long)S.rel_games[tid]);
      if (S.rel_picks[tid]) atomicAdd(&H[H_RELP + tid], (unsigned long long)S.rel_picks[tid]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Per-sim variant (small N compatibility path and cross-check of the flat kernel): one thread
// plays one game with plain nested loops, reads thresholds straight from global memory and
// writes a full mlbsim_replay_game record.  Same law, same Philox streams => the same games as
// mlbsim_native_kernel.
// ---------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(128)
mlbsim_persim_kernel(const PersimArgs args) {
  const mlbsim_table* __restrict__ T = args.table;
  for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; g < args.n_sims;
       g += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long sim = args.sim_lo + g;
    const uint32_t c0 = (uint32_t)sim;
    const uint32_t c1base = ((uint32_t)(sim >> 32) << RNG_SIMHI_SHIFT) | (args.matchup_id << RNG_MATCHUP_SHIFT);
    uint32_t seq = 0, w0 = 0, w1 = 0, w1_first = 0;
    __align__(16) mlbsim_replay_game R;
    {
      uint32_t* z = reinterpret_cast<uint32_t*>(&R);
#pragma unroll 1
      for (int i = 0; i < (int)(sizeof(R) / 4); ++i) z[i] = 0;
    }
    int score[2] = {0, 0}, slot[2] = {0, 0}, n_used[2] = {0, 0};
    int pc[2] = {0, 0}, tto[2] = {1, 1}, pit[2] = {0, 0};
    int inning = 1;
    for (;;) {
      int rr[2] = {0, 0};
      for (int side = 0; side < 2; ++side) {
        if (side == 1 && inning == 9 && score[1] > score[0]) break;
        const uint32_t nbp = T->bullpen_len[side];
        if ((pc[side] > PITCH_LIMIT || tto[side] >= 3) && nbp > 0) {
          const uint32_t idx = ((w1 & 0xFFFFu) * nbp) >> 16;  // low half of the game's latest w1
          pit[side] = 1 + (int)idx; pc[side] = 0; tto[side] = 1;
          uint8_t* used = side == 0 ? R.used_home : R.used_away;
          if (n_used[side] < MLBSIM_REPLAY_MAX_USED) used[n_used[side]] = (uint8_t)idx;
          else R.status |= MLBSIM_ST_USED_OVERFLOW;
          n_used[side]++;
        }
        const int L = (int)T->lineup_len[side];
        int outs = 0, bases = 0, runs = 0, n = 0, sl = slot[side];
        bool reached = false;
        while (outs < 3 && n < MAX_PA_PER_HALF) {
          philox2x32_native(c0, c1base + seq, args.keys, w0, w1);
          if (n == 0) w1_first = w1;
          if (sl == 0 && n > 0) tto[side] += 1;
          const int tier = (tto[side] >= 4 ? 4 : tto[side]) - 1;
          int o;
          bool dB;
          if ((T->flags & MLBSIM_FLAG_FATIGUE) && pit[side] == 0 && pc[side] > FATIGUE_START) {
            const uint32_t* row = T->fthr[side][tier][sl];
            const int32_t* sp = T->fslope[side][tier][sl];
            const uint32_t e = (uint32_t)(pc[side] - FATIGUE_START);
            const uint32_t u = w0 >> 1;
            o = 0;
#pragma unroll
            for (int j = 0; j < 6; ++j) o += (u >= row[j] + e * (uint32_t)sp[j]) ? 1 : 0;
            dB = (w1 & 0xFFFFu) < T_080_16;
          } else {
            const uint32_t col = w0 >> 29;
            const uint32_t e = T->alias[side][pit[side]][tier][sl][col];
            o = ((w0 << 3) > e) ? (int)col : (int)(e & 7u);
            dB = o == MLBSIM_V_1B_ADV;
            if (dB) o = MLBSIM_1B;
          }
          const uint32_t tA = subdraw_a_threshold(o, outs == 2);
          const uint32_t ent = transition_entry(o, w1 < tA, dB, outs, bases);
          const int r = (ent >> 16) & 7;
          outs = (ent & 31u) >> 3; bases = ent & 7u; runs += r;
          if ((ent >> 24) & 1u) reached = true;
          R.pa[side][o] += 1;
          pc[side] += 1; n += 1; sl = (sl + 1 == L) ? 0 : sl + 1;
          ++seq;
        }
        if (reached) runs += (runs == 0) ? (int)(w1_first < T_END_BOTH) + (int)(w1_first < T_END_ANY) : (int)(w1_first < T_END_GHOST);
        slot[side] = sl;
        score[side] += runs;
        rr[side] = runs;
      }
      if (inning <= MLBSIM_REPLAY_MAX_INNINGS) {
        R.away_runs[inning - 1] = (uint8_t)(rr[0] > 255 ? 255 : rr[0]);
        R.home_runs[inning - 1] = (uint8_t)(rr[1] > 255 ? 255 : rr[1]);
      } else {
        R.status |= MLBSIM_ST_INNINGS_OVERFLOW;
      }
      if (inning >= 9 && score[0] != score[1]) break;
      if (inning >= MLBSIM_MAX_INNINGS) { R.status |= MLBSIM_ST_TRUNCATED; break; }   // termination guard
      inning += 1;
    }
    R.home_score = score[1]; R.away_score = score[0]; R.n_innings = inning;
    R.tape_used = (int32_t)seq;
    R.n_used[0] = n_used[0]; R.n_used[1] = n_used[1];
    const uint4* s4 = reinterpret_cast<const uint4*>(&R);
    uint4* d4 = reinterpret_cast<uint4*>(args.out + g);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(R) / 16); ++i) d4[i] = s4[i];
  }
}

size_t mlbsim_native_smem_bytes(int /*threads_per_block*/) {   // dynamic part only; independent of the block size
#if OPT_ALIAS_OR
  return (size_t)DYN_JOINT_BYTES + (size_t)DYN_PA_BYTES + 0x8000u + (size_t)ALIAS_BYTES;   // + alignment slack + the alias table
#else
  return (size_t)DYN_JOINT_BYTES + (size_t)DYN_PA_BYTES;
#endif
}
