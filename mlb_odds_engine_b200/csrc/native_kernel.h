// Launch interface between the C-ABI translation unit and the kernels (internal, not exported).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../../include/mlbsim.h"
#include "philox.cuh"

#ifndef MLBSIM_NATIVE_MAX_THREADS
#define MLBSIM_NATIVE_MAX_THREADS 640
#endif

// Device-only image of one matchup, laid out exactly as the native kernel's shared memory wants it (built once per
// upload by mlbsim_stage_tables_kernel): the kernel stages a matchup with cp.async.bulk copies, no address arithmetic.
#define MLBSIM_STAGED_ALIAS_WORDS (2 * 8 * MLBSIM_N_TIERS * 16 * MLBSIM_ROW_WORDS) /* [side][pit 8][tier][16 slots][8]: 32 KB */
#define MLBSIM_STAGED_FAT_WORDS (2 * MLBSIM_N_TIERS * MLBSIM_MAX_LINEUP * MLBSIM_ROW_WORDS)
struct alignas(16) mlbsim_staged_table {
  uint32_t alias[MLBSIM_STAGED_ALIAS_WORDS];
  uint32_t fthr[MLBSIM_STAGED_FAT_WORDS];   // fthr and fslope stay contiguous: one copy
  int32_t fslope[MLBSIM_STAGED_FAT_WORDS];
  uint32_t hdr[8];                          // flags, lineup_len[2], bullpen_len[2], 0, 0, 0
};

struct NativeArgs {
  const mlbsim_staged_table* tables;  // device, n_matchups entries (already offset to matchup_lo)
  unsigned long long* hist;      // device, n_matchups * MLBSIM_HIST_WORDS uint64 counters
  unsigned long long* counters;  // device, n_matchups work counters, zero at launch
  uint64_t sim_lo;               // first sim index of this launch
  uint32_t n_sims;               // sims per matchup in this launch (< 2^32)
  uint32_t n_matchups;
  uint32_t matchup_base;         // global id of tables[0] (Philox counter field)
  uint32_t chunk;                // sims a warp takes per fetch
  uint32_t flags;                // bit 0: skip PA outcome counters
  Philox2Keys keys;
};

struct ReplayArgs {
  const mlbsim_params* params;  // device
  const double* tape;           // device
  const uint64_t* offsets;      // device, n_games + 1
  mlbsim_replay_game* out;      // device
  long long n_games;
  // the launch-uniform split points of resolve_contact, by value: kernel arguments sit in the constant bank, so a compare
  // reads them as an operand (no shared-memory load, no register)
  double cdf_bip[4], cdf_hit[3], cdf_fb[2];
  double consts[5];   // 1.01, 0.10, 0.4, 0.14, 0.8: the step's float64 literals without an immediate form (REPLAY_CONST_BANK)
};

struct PersimArgs {
  const mlbsim_table* table;  // device, one matchup
  mlbsim_replay_game* out;    // device, n_sims records
  unsigned long long sim_lo;
  unsigned long long n_sims;
  uint32_t matchup_id;        // Philox counter field
  Philox2Keys keys;
};

#ifdef __CUDACC__
extern "C" __global__ void mlbsim_persim_kernel(const PersimArgs args);
extern "C" __global__ void mlbsim_native_kernel(const NativeArgs args);
extern "C" __global__ void mlbsim_stage_tables_kernel(const mlbsim_table* tables, mlbsim_staged_table* staged, int n);
extern "C" __global__ void mlbsim_replay_flat_kernel(const ReplayArgs args, unsigned long long* counter, uint32_t chunk);
extern "C" __global__ void mlbsim_replay_flat_nonoise_kernel(const ReplayArgs args, unsigned long long* counter, uint32_t chunk);
extern "C" __global__ void mlbsim_replay_flat_regular_kernel(const ReplayArgs args, unsigned long long* counter, uint32_t chunk);
extern "C" __global__ void mlbsim_summary_kernel(const mlbsim_hist* hist, mlbsim_summary* out);
extern "C" __global__ void mlbsim_alias_rows_kernel(const double* probs, int n_rows, int n_own, uint32_t* out);

// table_kernel.cu: mlbsim_params -> mlbsim_table on the device; grid = n_matchups * MLBSIM_TABLE_BUILD_SPLIT blocks of 256
// threads; status[m] (zeroed by the caller) becomes 1 when somebody can reach base in matchup m
#define MLBSIM_TABLE_BUILD_SPLIT 8
extern "C" __global__ void mlbsim_build_tables_kernel(const mlbsim_params* params, mlbsim_table* tables, uint32_t* status, int n_matchups);
extern "C" cudaError_t mlbsim_table_set_nodes(const double* y, const double* w, cudaStream_t stream);
#endif

size_t mlbsim_native_smem_bytes(int threads_per_block);
#define MLBSIM_SUMMARY_THREADS 256
// Replay kernel.  MLBSIM_REPLAY_RING = 1: every lane streams its tape through a 256-byte ring of its own in shared memory
// (cp.async, each byte fetched from HBM once, the plate-appearance window read with LDS); 0: windows loaded straight from
// global memory (the round-2 "flat" kernel, kept for A/B).
#ifndef MLBSIM_REPLAY_RING
#define MLBSIM_REPLAY_RING 1
#endif
#define MLBSIM_REPLAY_RING_LANE_BYTES 256
#ifndef MLBSIM_REPLAY_FLAT_THREADS
#define MLBSIM_REPLAY_FLAT_THREADS (MLBSIM_REPLAY_RING ? 800 : 256)   /* ring: 25 warps = what 80 registers allow (768 / 800 / 832 at 72 registers: 5.79 / 5.95 / 5.94e8, profiles/r02_rp10.log) */
#endif
#ifndef MLBSIM_REPLAY_FLAT_BLOCKS
#define MLBSIM_REPLAY_FLAT_BLOCKS (MLBSIM_REPLAY_RING ? 1 : 3)   /* the rings take 200 KB of shared memory; the window-load kernel runs 3 x 256 threads */
#endif
#define MLBSIM_REPLAY_DYN_SMEM (MLBSIM_REPLAY_RING ? (size_t)MLBSIM_REPLAY_FLAT_THREADS * MLBSIM_REPLAY_RING_LANE_BYTES : (size_t)0)
