// summary_kernel.cu — per-matchup scalar reduction of native-mode histograms (sweeps, calibration).
//
// Replaces the per-grid-point Python reductions of tools/test_stuff_plus_sensitivity.py:71-85 (outcome
// rates from PA_RECAP_LOG) and tools/simulate_calibration_game.py:94-128 (run means / SDs, win counts):
// one block per matchup reads its mlbsim_hist (165 KB, coalesced u64 loads) and writes one 320-byte
// mlbsim_summary of exact integer moments.  HBM-bound: 165 KB in, 320 B out per matchup.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mlbsim.h"
#include "native_kernel.h"

namespace {

constexpr int N_ACC = 16;  // sum_runs[5][2], sum_sq[2], sum_cross, home_wins, away_wins, (pad)
constexpr int THREADS = MLBSIM_SUMMARY_THREADS;

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

}  // namespace

extern "C" __global__ void __launch_bounds__(MLBSIM_SUMMARY_THREADS)
mlbsim_summary_kernel(const mlbsim_hist* __restrict__ hist, mlbsim_summary* __restrict__ out) {
  const mlbsim_hist* H = hist + blockIdx.x;
  mlbsim_summary* S = out + blockIdx.x;
  __shared__ unsigned long long part[THREADS / 32][N_ACC];
  unsigned long long acc[N_ACC];
#pragma unroll
  for (int i = 0; i < N_ACC; ++i) acc[i] = 0;

  constexpr int CELLS = MLBSIM_HIST_BINS * MLBSIM_HIST_BINS;
  const unsigned long long* J = reinterpret_cast<const unsigned long long*>(&H->joint[0][0][0]);
#pragma unroll
  for (int cap = 0; cap < MLBSIM_N_CAPS; ++cap) {
    for (int i = threadIdx.x; i < CELLS; i += THREADS) {
      const unsigned long long c = __ldg(J + cap * CELLS + i);
      if (c == 0) continue;
      const unsigned long long a = (unsigned)i / MLBSIM_HIST_BINS, h = (unsigned)i % MLBSIM_HIST_BINS;
      acc[cap * 2 + 0] += c * a;
      acc[cap * 2 + 1] += c * h;
      if (cap == MLBSIM_N_CAPS - 1) {
        acc[10] += c * a * a;
        acc[11] += c * h * h;
        acc[12] += c * a * h;
        if (h > a) acc[13] += c;
        if (a > h) acc[14] += c;
      }
    }
  }
  // innings: one warp's worth of bins
  if (threadIdx.x < MLBSIM_INN_BINS) acc[15] = (unsigned long long)__ldg(reinterpret_cast<const unsigned long long*>(&H->innings[threadIdx.x])) * threadIdx.x;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < N_ACC; ++i) {
    const unsigned long long v = warp_sum(acc[i]);
    if (lane == 0) part[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < N_ACC) {
    unsigned long long v = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) v += part[w][threadIdx.x];
    const int i = threadIdx.x;
    if (i < 10) S->sum_runs[i >> 1][i & 1] = v;
    else if (i < 12) S->sum_sq[i - 10] = v;
    else if (i == 12) S->sum_cross = v;
    else if (i == 13) S->home_wins = v;
    else if (i == 14) S->away_wins = v;
    else S->sum_innings = v;
  }
  if (threadIdx.x >= 32 && threadIdx.x < 48) S->pa[(threadIdx.x - 32) >> 3][(threadIdx.x - 32) & 7] = H->pa[(threadIdx.x - 32) >> 3][(threadIdx.x - 32) & 7];
  if (threadIdx.x == 48) S->n_games = H->n_games;
  if (threadIdx.x == 49) S->n_truncated = H->n_truncated;
  if (threadIdx.x >= 64 && threadIdx.x < 70) S->reserved[threadIdx.x - 64] = 0;
}
