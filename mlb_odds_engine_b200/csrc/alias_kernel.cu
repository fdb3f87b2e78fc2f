// alias_kernel.cu — Walker/Vose alias rows on the device: one thread builds one row (alias_row.cuh).
//
// Bit-identical to mlb_odds_engine_b200/tables.py::alias_rows (tested on the GPU box against random rows and against every
// row of real tables): tables built through it hash to the same digests as tables built on the CPU.  A 1 024-point sweep
// has 73 728 rows; numpy needs ~60 ms for them, this ~20 us.
#include <cuda_runtime.h>
#include <stdint.h>

#include "alias_row.cuh"
#include "native_kernel.h"

extern "C" __global__ void __launch_bounds__(128)
mlbsim_alias_rows_kernel(const double* __restrict__ probs, int n_rows, int n_own, uint32_t* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  mlbsim_alias_row(probs + (size_t)r * n_own, n_own, out + (size_t)r * MLBSIM_ROW_WORDS);
}
