// alias_row.cuh — one Walker/Vose alias row on the device (shared by alias_kernel.cu and table_kernel.cu).
//
// Host statement: mlb_odds_engine_b200/tables.py::alias_row (and its lock-step batch form alias_rows).  The SAME float64
// operations in the SAME order (explicit round-to-nearest intrinsics: no FMA contraction), so the output is bit-identical
// to the host's for the same input probabilities.
//
// Row law: probabilities p[0..n_own) (n_own = 7 or 8) on 8 columns; q = p * 8 / sum(p); columns with q < 1 are topped up
// from columns with q >= 1, highest column index first (two stacks).  Entry = (2^29 - 1 - cut29) << 3 | alias (own outcome
// iff (draw << 3) > entry: exactly cut29 of a column's 2^29 draws); a full column is encoded as "cut 0, alias = own outcome".  (What the alias row is for: include/mlbsim.h, native mode.)
#pragma once
#include <stdint.h>

#include "../../include/mlbsim.h"

__device__ __forceinline__ void mlbsim_alias_row(const double* p_in, int n_own, uint32_t* out) {
  constexpr int C = MLBSIM_ROW_WORDS;   // 8 columns
  double p[C], q[C], cut[C];
  int alias[C], small[C], large[C];
#pragma unroll
  for (int i = 0; i < C; ++i) p[i] = i < n_own ? p_in[i] : 0.0;
  // numpy's sum over a contiguous axis: fewer than 8 elements left to right, exactly 8 as ((0+1)+(2+3))+((4+5)+(6+7))
  double s;
  if (n_own == C) {
    s = __dadd_rn(__dadd_rn(__dadd_rn(p[0], p[1]), __dadd_rn(p[2], p[3])), __dadd_rn(__dadd_rn(p[4], p[5]), __dadd_rn(p[6], p[7])));
  } else {
    s = p[0];
    for (int i = 1; i < n_own; ++i) s = __dadd_rn(s, p[i]);
  }
  int ns = 0, nl = 0;
#pragma unroll
  for (int i = 0; i < C; ++i) {
    q[i] = i < n_own ? __ddiv_rn(__dmul_rn(p[i], (double)C), s) : 0.0;
    cut[i] = 1.0;
    alias[i] = i;
  }
  for (int i = 0; i < C; ++i) {
    if (q[i] < 1.0) small[ns++] = i; else large[nl++] = i;
  }
  while (ns > 0 && nl > 0) {
    const int sm = small[--ns], lg = large[--nl];
    cut[sm] = q[sm];
    alias[sm] = lg;
    q[lg] = __dsub_rn(q[lg], __dsub_rn(1.0, q[sm]));
    if (q[lg] < 1.0) small[ns++] = lg; else large[nl++] = lg;
  }
  int top = 0;                                  // np.argmax(p): first index of the maximum
  for (int i = 1; i < n_own; ++i) if (p[i] > p[top]) top = i;
  for (int j = 0; j < ns; ++j) { const int c = small[j]; cut[c] = 1.0; alias[c] = c < n_own ? c : top; }
  for (int j = 0; j < nl; ++j) { const int c = large[j]; cut[c] = 1.0; alias[c] = c < n_own ? c : top; }
  const long long one = 1ll << 29;
  for (int i = 0; i < C; ++i) {
    long long c = (long long)floor(__dadd_rn(__dmul_rn(cut[i], (double)one), 0.5));
    int a = alias[i];
    const bool full = c >= one;
    if (full && i < n_own) a = i;
    if (full || i >= n_own) c = 0;
    if (c < 0) c = 0;
    out[i] = (uint32_t)(((one - 1 - c) << 3) | (long long)a);   // complemented cut: own iff (w0 << 3) > entry
  }
}
