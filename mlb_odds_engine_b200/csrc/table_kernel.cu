// table_kernel.cu — the native-mode sampling tables of a batch of matchups, built on the device from their raw numbers.
//
// Host statement: mlb_odds_engine_b200/tables.py::tables_from_arrays (the marginal plate-appearance law of
// core/pa_simulator.py:91-143 + core/bip_resolution.py + core/fatigue_modeling.py with the Beta(30p, 30(1-p)) noise of
// pa_simulator.py:23-32 integrated out).  Input: one mlbsim_params per matchup (the replay-mode struct: the same raw
// float64 numbers).  Output: one mlbsim_table per matchup, written straight into the context's device tables — a
// 1 024-point sweep never materialises 20 MB of tables on the host.
//
// Eight blocks per matchup, one warp per table row (side, pitcher, TTO tier, lineup slot):
//   * effective K / BB rates, the contact split and the cumulative law: plain float64 arithmetic in the host's operation
//     order, compiled with -fmad=false — bit-identical to numpy;
//   * E[min(1, X + 1.01 Y)], X ~ Beta(30k, 30(1-k)), Y ~ Beta(30bb, 30(1-bb)): 48-node Gauss-Legendre over y (lanes =
//     nodes), closed form in x through the regularised incomplete beta function (continued fraction, evaluated directly
//     in the UPPER tail, where the host computes 1 - betainc and cancels).  This is the one place where the device and
//     the host (scipy) do not perform the same operations: the integral agrees to ~1e-16 absolute, so P(K or BB) can
//     differ in its last bit and a 29/31-bit threshold derived from it by one unit, rarely (tests/test_gpu_parity.py
//     bounds it: every entry within one unit, alias targets equal);
//   * Walker alias row (alias_row.cuh, bit-identical to the host for the same probabilities), and for starters that are
//     never replaced the 31-bit thresholds at 75 pitches with the integer slope per pitch fitted over 100 pitches.
// A matchup in which no plate appearance of either side can put a runner on base (the game would never end:
// core/game_simulator.py:110-111 has no inning cap) is reported through `status` (0 = nobody can reach base).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "alias_row.cuh"
#include "native_kernel.h"

#define GL_NODES 48
__constant__ double c_gl_y[GL_NODES];   // Gauss-Legendre nodes on (0, 1)
__constant__ double c_gl_w[GL_NODES];   // weights (sum 1)

extern "C" cudaError_t mlbsim_table_set_nodes(const double* y, const double* w, cudaStream_t stream) {
  cudaError_t e = cudaMemcpyToSymbolAsync(c_gl_y, y, sizeof(double) * GL_NODES, 0, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return e;
  return cudaMemcpyToSymbolAsync(c_gl_w, w, sizeof(double) * GL_NODES, 0, cudaMemcpyHostToDevice, stream);
}

namespace {

constexpr double NOISE_WEIGHT = 30.0;        // pa_simulator.py:23
constexpr double INFIELD_FALLBACK = 0.10;    // pa_simulator.py:74
constexpr double SLOPE_SPAN = 100.0;         // pitches over which the integer slope is fitted (tables.py)
constexpr double P_FIRST_TO_SECOND = 0.8;    // half_inning_simulator.py:105

// continued fraction of the incomplete beta function (modified Lentz)
__device__ double beta_cf(double a, double b, double x) {
  const double tiny = 1e-300, eps = 1e-16;
  const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (fabs(d) < tiny) d = tiny;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m <= 300; ++m) {
    const double m2 = 2.0 * m;
    double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c;
    if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    h *= d * c;
    aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d;
    if (fabs(d) < tiny) d = tiny;
    c = 1.0 + aa / c;
    if (fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (fabs(del - 1.0) < eps) break;
  }
  return h;
}

// Q = 1 - I_x(a, b) (upper tail) and t = x^a (1-x)^b / (a B(a, b)) = I_x(a, b) - I_x(a + 1, b), for 0 < x < 1
__device__ void beta_upper(double a, double b, double x, double lbeta, double& Q, double& t) {
  const double lt = a * log(x) + b * log1p(-x) - lbeta;
  const double bt = exp(lt);
  t = bt / a;
  if (x < (a + 1.0) / (a + b + 2.0)) Q = 1.0 - bt * beta_cf(a, b, x) / a;
  else Q = bt * beta_cf(b, a, 1.0 - x) / b;
}

// E[min(1, X + 1.01 Y)] — all 32 lanes of the warp call this with the same (k, bb); k < 0 means "X = 0" (k <= 0 on the host)
__device__ double expect_clip_warp(double k, double bb, int lane) {
  const bool with_k = k > 0.0;
  const double ay = NOISE_WEIGHT * bb, by = NOISE_WEIGHT * (1.0 - bb);
  const double lby = lgamma(ay) + lgamma(by) - lgamma(ay + by);
  double ax = 0.0, bx = 0.0, lbx = 0.0;
  if (with_k) {
    ax = NOISE_WEIGHT * k, bx = NOISE_WEIGHT * (1.0 - k);
    lbx = lgamma(ax) + lgamma(bx) - lgamma(ax + bx);
  }
  double part = 0.0;
  for (int j = lane; j < GL_NODES; j += 32) {
    const double y = c_gl_y[j];
    const double wy = c_gl_w[j] * exp((ay - 1.0) * log(y) + (by - 1.0) * log1p(-y) - lby);
    const double c = 1.0 - 1.01 * y;    // X must exceed c for the clip to bite
    double excess;
    if (!with_k) {
      excess = fmax(-c, 0.0);
    } else if (c < 0.0) {
      excess = k - c;
    } else {
      // E[(X - c)^+] = k (1 - I_c(a+1, b)) - c (1 - I_c(a, b)) = (k - c) Q + k t
      double Q, t;
      beta_upper(ax, bx, c, lbx, Q, t);
      excess = (k - c) * Q + k * t;
    }
    part += wy * excess;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
  const double mean = (with_k ? k : 0.0) + 1.01 * bb;
  return fmin(fmax(mean - part, 0.0), 1.0);
}

// (P(K), P(K or BB)) — tables.py::_p_k_or_bb_batch
__device__ void p_k_or_bb(double k, double bb, bool noise, int lane, double& pk, double& pkbb) {
  pk = fmin(fmax(k, 0.0), 1.0);
  if (!noise) {
    pkbb = fmax(pk, fmin(fmax(k + bb * 1.01, 0.0), 1.0));
    return;
  }
  if (bb <= 0.0) pkbb = pk;
  else if (bb >= 1.0 || k >= 1.0) pkbb = 1.0;
  else pkbb = expect_clip_warp(k <= 0.0 ? -1.0 : k, bb, lane);
}

struct RowLaw {
  double c[6];   // cumulative K, K+BB, +1B, +2B, +3B, +HR
};

// tables.py::_cdf_rows for one row at fatigue level fl (0 = at most 75 pitches)
__device__ RowLaw row_law(const mlbsim_params& p, int s, int pit, int tier, int slot, double fl, int lane) {
  const double TTO[4] = {0.0, 0.015, 0.035, 0.060};   // fatigue_modeling.py:25-32
  const double pen = TTO[tier];
  const double ak = p.pit_k[s][pit] * (1 - pen) * (1.0 - 0.02 * fl);
  const double ab = p.pit_bb[s][pit] * (1 + pen) * (1.0 + 0.03 * fl);
  const double k = (p.bat_k[s][slot] + ak) / 2 * p.k_mod;
  const double bb = (p.bat_bb[s][slot] + ab) / 2 * p.bb_mod;
  double pk, pkbb;
  p_k_or_bb(k, bb, p.use_noise != 0, lane, pk, pkbb);
  const double contact = 1.0 - pkbb;
  const double w0 = p.cdf_bip[0] - 0.0, w1 = p.cdf_bip[1] - p.cdf_bip[0], w2 = p.cdf_bip[2] - p.cdf_bip[1], w3 = p.cdf_bip[3] - p.cdf_bip[2];
  const double h0 = p.cdf_hit[0] - 0.0, h1 = p.cdf_hit[1] - p.cdf_hit[0], h2 = p.cdf_hit[2] - p.cdf_hit[1];
  const double f0 = p.cdf_fb[0] - 0.0, f1 = p.cdf_fb[1] - p.cdf_fb[0];
  const double hr = fmin(fmax(p.pit_hr[s][pit], 0.0), 1.0);
  const double* hp = p.hit_prob[s][pit][slot];
  const double ph = hp[0] * w0 + hp[1] * w1 + hp[2] * w2 + hp[3] * w3;
  const double s1 = (1 - hr) * (ph * h0 + (1 - ph) * INFIELD_FALLBACK * f0);
  const double s2 = (1 - hr) * (ph * h1 + (1 - ph) * INFIELD_FALLBACK * f1);
  const double s3 = (1 - hr) * ph * h2;
  RowLaw L;
  L.c[0] = pk;
  L.c[1] = pkbb;
  L.c[2] = pkbb + contact * s1;
  L.c[3] = pkbb + contact * (s1 + s2);
  L.c[4] = pkbb + contact * (s1 + s2 + s3);
  L.c[5] = pkbb + contact * (s1 + s2 + s3 + hr);
  double run = L.c[0];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    run = fmax(run, L.c[i]);
    L.c[i] = fmin(run, 1.0);
  }
  return L;
}

__device__ long long quantise31(double c) {
  const double q = floor(c * 2147483648.0 + 0.5);
  return (long long)fmin(fmax(q, 0.0), 2147483648.0);
}

}  // namespace

// grid = n_matchups * MLBSIM_TABLE_BUILD_SPLIT blocks: block (m, part) fills its share of table m's words with the defaults and
// builds rows part, part + SPLIT * warps, ... — a single matchup still spreads over 8 blocks x 8 warps.  `status` must be
// zeroed by the caller: bit 0 is set as soon as somebody can reach base.
extern "C" __global__ void __launch_bounds__(256)
mlbsim_build_tables_kernel(const mlbsim_params* __restrict__ params, mlbsim_table* __restrict__ tables, uint32_t* __restrict__ status,
                           int n_matchups) {
  const int m = blockIdx.x / MLBSIM_TABLE_BUILD_SPLIT, part = blockIdx.x % MLBSIM_TABLE_BUILD_SPLIT;
  if (m >= n_matchups) return;
  const mlbsim_params& p = params[m];
  mlbsim_table& T = tables[m];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  constexpr int ROWS = 2 * MLBSIM_MAX_PITCHERS * MLBSIM_N_TIERS * MLBSIM_MAX_LINEUP;
  constexpr int HDR = offsetof(mlbsim_table, alias) / 4;
  const bool flagged = p.bullpen_len[0] == 0 || p.bullpen_len[1] == 0;   // some starter is never replaced
  uint32_t* words = reinterpret_cast<uint32_t*>(&T);
  if (part == 0 && threadIdx.x < HDR) {
    uint32_t v = 0u;
    switch (threadIdx.x) {
      case offsetof(mlbsim_table, magic) / 4: v = MLBSIM_TABLE_MAGIC; break;
      case offsetof(mlbsim_table, flags) / 4: v = flagged ? MLBSIM_FLAG_FATIGUE : 0u; break;
      case offsetof(mlbsim_table, lineup_len) / 4: v = (uint32_t)p.lineup_len[0]; break;
      case offsetof(mlbsim_table, lineup_len) / 4 + 1: v = (uint32_t)p.lineup_len[1]; break;
      case offsetof(mlbsim_table, bullpen_len) / 4: v = (uint32_t)p.bullpen_len[0]; break;
      case offsetof(mlbsim_table, bullpen_len) / 4 + 1: v = (uint32_t)p.bullpen_len[1]; break;
      default: break;
    }
    words[threadIdx.x] = v;
  }
  bool reach = false;
  for (int r = part * n_warps + warp; r < ROWS; r += n_warps * MLBSIM_TABLE_BUILD_SPLIT) {
    const int slot = r % MLBSIM_MAX_LINEUP, tier = (r / MLBSIM_MAX_LINEUP) % MLBSIM_N_TIERS,
              pit = (r / (MLBSIM_MAX_LINEUP * MLBSIM_N_TIERS)) % MLBSIM_MAX_PITCHERS, s = r / (MLBSIM_MAX_LINEUP * MLBSIM_N_TIERS * MLBSIM_MAX_PITCHERS);
    const bool valid = pit <= p.bullpen_len[s] && slot < p.lineup_len[s];   // (warp-uniform)
    const bool fat = valid && flagged && pit == 0;
    uint32_t* arow = T.alias[s][pit][tier][slot];
    if (!valid) {   // unused rows: zero words (never reached)
      if (lane < MLBSIM_ROW_WORDS) arow[lane] = 0u;
    } else {
      const RowLaw L0 = row_law(p, s, pit, tier, slot, 0.0, lane);
      RowLaw L1 = L0;
      if (fat) L1 = row_law(p, s, pit, tier, slot, SLOPE_SPAN / 25, lane);
      if (lane == 0) {
        double probs[8];
        probs[0] = L0.c[0] - 0.0;
#pragma unroll
        for (int i = 1; i < 6; ++i) probs[i] = L0.c[i] - L0.c[i - 1];
        probs[6] = 1.0 - L0.c[5];
        if (probs[0] + probs[6] < 1.0) reach = true;       // somebody can reach base (K = column 0, OUT = column 6)
        const double adv = probs[2] * P_FIRST_TO_SECOND;   // tables.py::outcome_variants
        probs[7] = adv;
        probs[2] = probs[2] - adv;
        mlbsim_alias_row(probs, 8, arow);
        if (fat) {
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const long long q0 = quantise31(L0.c[i]), q1 = quantise31(L1.c[i]);
            T.fthr[s][tier][slot][i] = (uint32_t)q0;
            T.fslope[s][tier][slot][i] = (int32_t)rint((double)(q1 - q0) / SLOPE_SPAN);
          }
        }
      }
    }
    // pitch-count rows exist for the starters only (pit == 0 owns them); defaults: threshold 2^31, slope 0
    if (pit == 0 && lane < MLBSIM_ROW_WORDS && (!fat || lane >= 6)) {
      T.fthr[s][tier][slot][lane] = 2147483648u;
      T.fslope[s][tier][slot][lane] = 0;
    }
  }
  if (reach) atomicOr(&status[m], 1u);
}
