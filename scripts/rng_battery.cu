// Statistical battery for the native law's generator at reduced round counts ("what the generator costs").
//
// The product draws ONE Philox2x32-10 block (w0, w1) per plate appearance, counter = (sim index, PA number | matchup),
// key = f(seed) (csrc/philox.cuh).  The counters are sequential and low-entropy, so what a shortened generator must
// deliver is diffusion from adjacent counters — this battery measures exactly that, on the stream in the order and with
// the bit fields the kernel consumes (alias column = w0 >> 29, 29-bit fraction, w1 thresholds, low 16 bits of w1):
//   avalanche      flip each of the 64 counter bits and 32 key bits: P(output bit flips) = 1/2 for all 96 x 64 pairs
//   hamming        popcount(out(c) ^ out(c')) ~ Binomial(64, 1/2) for c' = next sim / next PA
//   freq16         chi-square of the four 16-bit halves of (w0, w1) over 65 536 cells
//   serial         12-bit x 12-bit pair tables: w0 of consecutive PAs, w0 of consecutive sims, (w0, w1) of one block,
//                  (w1 of a PA, w0 of the next)
//   columns        the 3-bit alias columns of 5 consecutive PAs (32 768 cells)
//   runs           run-up lengths of w0 over the PAs of a game (Knuth's test with the breaking element discarded)
//   birthday       Marsaglia's birthday spacings, 4 096 birthdays in a year of 2^32 days (lambda = 4), on w0 and w1
// Raw statistics go to stdout as JSON; scripts/rng_battery.py turns them into p-values.  Test code: not part of the product
// library.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o variants/rng_battery scripts/rng_battery.cu
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../mlb_odds_engine_b200/csrc/philox.cuh"

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));                \
      exit(1);                                                                                  \
    }                                                                                           \
  } while (0)

template <int R>
__device__ __forceinline__ void gen(uint32_t c0, uint32_t c1, uint32_t key, uint32_t& o0, uint32_t& o1) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const uint64_t p = (uint64_t)PHILOX2_M * c0;
    c0 = (uint32_t)(p >> 32) ^ key ^ c1;
    c1 = (uint32_t)p;
    key += PHILOX_W0;
  }
  o0 = c0;
  o1 = c1;
}

// the product's counter for (sim, PA number, matchup)
__device__ __forceinline__ void ctr(uint64_t sim, uint32_t pa, uint32_t matchup, uint32_t& c0, uint32_t& c1) {
  c0 = (uint32_t)sim;
  c1 = pa | (matchup << RNG_MATCHUP_SHIFT) | ((uint32_t)(sim >> 32) << RNG_SIMHI_SHIFT);
}

#define PA_SPAN 128u   // PA numbers 0..127 of a game

// ---- avalanche -------------------------------------------------------------------------------------------------------
template <int R>
__global__ void k_avalanche(uint32_t key, uint64_t n_base, unsigned long long* counts /* [96][65]: 64 output bits, then n */) {
  __shared__ unsigned int s[96 * 65];
  for (int i = threadIdx.x; i < 96 * 65; i += blockDim.x) s[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_base; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t c0, c1, a0, a1;
    ctr(i / PA_SPAN, (uint32_t)(i % PA_SPAN), 0, c0, c1);
    gen<R>(c0, c1, key, a0, a1);
    for (int b = 0; b < 96; ++b) {
      uint32_t b0, b1;
      bool use = true;   // the bases are sequential: (c, c ^ bit) would be met from both ends, so only the end with the bit clear counts
      if (b < 32) gen<R>(c0 ^ (1u << b), c1, key, b0, b1), use = !((c0 >> b) & 1u);
      else if (b < 64) gen<R>(c0, c1 ^ (1u << (b - 32)), key, b0, b1), use = !((c1 >> (b - 32)) & 1u);
      else gen<R>(c0, c1, key ^ (1u << (b - 64)), b0, b1);
      const uint32_t d0 = use ? a0 ^ b0 : 0u, d1 = use ? a1 ^ b1 : 0u;
      const unsigned mu = __ballot_sync(0xFFFFFFFFu, use);
      if (lane == 0) atomicAdd(&s[b * 65 + 64], __popc(mu));
      for (int o = 0; o < 32; ++o) {
        const unsigned m0 = __ballot_sync(0xFFFFFFFFu, (d0 >> o) & 1u), m1 = __ballot_sync(0xFFFFFFFFu, (d1 >> o) & 1u);
        if (lane == 0) {
          atomicAdd(&s[b * 65 + o], __popc(m0));
          atomicAdd(&s[b * 65 + 32 + o], __popc(m1));
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 96 * 65; i += blockDim.x) atomicAdd(&counts[i], (unsigned long long)s[i]);
}

// ---- hamming weight of the output difference between neighbouring counters -------------------------------------------
template <int R>
__global__ void k_hamming(uint32_t key, uint64_t n, unsigned long long* hist /* [2][65] */) {
  __shared__ unsigned int s[2 * 65];
  for (int i = threadIdx.x; i < 130; i += blockDim.x) s[i] = 0;
  __syncthreads();
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t c0, c1, a0, a1, b0, b1;
    const uint64_t sim = i / PA_SPAN;
    const uint32_t pa = (uint32_t)(i % PA_SPAN);
    ctr(sim, pa, 0, c0, c1);
    gen<R>(c0, c1, key, a0, a1);
    ctr(sim + 1, pa, 0, c0, c1);
    gen<R>(c0, c1, key, b0, b1);
    atomicAdd(&s[__popc(a0 ^ b0) + __popc(a1 ^ b1)], 1u);
    ctr(sim, pa + 1, 0, c0, c1);
    gen<R>(c0, c1, key, b0, b1);
    atomicAdd(&s[65 + __popc(a0 ^ b0) + __popc(a1 ^ b1)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 130; i += blockDim.x) atomicAdd(&hist[i], (unsigned long long)s[i]);
}

// ---- 16-bit field frequencies ------------------------------------------------------------------------------------------
template <int R>
__global__ void k_freq16(uint32_t key, uint64_t n, unsigned int* hist /* [4][65536] */) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t c0, c1, a0, a1;
    ctr(i / PA_SPAN, (uint32_t)(i % PA_SPAN), 0, c0, c1);
    gen<R>(c0, c1, key, a0, a1);
    atomicAdd(&hist[a0 >> 16], 1u);
    atomicAdd(&hist[65536 + (a0 & 0xFFFFu)], 1u);
    atomicAdd(&hist[2 * 65536 + (a1 >> 16)], 1u);
    atomicAdd(&hist[3 * 65536 + (a1 & 0xFFFFu)], 1u);
  }
}

// ---- serial pairs (12 x 12 bits), one table per launch so it stays in L2 ------------------------------------------------
template <int R, int KIND>
__global__ void k_serial(uint32_t key, uint64_t n, unsigned int* hist /* [1 << 24] */) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t c0, c1, a0, a1, b0, b1;
    const uint64_t sim = i / PA_SPAN;
    const uint32_t pa = (uint32_t)(i % PA_SPAN);
    ctr(sim, pa, 0, c0, c1);
    gen<R>(c0, c1, key, a0, a1);
    uint32_t x, y;
    if (KIND == 0) {          // w0 of consecutive PAs of a game
      ctr(sim, pa + 1, 0, c0, c1);
      gen<R>(c0, c1, key, b0, b1);
      x = a0 >> 20, y = b0 >> 20;
    } else if (KIND == 1) {   // w0 of the same PA of consecutive sims
      ctr(sim + 1, pa, 0, c0, c1);
      gen<R>(c0, c1, key, b0, b1);
      x = a0 >> 20, y = b0 >> 20;
    } else if (KIND == 2) {   // the two words of one block
      x = a0 >> 20, y = a1 >> 20;
    } else {                  // low 12 bits of w1 (reliever pick field) against the next PA's w0
      ctr(sim, pa + 1, 0, c0, c1);
      gen<R>(c0, c1, key, b0, b1);
      x = a1 & 0xFFFu, y = b0 >> 20;
    }
    atomicAdd(&hist[(x << 12) | y], 1u);
  }
}

// ---- alias columns of 5 consecutive PAs ---------------------------------------------------------------------------------
template <int R>
__global__ void k_columns(uint32_t key, uint64_t n, unsigned int* hist /* [1 << 15] */) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t sim = i / (PA_SPAN / 5);
    const uint32_t pa = 5u * (uint32_t)(i % (PA_SPAN / 5));
    uint32_t cell = 0;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      uint32_t c0, c1, a0, a1;
      ctr(sim, pa + j, 0, c0, c1);
      gen<R>(c0, c1, key, a0, a1);
      cell = (cell << 3) | (a0 >> 29);
    }
    atomicAdd(&hist[cell], 1u);
  }
}

// ---- run-up lengths of w0 over the PAs of one game (element after a break is discarded: runs are independent) --------------
template <int R>
__global__ void k_runs(uint32_t key, uint32_t runs_per_thread, unsigned long long* hist /* [8]: lengths 1..7, 8+ */) {
  __shared__ unsigned int s[8];
  if (threadIdx.x < 8) s[threadIdx.x] = 0;
  __syncthreads();
  unsigned int loc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // every thread walks its own stretch of (sim, PA) counters until it has COMPLETED `runs_per_thread` runs, so no run is
  // censored (mean consumption e = 2.72 values per run; the stretch of 128 sims x 128 PAs per 4096 runs is never exhausted)
  uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * ((uint64_t)runs_per_thread * 4u);
  uint32_t prev = 0, done = 0;
  int len = 0;
  while (done < runs_per_thread) {
    uint32_t c0, c1, a0, a1;
    ctr(i / PA_SPAN, (uint32_t)(i % PA_SPAN), 0, c0, c1);
    ++i;
    gen<R>(c0, c1, key, a0, a1);
    if (len == 0 || a0 > prev) {
      ++len;
      prev = a0;
    } else {          // this value broke the run: count the run, discard the value, start afresh with the next one
      ++loc[(len > 8 ? 8 : len) - 1];
      ++done;
      len = 0;
    }
  }
  for (int j = 0; j < 8; ++j) atomicAdd(&s[j], loc[j]);
  __syncthreads();
  if (threadIdx.x < 8) atomicAdd(&hist[threadIdx.x], (unsigned long long)s[threadIdx.x]);
}

// ---- birthday spacings: one block per trial, 4096 birthdays, bitonic sorts in shared memory -----------------------------
#define BDAY_M 4096
__device__ void bitonic_sort(uint32_t* v) {
  for (int k = 2; k <= BDAY_M; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < BDAY_M; t += blockDim.x) {
        const int p = t ^ j;
        if (p > t) {
          const bool up = (t & k) == 0;
          const uint32_t a = v[t], b = v[p];
          if ((a > b) == up) {
            v[t] = b;
            v[p] = a;
          }
        }
      }
      __syncthreads();
    }
}

template <int R, int WORD>
__global__ void k_birthday(uint32_t key, unsigned long long* hist /* [32] */) {
  __shared__ uint32_t v[BDAY_M];
  __shared__ unsigned int dup;
  const uint64_t trial = blockIdx.x;
  if (threadIdx.x == 0) dup = 0;
  for (int t = threadIdx.x; t < BDAY_M; t += blockDim.x) {
    uint32_t c0, c1, a0, a1;
    ctr(trial * BDAY_M + t, (uint32_t)(trial % PA_SPAN), 0, c0, c1);   // 4096 consecutive sims at one PA number
    gen<R>(c0, c1, key, a0, a1);
    v[t] = WORD ? a1 : a0;
  }
  __syncthreads();
  bitonic_sort(v);
  uint32_t sp[BDAY_M / 1024];
  for (int q = 0, t = threadIdx.x; t < BDAY_M; t += blockDim.x, ++q) sp[q] = t ? v[t] - v[t - 1] : v[0];
  __syncthreads();
  for (int q = 0, t = threadIdx.x; t < BDAY_M; t += blockDim.x, ++q) v[t] = sp[q];
  __syncthreads();
  bitonic_sort(v);
  unsigned int mine = 0;
  for (int t = threadIdx.x; t < BDAY_M; t += blockDim.x)
    if (t && v[t] == v[t - 1]) ++mine;
  atomicAdd(&dup, mine);
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(&hist[dup > 31 ? 31 : dup], 1ull);
}

// ---- host ------------------------------------------------------------------------------------------------------------
static double chi2_uniform(const std::vector<unsigned int>& h, size_t off, size_t cells, double n) {
  const double e = n / (double)cells;
  double c = 0;
  for (size_t i = 0; i < cells; ++i) {
    const double d = (double)h[off + i] - e;
    c += d * d / e;
  }
  return c;
}

template <int R>
static void run(uint32_t key, int scale) {
  const int blocks = 148 * 8, threads = 256;
  printf("{\"rounds\": %d", R);
  // avalanche
  {
    // a whole number of grid passes: every warp runs its ballots with all 32 lanes in the loop
    const uint64_t n = (((1ull << (18 + scale)) + (uint64_t)blocks * threads - 1) / ((uint64_t)blocks * threads)) * ((uint64_t)blocks * threads);
    unsigned long long* d;
    CK(cudaMalloc(&d, 96 * 65 * 8));
    CK(cudaMemset(d, 0, 96 * 65 * 8));
    k_avalanche<R><<<blocks, threads>>>(key, n, d);
    std::vector<unsigned long long> h(96 * 65);
    CK(cudaMemcpy(h.data(), d, 96 * 65 * 8, cudaMemcpyDeviceToHost));
    CK(cudaFree(d));
    double chi = 0, zmax = 0;
    int worst = 0;
    for (int b = 0; b < 96; ++b) {
      const double nb = (double)h[b * 65 + 64];
      for (int o = 0; o < 64; ++o) {
        const double z = ((double)h[b * 65 + o] - 0.5 * nb) / std::sqrt(0.25 * nb);
        chi += z * z;
        if (std::fabs(z) > zmax) zmax = std::fabs(z), worst = b * 64 + o;
      }
    }
    printf(", \"avalanche\": {\"n\": %llu, \"chi2\": %.3f, \"dof\": %d, \"zmax\": %.3f, \"worst_in_bit\": %d, \"worst_out_bit\": %d, \"cells\": %d}",
           (unsigned long long)n, chi, 96 * 64, zmax, worst / 64, worst % 64, 96 * 64);
  }
  // hamming
  {
    const uint64_t n = 1ull << (26 + scale);
    unsigned long long* d;
    CK(cudaMalloc(&d, 130 * 8));
    CK(cudaMemset(d, 0, 130 * 8));
    k_hamming<R><<<blocks, threads>>>(key, n, d);
    unsigned long long h[130];
    CK(cudaMemcpy(h, d, 130 * 8, cudaMemcpyDeviceToHost));
    CK(cudaFree(d));
    for (int w = 0; w < 2; ++w) {
      printf(", \"hamming_%s\": {\"n\": %llu, \"hist\": [", w ? "next_pa" : "next_sim", (unsigned long long)n);
      for (int i = 0; i < 65; ++i) printf("%s%llu", i ? ", " : "", h[w * 65 + i]);
      printf("]}");
    }
  }
  // freq16
  {
    const uint64_t n = 1ull << (27 + scale);
    unsigned int* d;
    CK(cudaMalloc(&d, 4 * 65536 * 4));
    CK(cudaMemset(d, 0, 4 * 65536 * 4));
    k_freq16<R><<<blocks, threads>>>(key, n, d);
    std::vector<unsigned int> h(4 * 65536);
    CK(cudaMemcpy(h.data(), d, 4 * 65536 * 4, cudaMemcpyDeviceToHost));
    CK(cudaFree(d));
    const char* nm[4] = {"w0_hi", "w0_lo", "w1_hi", "w1_lo"};
    for (int f = 0; f < 4; ++f)
      printf(", \"freq16_%s\": {\"n\": %llu, \"chi2\": %.3f, \"dof\": 65535}", nm[f], (unsigned long long)n, chi2_uniform(h, f * 65536, 65536, (double)n));
  }
  // serial
  {
    const uint64_t n = 1ull << (27 + scale);
    unsigned int* d;
    CK(cudaMalloc(&d, (1u << 24) * 4));
    std::vector<unsigned int> h(1u << 24);
    const char* nm[4] = {"w0_next_pa", "w0_next_sim", "w0_w1", "w1lo_next_w0"};
    for (int kind = 0; kind < 4; ++kind) {
      CK(cudaMemset(d, 0, (1u << 24) * 4));
      if (kind == 0) k_serial<R, 0><<<blocks, threads>>>(key, n, d);
      if (kind == 1) k_serial<R, 1><<<blocks, threads>>>(key, n, d);
      if (kind == 2) k_serial<R, 2><<<blocks, threads>>>(key, n, d);
      if (kind == 3) k_serial<R, 3><<<blocks, threads>>>(key, n, d);
      CK(cudaMemcpy(h.data(), d, (1u << 24) * 4, cudaMemcpyDeviceToHost));
      printf(", \"serial_%s\": {\"n\": %llu, \"chi2\": %.3f, \"dof\": %u}", nm[kind], (unsigned long long)n, chi2_uniform(h, 0, 1u << 24, (double)n),
             (1u << 24) - 1);
    }
    CK(cudaFree(d));
  }
  // columns
  {
    const uint64_t n = 1ull << (26 + scale);
    unsigned int* d;
    CK(cudaMalloc(&d, (1u << 15) * 4));
    CK(cudaMemset(d, 0, (1u << 15) * 4));
    k_columns<R><<<blocks, threads>>>(key, n, d);
    std::vector<unsigned int> h(1u << 15);
    CK(cudaMemcpy(h.data(), d, (1u << 15) * 4, cudaMemcpyDeviceToHost));
    CK(cudaFree(d));
    printf(", \"columns5\": {\"n\": %llu, \"chi2\": %.3f, \"dof\": %u}", (unsigned long long)n, chi2_uniform(h, 0, 1u << 15, (double)n), (1u << 15) - 1);
  }
  // runs
  {
    const uint32_t n = 1u << (8 + scale);   // runs per thread
    unsigned long long* d;
    CK(cudaMalloc(&d, 8 * 8));
    CK(cudaMemset(d, 0, 8 * 8));
    k_runs<R><<<blocks, threads>>>(key, n, d);
    unsigned long long h[8];
    CK(cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost));
    CK(cudaFree(d));
    printf(", \"runs_up\": {\"runs\": %llu, \"hist\": [", (unsigned long long)n * blocks * threads);
    for (int i = 0; i < 8; ++i) printf("%s%llu", i ? ", " : "", h[i]);
    printf("]}");
  }
  // birthday spacings
  {
    const unsigned trials = 1u << (11 + scale);
    unsigned long long* d;
    CK(cudaMalloc(&d, 32 * 8));
    for (int word = 0; word < 2; ++word) {
      CK(cudaMemset(d, 0, 32 * 8));
      if (word == 0) k_birthday<R, 0><<<trials, 1024>>>(key, d);
      else k_birthday<R, 1><<<trials, 1024>>>(key, d);
      unsigned long long h[32];
      CK(cudaMemcpy(h, d, 32 * 8, cudaMemcpyDeviceToHost));
      printf(", \"birthday_w%d\": {\"trials\": %u, \"m\": %d, \"lambda\": 4.0, \"hist\": [", word, trials, BDAY_M);
      for (int i = 0; i < 32; ++i) printf("%s%llu", i ? ", " : "", h[i]);
      printf("]}");
    }
    CK(cudaFree(d));
  }
  CK(cudaDeviceSynchronize());
  printf("}\n");
  fflush(stdout);
}

int main(int argc, char** argv) {
  // usage: rng_battery [scale 0..6] [seed] [rounds ...]   — sample sizes are (base << scale)
  const int scale = argc > 1 ? atoi(argv[1]) : 4;
  const uint64_t seed = argc > 2 ? strtoull(argv[2], nullptr, 0) : 20250404ull;
  const uint32_t key = philox2_keys_from_seed(seed).k[0];   // the product's key schedule: round keys key + r * W0
  std::vector<int> rounds;
  for (int i = 3; i < argc; ++i) rounds.push_back(atoi(argv[i]));
  if (rounds.empty()) rounds = {3, 4, 5, 6, 7, 8, 10};
  for (int r : rounds) {
    switch (r) {
      case 2: run<2>(key, scale); break;
      case 3: run<3>(key, scale); break;
      case 4: run<4>(key, scale); break;
      case 5: run<5>(key, scale); break;
      case 6: run<6>(key, scale); break;
      case 7: run<7>(key, scale); break;
      case 8: run<8>(key, scale); break;
      case 9: run<9>(key, scale); break;
      case 10: run<10>(key, scale); break;
      default: fprintf(stderr, "rounds %d not instantiated\n", r); return 2;
    }
  }
  return 0;
}
