// Counter-based Philox RNGs (Salmon, Moraes, Dror, Shaw — SC'11), written for sm_100a.
//
// The hot loop uses Philox2x32-10: one 32x32->64 multiply (IMAD.WIDE.U32, the scarce resource on
// the fmaheavy pipe: 4 issue cycles per warp) and one three-input xor (LOP3) per round, i.e. half
// the multiplies per random bit of Philox4x32-10.  The 32-bit key is derived from the 64-bit
// launch seed on the host (with Philox4x32-10), so the ten round keys are launch constants that
// reach the kernel through the parameter constant bank — LOP3 reads them as c[0x0][..] operands:
// no per-thread key registers and no key-bump adds.
//
// Counter layout (what makes sim i of matchup m reproducible on any launch geometry / GPU count):
//   c0 = sim index bits 31..0
//   c1 = plate-appearance number within the game [12:0] | reserved 0 [13] | matchup id [25:14] |
//        sim index bits 37..32 [31:26]
// One block (w0, w1) per plate appearance: w0 = outcome (alias draw), w1 = base-running draws;
// the w1 of a half's first PA (bases empty: unused by the transition) is that half's
// end-of-inning draw, the low 16 bits of the w1 of its last PA (an out) pick the next reliever.
#pragma once
#include <stdint.h>

#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u
#define PHILOX2_M 0xD256D193u
// Rounds of the generator behind the native law.  10 = Philox2x32-10 as published (the product).  Fewer rounds are an
// EXPERIMENT switch only (what the generator costs, measured): the oracle twin must be built with the same
// value (oracle/Makefile: NATIVE_PHILOX_ROUNDS), and nothing but 10 is backed by the Random123 known-answer vectors.
#ifndef NATIVE_PHILOX_ROUNDS
#define NATIVE_PHILOX_ROUNDS 10
#endif

#define RNG_KIND_HALF_END (1u << 13)
#define RNG_MATCHUP_SHIFT 14
#define RNG_SIMHI_SHIFT 26
#define RNG_MAX_MATCHUPS 4096u          /* 12 bits */
#define RNG_MAX_SIM_INDEX (1ull << 38)  /* 32 + 6 bits */

struct Philox2Keys {
  uint32_t k[10];
};

static inline void philox4x32_10_host(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)PHILOX_M0 * c0, p1 = (uint64_t)PHILOX_M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    k0 += PHILOX_W0; k1 += PHILOX_W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 64-bit seed -> the 32-bit Philox2x32 key of this launch family (word 0 of Philox4x32-10 over a
// fixed domain-separation counter) -> its ten round keys
static inline Philox2Keys philox2_keys_from_seed(uint64_t seed) {
  const uint32_t ctr[4] = {0x6d6c6273u /* "mlbs" */, 0x696d3278u /* "im2x" */, 0u, 0u};
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t out[4];
  philox4x32_10_host(ctr, key, out);
  Philox2Keys K;
  uint32_t k = out[0];
  for (int r = 0; r < 10; ++r) {
    K.k[r] = k;
    k += PHILOX_W0;
  }
  return K;
}

#ifdef __CUDACC__
__device__ __forceinline__ void philox2x32_native(uint32_t c0, uint32_t c1, const Philox2Keys& K, uint32_t& o0, uint32_t& o1) {
#pragma unroll
  for (int r = 0; r < NATIVE_PHILOX_ROUNDS; ++r) {
    const uint64_t p = (uint64_t)PHILOX2_M * c0;
    c0 = (uint32_t)(p >> 32) ^ K.k[r] ^ c1;
    c1 = (uint32_t)p;
  }
  o0 = c0;
  o1 = c1;
}
#endif
