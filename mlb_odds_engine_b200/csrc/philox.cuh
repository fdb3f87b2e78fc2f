// Philox4x32-10 counter-based RNG (Salmon, Moraes, Dror, Shaw — SC'11), written for sm_100a.
//
// The key is the launch seed, so the ten round keys are launch constants: the host expands them
// once (philox_expand_key) and they reach the kernel through the parameter constant bank, where
// LOP3 reads them as c[0x0][..] operands — no per-thread key registers and no key-bump adds.
// A round is then 2x IMAD.WIDE.U32 + 2x LOP3 (three-input xor).
#pragma once
#include <stdint.h>

#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

struct PhiloxKeys {
  uint32_t k0[10];
  uint32_t k1[10];
};

static inline PhiloxKeys philox_expand_key(uint64_t seed) {
  PhiloxKeys k;
  uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    k.k0[r] = a;
    k.k1[r] = b;
    a += PHILOX_W0;
    b += PHILOX_W1;
  }
  return k;
}

#ifdef __CUDACC__
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& K,
                                              uint32_t& o0, uint32_t& o1, uint32_t& o2, uint32_t& o3) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
    const uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.k0[r];
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.k1[r];
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
  }
  o0 = c0;
  o1 = c1;
  o2 = c2;
  o3 = c3;
}
#endif
