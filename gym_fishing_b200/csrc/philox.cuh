// Counter-based Gaussian noise for the fishing engine: Philox4x32-10 (Salmon et al., SC'11) + Box-Muller in
// binary64.  This replaces the reference's global numpy MT19937 draw np.random.normal(0, 1)
// (gym_fishing/envs/base_fishing_env.py:130, fishing_tipping_env.py:31, fishing_model_error.py:37-43): the
// stream is a pure function of (seed, noise id, step, stream id), so results do not depend on how environments
// are partitioned over threads, launches or GPUs.  oracle/philox_ref.py restates the same spec in numpy.
//
//   key      = (seed & 0xffffffff, seed >> 32)
//   counter  = (id & 0xffffffff, id >> 32, t & 0xffffffff, (stream << 24) | ((t >> 32) & 0xffffff))
//   (w0..w3) = philox4x32_10(counter, key)
//   u1       = ((w1 & 0xfffff) * 2^32 + w0 + 0.5) * 2^-52          in (0, 1), exact in binary64
//   u2       = ((w3 & 0xfffff) * 2^32 + w2 + 0.5) * 2^-52
//   R        = sqrt(-2 ln u1);   z0 = R cos(2 pi u2);   z1 = R sin(2 pi u2)
//
// Process noise (stream 0) uses id = env_id >> 1 and hands z0 to the even and z1 to the odd environment of the
// pair, so one Philox block + one Box-Muller serves two environments.
//
// The Box-Muller itself lives in normal.cuh (round 2: table-assisted log and sin/cos, 28 FP64 instructions per pair of
// draws; the round-1 evaluation is kept in normal_v1.cuh for A/B timing only).
#pragma once
#include <stdint.h>

#ifndef FISHING_NORMAL_V1
#include "normal.cuh"
#endif

namespace fishing {

#ifdef FISHING_NORMAL_V1
struct Normal2 {
  double z0, z1;
};
#endif

// A/B timing builds only (tools/ab_build.py): fewer rounds, or no rounds at all, to MEASURE what the multiplies cost.  The
// product is always built with the full ten rounds (tests/test_philox_ref.py: Random123 known answers).
#ifndef FISHING_PHILOX_ROUNDS
#define FISHING_PHILOX_ROUNDS 10
#endif

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&w)[4]) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < FISHING_PHILOX_ROUNDS; ++i) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += W0;
    k1 += W1;
  }
  w[0] = c0;
  w[1] = c1;
  w[2] = c2;
  w[3] = c3;
}

// The ten round keys (k + i*W) depend only on the seed: the host precomputes them into the kernel parameters, where the
// xor of each round can take them straight from the constant bank (saves the 18 key additions per Philox block).
struct PhiloxKeys {
  uint32_t k0[10], k1[10];
  uint32_t m[2];   // the two Philox multipliers again (only read by the FISHING_AB_SPLIT_MUL timing build)
};
inline PhiloxKeys philox_round_keys(uint64_t seed) {
  PhiloxKeys k;
  uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
  for (int i = 0; i < 10; ++i) {
    k.k0[i] = a;
    k.k1[i] = b;
    a += 0x9E3779B9u;
    b += 0xBB67AE85u;
  }
  k.m[0] = 0xD2511F53u;
  k.m[1] = 0xCD9E8D57u;
  return k;
}

__device__ __forceinline__ void philox4x32_10_keys(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& k,
                                                   uint32_t (&w)[4]) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int i = 0; i < FISHING_PHILOX_ROUNDS; ++i) {
#ifdef FISHING_AB_SPLIT_MUL
    // timing build: the low halves multiply by a copy of the constant the compiler cannot see, so that ptxas does not
    // fuse mul.hi + mul.lo into one IMAD.WIDE
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = k.m[0] * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = k.m[1] * c2;
#else
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
#endif
    c0 = hi1 ^ c1 ^ k.k0[i];
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k.k1[i];
    c3 = lo0;
  }
  w[0] = c0;
  w[1] = c1;
  w[2] = c2;
  w[3] = c3;
}

#ifdef FISHING_NORMAL_V1
// A/B build: the round-1 evaluation (normal_v1.cuh).  The table accessors exist but are never read.
#include "normal_v1.cuh"
struct GlobalTabs {};
struct SharedTabs {};
template <class Tabs>
__device__ __forceinline__ double exp_tab(double x, const Tabs&) { return exp(x); }
constexpr int NORMAL_TABLE_ENTRIES = 1;
template <int BS>
__device__ __forceinline__ SharedTabs stage_normal_tables(double2*) { return SharedTabs(); }
template <class Tabs>
__device__ __forceinline__ Normal2 normal_from_words(const uint32_t (&w)[4], const Tabs&) { return normal_from_words_v1(w); }
#endif

// Uniform integer in [0, n) from the first Philox word (fishing-v11 model draw, np.random.choice(models)):
// floor(w0 * n / 2^32).
__device__ __forceinline__ uint32_t uniform_index(uint64_t seed, uint64_t id, uint64_t t, uint32_t stream, uint32_t n) {
  uint32_t w[4];
  philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), (uint32_t)t, (stream << 24) | ((uint32_t)(t >> 32) & 0xFFFFFFu),
                (uint32_t)seed, (uint32_t)(seed >> 32), w);
  return __umulhi(w[0], n);
}

// the pair of draws of (id, t, stream), round keys precomputed on the host; `tb` says where the Box-Muller tables are read
// (GlobalTabs: read-only global loads, SharedTabs: the block's staged copy) -- the result does not depend on it
template <class Tabs = GlobalTabs>
__device__ __forceinline__ Normal2 normal_pair_keys(const PhiloxKeys& k, uint64_t id, uint64_t t, uint32_t stream,
                                                    const Tabs& tb = Tabs()) {
  uint32_t w[4];
  philox4x32_10_keys((uint32_t)id, (uint32_t)(id >> 32), (uint32_t)t, (stream << 24) | ((uint32_t)(t >> 32) & 0xFFFFFFu), k, w);
  return normal_from_words(w, tb);
}

template <class Tabs = GlobalTabs>
__device__ __forceinline__ Normal2 normal_pair(uint64_t seed, uint64_t id, uint64_t t, uint32_t stream, const Tabs& tb = Tabs()) {
  uint32_t w[4];
  philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), (uint32_t)t, (stream << 24) | ((uint32_t)(t >> 32) & 0xFFFFFFu),
                (uint32_t)seed, (uint32_t)(seed >> 32), w);
  return normal_from_words(w, tb);
}

}  // namespace fishing
