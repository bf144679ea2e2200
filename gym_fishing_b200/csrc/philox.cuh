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
#pragma once
#include <stdint.h>

namespace fishing {

struct Normal2 {
  double z0, z1;
};

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&w)[4]) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
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

// 52 random bits -> (k + 0.5) * 2^-52 without an integer->double conversion: place the bits in the mantissa of a
// number in [1, 2) and subtract (1 - 2^-53); the subtraction is exact.
__device__ __forceinline__ double u52_to_open_unit(uint32_t lo, uint32_t hi) {
  const double d = __hiloint2double(0x3FF00000 | (hi & 0xFFFFF), lo);
  return __dsub_rn(d, 0x1.fffffffffffffp-1);
}

__device__ __forceinline__ Normal2 normal_pair(uint64_t seed, uint64_t id, uint64_t t, uint32_t stream) {
  uint32_t w[4];
  philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), (uint32_t)t, (stream << 24) | ((uint32_t)(t >> 32) & 0xFFFFFFu),
                (uint32_t)seed, (uint32_t)(seed >> 32), w);
  const double u1 = u52_to_open_unit(w[0], w[1]);
  const double u2 = u52_to_open_unit(w[2], w[3]);
  const double R = __dsqrt_rn(__dmul_rn(-2.0, log(u1)));
  double sn, cs;
  sincospi(__dmul_rn(2.0, u2), &sn, &cs);
  Normal2 out;
  out.z0 = __dmul_rn(R, cs);
  out.z1 = __dmul_rn(R, sn);
  return out;
}

}  // namespace fishing
