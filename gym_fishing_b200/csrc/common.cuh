// Small shared helpers: streaming (evict-first) vector loads/stores, the launch counter and SM count.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fishing {

// Data that is touched once per step (actions, injected noise, reward/done slices of a rollout buffer) should not
// displace the L2-resident state: use the streaming cache operators.
template <typename T>
__device__ __forceinline__ T ld_stream(const T* p) {
  return __ldcs(p);
}
template <typename T>
__device__ __forceinline__ void st_stream(T* p, const T& v) {
  __stcs(p, v);
}
// aggregate pairs that __ldcs has no overload for: load through the matching builtin vector type
template <typename T, typename V>
__device__ __forceinline__ T ld_stream_as(const T* p) {
  static_assert(sizeof(T) == sizeof(V), "size mismatch");
  V v = __ldcs(reinterpret_cast<const V*>(p));
  return *reinterpret_cast<T*>(&v);
}

void count_launch();
int sm_count();

}  // namespace fishing
