// Device helpers: streaming loads, warp / block reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nsb {

// 128-bit streaming load through the read-only path without allocating in L1 (the basis is
// read once per kernel; keeping it out of L1 leaves the cache for h / D / index data).
__device__ __forceinline__ double2 ld_stream(const double2 *p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

__device__ __forceinline__ double ld_stream1(const double *p) {
  double r;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}

__device__ __forceinline__ double warp_reduce_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over a block of NT threads; result valid in thread 0 (fixed order -> deterministic).
template <int NT>
__device__ __forceinline__ double block_reduce_sum(double v) {
  __shared__ double red_[NT / 32];
  v = warp_reduce_sum(v);
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red_[warp] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NT / 32; ++i) s += red_[i];
  }
  return s;
}

}  // namespace nsb
