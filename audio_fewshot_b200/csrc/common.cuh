// Shared helpers for the sm_100a kernels behind include/afs_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "afs_b200.h"

namespace afs {

extern thread_local int g_last_cuda_error;
void count_launch();

inline int cuda_fail(cudaError_t e) {
  g_last_cuda_error = static_cast<int>(e);
  return AFS_ERR_CUDA;
}

#define AFS_CUDA_TRY(expr)                            \
  do {                                                \
    cudaError_t _e = (expr);                          \
    if (_e != cudaSuccess) return ::afs::cuda_fail(_e); \
  } while (0)

// One per kernel launch: counts it (afs_launch_count) and surfaces launch-configuration errors.
#define AFS_LAUNCH_CHECK()                      \
  do {                                          \
    ::afs::count_launch();                      \
    AFS_CUDA_TRY(cudaGetLastError());           \
  } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Streaming 128-bit load: data that is read once bypasses L1 allocation.
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float2 ldg_stream2(const float* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];"
               : "=f"(r.x), "=f"(r.y)
               : "l"(p));
  return r;
}

}  // namespace afs
