// Micro-benchmark: scalar FFMA/FADD vs packed FFMA2/FADD2 (sm_100a) at equal flops.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096;
__global__ void k_scalar(float* out, float a, float b) {
  float x[16];
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float* out, float a, float b) {
  float2 x[8];
  for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __ffma2_rn(x[i], a2, b2);
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_add_scalar(float* out, float b) {
  float x[16];
  for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = x[i] + b;
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_add_packed(float* out, float b) {
  float2 x[8];
  for (int i = 0; i < 8; ++i) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
  const float2 b2 = make_float2(b, -b);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __fadd2_rn(x[i], b2);
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); for (int i = 0; i < 5; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / 5;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  const int grid = 148 * 8, block = 256;
  const double flops_fma = 2.0 * 16 * ITERS * grid * block, flops_add = 1.0 * 16 * ITERS * grid * block;
  float t;
  t = timeit([&] { k_scalar<<<grid, block>>>(out, 1.0001f, 0.5f); });     printf("FFMA   %.3f ms  %.1f TFLOP/s\n", t, flops_fma / t / 1e9);
  t = timeit([&] { k_packed<<<grid, block>>>(out, 1.0001f, 0.5f); });     printf("FFMA2  %.3f ms  %.1f TFLOP/s\n", t, flops_fma / t / 1e9);
  t = timeit([&] { k_add_scalar<<<grid, block>>>(out, 0.5f); });          printf("FADD   %.3f ms  %.1f Tadd/s\n", t, flops_add / t / 1e9);
  t = timeit([&] { k_add_packed<<<grid, block>>>(out, 0.5f); });          printf("FADD2  %.3f ms  %.1f Tadd/s\n", t, flops_add / t / 1e9);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
