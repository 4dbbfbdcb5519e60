// MaxPool2d(kernel 3, stride 3, floor) on channels-last fp32 activations, for sm_100a.
//
// Reference: the nn.MaxPool2d(kernel_size=3, stride=3) after every Conv64F block
// (libfewshot_core/model/backbone/conv_four.py:65,71,77,84).  ATen's channels-last max-pool spends
// 0.16 ms per launch on B200 for a [800,64,42,52] activation (profiles/r01_launches_after_conv1.csv):
// one thread per output element, scalar loads.  Here a thread owns 4 consecutive channels of one pooled
// pixel: nine 128-bit loads (consecutive threads -> consecutive channels: fully coalesced), one 128-bit
// store.  HBM-bound: 4*C*(H*W + (H/3)*(W/3)) bytes per image.
#include "common.cuh"

namespace afs {
namespace {

__global__ void __launch_bounds__(256)
maxpool3_nhwc_kernel(const float4* __restrict__ x, int64_t total, int H, int W, int C4, int PH, int PW,
                     float4* __restrict__ out) {
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C4);
    int64_t r = idx / C4;
    const int pw = static_cast<int>(r % PW);
    r /= PW;
    const int ph = static_cast<int>(r % PH);
    const int64_t n = r / PH;
    const float4* base = x + ((n * H + 3 * ph) * W + 3 * pw) * C4 + c;
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const float4 v = __ldg(base + (static_cast<int64_t>(dy) * W + dx) * C4);
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    }
    out[idx] = m;
  }
}

}  // namespace
}  // namespace afs

extern "C" int afs_maxpool3_nhwc_fwd(const float* x, int32_t N, int32_t H, int32_t W, int32_t C, float* out,
                                     afs_stream_t stream_) {
  using namespace afs;
  if (x == nullptr || out == nullptr || N < 0 || H < 3 || W < 3 || C < 4 || (C & 3) != 0) return AFS_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) return AFS_ERR_INVALID_ARG;
  if (N == 0) return AFS_OK;
  const int PH = H / 3, PW = W / 3, C4 = C / 4;
  const int64_t total = static_cast<int64_t>(N) * PH * PW * C4;
  int64_t blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 32) blocks = kNumSMs * 32;
  maxpool3_nhwc_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      reinterpret_cast<const float4*>(x), total, H, W, C4, PH, PW, reinterpret_cast<float4*>(out));
  AFS_LAUNCH_CHECK();
  return AFS_OK;
}
