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

// out = MaxPool_k( act( a + b + bias[c] ) ), k in {1,2,3}, floor mode, channels-last.  The tail of a ResNet-12
// BasicBlock in eval mode (libfewshot_core/model/backbone/resnet_12.py:79-101: bn3 -> += residual -> LeakyReLU ->
// MaxPool2d(stride)) with both BatchNorms folded into the convolutions: one read of the two convolution outputs,
// one write of the pooled map, instead of five elementwise passes over the un-pooled activation.  With b == null
// and k == 1 it is the in-place bias + activation after a folded convolution.  The activation is monotone
// (slope >= 0), so it is applied once, after the max.
template <int K>
__global__ void __launch_bounds__(256)
add_bias_act_pool_nhwc_kernel(const float4* __restrict__ a, const float4* __restrict__ b, const float4* __restrict__ bias,
                              int64_t total, int H, int W, int C4, int PH, int PW, float slope, float4* out) {
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C4);
    int64_t r = idx / C4;
    const int pw = static_cast<int>(r % PW);
    r /= PW;
    const int ph = static_cast<int>(r % PH);
    const int64_t n = r / PH;
    const int64_t base = ((n * H + K * ph) * W + K * pw) * C4 + c;
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int dy = 0; dy < K; ++dy) {
#pragma unroll
      for (int dx = 0; dx < K; ++dx) {
        const int64_t o = base + (static_cast<int64_t>(dy) * W + dx) * C4;
        float4 v = a[o];  // plain loads: `out` may alias `a` when K == 1
        if (b != nullptr) {
          const float4 w = __ldg(b + o);
          v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        }
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
      }
    }
    if (bias != nullptr) {
      const float4 bb = __ldg(bias + c);
      m.x += bb.x; m.y += bb.y; m.z += bb.z; m.w += bb.w;
    }
    m.x = m.x > 0.f ? m.x : m.x * slope;
    m.y = m.y > 0.f ? m.y : m.y * slope;
    m.z = m.z > 0.f ? m.z : m.z * slope;
    m.w = m.w > 0.f ? m.w : m.w * slope;
    out[idx] = m;
  }
}

}  // namespace
}  // namespace afs

extern "C" int afs_add_bias_act_pool_nhwc_fwd(const float* a, const float* b, const float* bias, int32_t N, int32_t H,
                                              int32_t W, int32_t C, float negative_slope, int32_t k, float* out,
                                              afs_stream_t stream_) {
  using namespace afs;
  if (a == nullptr || out == nullptr || N < 0 || k < 1 || k > 3 || H < k || W < k || C < 4 || (C & 3) != 0 ||
      negative_slope < 0.f)
    return AFS_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(bias) |
       reinterpret_cast<uintptr_t>(out)) & 15)
    return AFS_ERR_INVALID_ARG;
  if (out == a && k != 1) return AFS_ERR_INVALID_ARG;  // in place only without pooling
  if (N == 0) return AFS_OK;
  const int PH = H / k, PW = W / k, C4 = C / 4;
  const int64_t total = static_cast<int64_t>(N) * PH * PW * C4;
  int64_t blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 32) blocks = kNumSMs * 32;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  const float4* s4 = reinterpret_cast<const float4*>(bias);
  float4* o4 = reinterpret_cast<float4*>(out);
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const unsigned g = static_cast<unsigned>(blocks);
  if (k == 1) add_bias_act_pool_nhwc_kernel<1><<<g, 256, 0, st>>>(a4, b4, s4, total, H, W, C4, PH, PW, negative_slope, o4);
  else if (k == 2) add_bias_act_pool_nhwc_kernel<2><<<g, 256, 0, st>>>(a4, b4, s4, total, H, W, C4, PH, PW, negative_slope, o4);
  else add_bias_act_pool_nhwc_kernel<3><<<g, 256, 0, st>>>(a4, b4, s4, total, H, W, C4, PH, PW, negative_slope, o4);
  AFS_LAUNCH_CHECK();
  return AFS_OK;
}

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
