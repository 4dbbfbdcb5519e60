// MaxPool2d(kernel 3, stride 3, floor) on channels-last fp32 activations, for sm_100a.
//
// Reference: the nn.MaxPool2d(kernel_size=3, stride=3) after every Conv64F block
// (libfewshot_core/model/backbone/conv_four.py:65,71,77,84).  ATen's channels-last max-pool spends
// 0.16 ms per launch on B200 for a [800,64,42,52] activation (profiles/r01_launches_after_conv1.csv):
// one thread per output element, scalar loads.  Here a thread owns 4 consecutive channels of one pooled
// pixel: nine 128-bit loads (consecutive threads -> consecutive channels: fully coalesced), one 128-bit
// store.  HBM-bound: 4*C*(H*W + (H/3)*(W/3)) bytes per image.
#include <cuda_bf16.h>

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

// The same tail for the separately stated bf16 trunk (ResNet.precision = "bf16"): a and b are bf16 channels-last
// convolution outputs, a thread owns 8 consecutive channels of one pooled pixel (one 128-bit load per operand and window
// position), sum, max, bias and activation in fp32, one rounding at the store (bf16) or none (fp32 output of the last
// block, which feeds the fp32 heads).
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

template <int K, bool OUT_F32>
__global__ void __launch_bounds__(256)
add_bias_act_pool_nhwc_bf16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, const float4* __restrict__ bias,
                                   int64_t total, int H, int W, int C8, int PH, int PW, float slope, void* out_) {
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C8);
    int64_t r = idx / C8;
    const int pw = static_cast<int>(r % PW);
    r /= PW;
    const int ph = static_cast<int>(r % PH);
    const int64_t n = r / PH;
    const int64_t base = ((n * H + K * ph) * W + K * pw) * C8 + c;
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
#pragma unroll
    for (int dy = 0; dy < K; ++dy) {
#pragma unroll
      for (int dx = 0; dx < K; ++dx) {
        const int64_t o = base + (static_cast<int64_t>(dy) * W + dx) * C8;
        float v[8];
        unpack8(a[o], v);  // plain loads: `out` may alias `a` when K == 1 and the output is bf16
        if (b != nullptr) {
          float w[8];
          unpack8(__ldg(b + o), w);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] += w[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], v[i]);
      }
    }
    if (bias != nullptr) {
      const float4 b0 = __ldg(bias + 2 * c), b1 = __ldg(bias + 2 * c + 1);
      m[0] += b0.x; m[1] += b0.y; m[2] += b0.z; m[3] += b0.w;
      m[4] += b1.x; m[5] += b1.y; m[6] += b1.z; m[7] += b1.w;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = m[i] > 0.f ? m[i] : m[i] * slope;
    if (OUT_F32) {
      float4* o = static_cast<float4*>(out_) + 2 * idx;
      o[0] = make_float4(m[0], m[1], m[2], m[3]);
      o[1] = make_float4(m[4], m[5], m[6], m[7]);
    } else {
      const __nv_bfloat162 p0 = __floats2bfloat162_rn(m[0], m[1]), p1 = __floats2bfloat162_rn(m[2], m[3]);
      const __nv_bfloat162 p2 = __floats2bfloat162_rn(m[4], m[5]), p3 = __floats2bfloat162_rn(m[6], m[7]);
      uint4 v;
      v.x = *reinterpret_cast<const uint32_t*>(&p0); v.y = *reinterpret_cast<const uint32_t*>(&p1);
      v.z = *reinterpret_cast<const uint32_t*>(&p2); v.w = *reinterpret_cast<const uint32_t*>(&p3);
      static_cast<uint4*>(out_)[idx] = v;
    }
  }
}

template <int K>
void launch_tail_bf16(unsigned g, cudaStream_t st, const uint4* a, const uint4* b, const float4* bias, int64_t total, int H,
                      int W, int C8, int PH, int PW, float slope, void* out, bool out_f32) {
  if (out_f32) add_bias_act_pool_nhwc_bf16_kernel<K, true><<<g, 256, 0, st>>>(a, b, bias, total, H, W, C8, PH, PW, slope, out);
  else add_bias_act_pool_nhwc_bf16_kernel<K, false><<<g, 256, 0, st>>>(a, b, bias, total, H, W, C8, PH, PW, slope, out);
}

}  // namespace
}  // namespace afs

extern "C" int afs_add_bias_act_pool_nhwc_bf16_fwd(const void* a, const void* b, const float* bias, int32_t N, int32_t H,
                                                   int32_t W, int32_t C, float negative_slope, int32_t k, void* out,
                                                   int32_t out_f32, afs_stream_t stream_) {
  using namespace afs;
  if (a == nullptr || out == nullptr || N < 0 || k < 1 || k > 3 || H < k || W < k || C < 8 || (C & 7) != 0 ||
      negative_slope < 0.f)
    return AFS_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(bias) |
       reinterpret_cast<uintptr_t>(out)) & 15)
    return AFS_ERR_INVALID_ARG;
  if (out == a && (k != 1 || out_f32)) return AFS_ERR_INVALID_ARG;  // in place only without pooling, same dtype
  if (N == 0) return AFS_OK;
  const int PH = H / k, PW = W / k, C8 = C / 8;
  const int64_t total = static_cast<int64_t>(N) * PH * PW * C8;
  int64_t blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 32) blocks = kNumSMs * 32;
  const uint4* a8 = static_cast<const uint4*>(a);
  const uint4* b8 = static_cast<const uint4*>(b);
  const float4* s4 = reinterpret_cast<const float4*>(bias);
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const unsigned g = static_cast<unsigned>(blocks);
  if (k == 1) launch_tail_bf16<1>(g, st, a8, b8, s4, total, H, W, C8, PH, PW, negative_slope, out, out_f32 != 0);
  else if (k == 2) launch_tail_bf16<2>(g, st, a8, b8, s4, total, H, W, C8, PH, PW, negative_slope, out, out_f32 != 0);
  else launch_tail_bf16<3>(g, st, a8, b8, s4, total, H, W, C8, PH, PW, negative_slope, out, out_f32 != 0);
  AFS_LAUNCH_CHECK();
  return AFS_OK;
}

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
