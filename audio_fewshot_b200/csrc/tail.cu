// Tail of the Conv64F inference path for sm_100a: the last 3x3/3 max-pool fused with the 64 -> 1600 linear layer.
//
// Reference: libfewshot_core/model/backbone/conv_four.py:84 (layer4_pool) and :89-92, 120-123 (logits = Dropout ->
// BatchNorm1d -> Linear, applied to out4.view(N, -1)); in eval mode BatchNorm1d folds into the linear layer.  On the
// audio shape the pooled map is [64, 1, 1], so the flatten is the 64 pooled channels.  The library sequence was an
// ATen max-pool launch, a SIMT cuBLAS sgemm and a cublasLt epilogue kernel: 56 us per 3 200 clips on B200
// (profiles/r02_launches_final.csv); this kernel does the same arithmetic (fp32 FMA, channels summed in ascending
// order) in one launch.
//
// CTA = 64 clips x 128 outputs, 256 threads, thread tile 4 clips x 8 outputs: per channel one 128-bit read of the
// pooled values ([channel][clip] in shared memory) and two of the weights ([channel][output]) feed 32 FMAs.  The pooled
// values are recomputed by each of the 13 column CTAs of a clip tile (9 x 64 floats per clip out of L2).
#include "common.cuh"

namespace afs {
namespace {

constexpr int kTM = 64;    // clips per CTA
constexpr int kTN = 128;   // outputs per CTA
constexpr int kTC = 64;    // channels
constexpr int kTailThreads = 256;

__global__ void __launch_bounds__(kTailThreads)
pool3_linear_kernel(const float* __restrict__ x, int N, int H, int W, const float* __restrict__ wl,
                    const float* __restrict__ bl, int J, float* __restrict__ out) {
  extern __shared__ __align__(16) float s_tail[];
  float (*s_v)[kTM] = reinterpret_cast<float (*)[kTM]>(s_tail);                        // pooled activations [channel][clip]
  float (*s_w)[kTN + 4] = reinterpret_cast<float (*)[kTN + 4]>(s_tail + kTC * kTM);    // weights [channel][output] (+4: conflict-free transposed fill)
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * kTM;
  const int j0 = blockIdx.y * kTN;

  // pooled values: thread = (clip, 16-channel group); 9 positions x four 128-bit loads
  {
    const int clip = tid >> 2, cg = tid & 3;
    const int n = n0 + clip;
    float4 m[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) m[i] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    if (n < N) {
      const float4* base = reinterpret_cast<const float4*>(x + static_cast<int64_t>(n) * H * W * kTC) + cg * 4;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 v = __ldg(base + (dy * W + dx) * (kTC / 4) + i);
            m[i].x = fmaxf(m[i].x, v.x); m[i].y = fmaxf(m[i].y, v.y); m[i].z = fmaxf(m[i].z, v.z); m[i].w = fmaxf(m[i].w, v.w);
          }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) m[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = cg * 16 + 4 * i;
      s_v[c + 0][clip] = m[i].x; s_v[c + 1][clip] = m[i].y; s_v[c + 2][clip] = m[i].z; s_v[c + 3][clip] = m[i].w;
    }
  }
  // weights: wl [J, 64] row-major -> s_w[c][j]
  for (int i = tid; i < kTN * (kTC / 4); i += kTailThreads) {
    const int j = i >> 4, c4 = i & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j0 + j < J) v = __ldg(reinterpret_cast<const float4*>(wl + static_cast<int64_t>(j0 + j) * kTC) + c4);
    s_w[4 * c4 + 0][j] = v.x; s_w[4 * c4 + 1][j] = v.y; s_w[4 * c4 + 2][j] = v.z; s_w[4 * c4 + 3][j] = v.w;
  }
  __syncthreads();

  const int tj = tid & 15, tn = tid >> 4;  // outputs j0 + 8 tj .. + 7, clips n0 + 4 tn .. + 3
  float acc[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
#pragma unroll 8
  for (int c = 0; c < kTC; ++c) {
    const float4 v = *reinterpret_cast<const float4*>(&s_v[c][4 * tn]);
    const float4 w0 = *reinterpret_cast<const float4*>(&s_w[c][8 * tj]);
    const float4 w1 = *reinterpret_cast<const float4*>(&s_w[c][8 * tj + 4]);
    const float vv[4] = {v.x, v.y, v.z, v.w};
    const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(vv[a], ww[b], acc[a][b]);
  }
  const int jb = j0 + 8 * tj;
  if (jb < J) {  // J % 8 == 0 (checked by the launcher): whole groups of 8 outputs
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bl + jb));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bl + jb + 4));
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int n = n0 + 4 * tn + a;
      if (n < N) {
        float4* o = reinterpret_cast<float4*>(out + static_cast<int64_t>(n) * J + jb);
        o[0] = make_float4(acc[a][0] + b0.x, acc[a][1] + b0.y, acc[a][2] + b0.z, acc[a][3] + b0.w);
        o[1] = make_float4(acc[a][4] + b1.x, acc[a][5] + b1.y, acc[a][6] + b1.z, acc[a][7] + b1.w);
      }
    }
  }
}

}  // namespace
}  // namespace afs

extern "C" int afs_pool3_linear_fwd(const float* x, int32_t N, int32_t H, int32_t W, int32_t C, const float* wl,
                                    const float* bl, int32_t J, float* out, afs_stream_t stream_) {
  using namespace afs;
  if (x == nullptr || wl == nullptr || bl == nullptr || out == nullptr || N < 0 || H < 3 || W < 3 || J < 1)
    return AFS_ERR_INVALID_ARG;
  if (C != kTC || H / 3 != 1 || W / 3 != 1 || (J & 7) != 0) return AFS_ERR_UNSUPPORTED;
  if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wl) | reinterpret_cast<uintptr_t>(bl) |
        reinterpret_cast<uintptr_t>(out)) & 15) != 0)
    return AFS_ERR_INVALID_ARG;
  if (N == 0) return AFS_OK;
  const dim3 grid((N + kTM - 1) / kTM, (J + kTN - 1) / kTN);
  if (grid.y > 65535) return AFS_ERR_UNSUPPORTED;
  constexpr int kSmem = (kTC * kTM + kTC * (kTN + 4)) * static_cast<int>(sizeof(float));  // 50 176 B: above the 48 KB default
  AFS_CUDA_TRY(cudaFuncSetAttribute(pool3_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  pool3_linear_kernel<<<grid, kTailThreads, kSmem, static_cast<cudaStream_t>(stream_)>>>(x, N, H, W, wl, bl, J, out);
  AFS_LAUNCH_CHECK();
  return AFS_OK;
}
