// Brownian-distance-covariance pooling for sm_100a.
//
// Arithmetic follows BDCovpool + Triuvec (reference libfewshot_core/model/
// backbone/utils/bdc_pool.py:69-93), with X in R^{C x M} the channel rows of
// one clip's feature map:
//   G = X X^T;  D_ij = max(G_ii + G_jj - 2 G_ij, 0);  A_ij = sqrt(exp(t) D_ij + 1e-5)
//   B_ij = A_ij - rowsum_i/C - colsum_j/C + total/C^2;   out = row-major triu(B)
// The reference spends 7 bmm (three of them against an all-ones matrix), two
// [B,C,C] temporaries per call and a CPU-built gather index.  Here one CTA per
// clip streams X once through shared memory (transposed so both Gram operands
// are 128-bit shared loads), keeps the 64x64 Gram in registers (4x4 per
// thread), and finishes the whole epilogue on chip.  Algorithmic bytes per
// clip: 4*C*M + 4*C(C+1)/2.
#include "common.cuh"

namespace afs {
namespace {

constexpr int kThreads = 256;
constexpr int kC = 64;          // padded channel count
constexpr int kMC = 64;         // spatial positions per shared-memory chunk
constexpr int kXStride = kC + 4;  // sX[m][c], 16-byte aligned rows, conflict-free 128-bit stores
constexpr int kGStride = kC + 1;

__global__ void __launch_bounds__(kThreads)
bdc_kernel(const float* __restrict__ x, int C, int M, const float* __restrict__ log_temp, int triu,
           float* __restrict__ out) {
  __shared__ __align__(16) float sX[kMC * kXStride];
  __shared__ float sA[kC * kGStride];
  __shared__ float s_diag[kC];
  __shared__ float s_rowsum[kC];
  __shared__ float s_colsum[kC];
  __shared__ float s_total;

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int ti = tid >> 4;  // Gram rows 4*ti .. 4*ti+3
  const int tj = tid & 15;  // Gram cols 4*tj .. 4*tj+3
  const float* xb = x + static_cast<int64_t>(b) * C * M;

  float g[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) g[i][j] = 0.f;

  // loader mapping: thread -> (position lm, channel quad c4); 64 positions x 16 quads = 1024
  // slots, 4 per thread; lanes run along m so the global loads are coalesced.
  for (int m0 = 0; m0 < M; m0 += kMC) {
    __syncthreads();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int slot = tid + kThreads * it;
      const int lm = slot & (kMC - 1);
      const int c4 = slot >> 6;
      const int m = m0 + lm;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < M) {
        const int c = 4 * c4;
        if (c + 0 < C) v.x = __ldg(xb + static_cast<int64_t>(c + 0) * M + m);
        if (c + 1 < C) v.y = __ldg(xb + static_cast<int64_t>(c + 1) * M + m);
        if (c + 2 < C) v.z = __ldg(xb + static_cast<int64_t>(c + 2) * M + m);
        if (c + 3 < C) v.w = __ldg(xb + static_cast<int64_t>(c + 3) * M + m);
      }
      *reinterpret_cast<float4*>(&sX[lm * kXStride + 4 * c4]) = v;
    }
    __syncthreads();
#pragma unroll 8
    for (int lm = 0; lm < kMC; ++lm) {
      const float4 a = *reinterpret_cast<const float4*>(&sX[lm * kXStride + 4 * ti]);
      const float4 c = *reinterpret_cast<const float4*>(&sX[lm * kXStride + 4 * tj]);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        g[i][0] = fmaf(av[i], c.x, g[i][0]);
        g[i][1] = fmaf(av[i], c.y, g[i][1]);
        g[i][2] = fmaf(av[i], c.z, g[i][2]);
        g[i][3] = fmaf(av[i], c.w, g[i][3]);
      }
    }
  }

  if (ti == tj) {
#pragma unroll
    for (int i = 0; i < 4; ++i) s_diag[4 * ti + i] = g[i][i];
  }
  __syncthreads();

  const float et = expf(__ldg(log_temp));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = 4 * ti + i, c = 4 * tj + j;
      float d = s_diag[c] + s_diag[r] - 2.f * g[i][j];
      d = fmaxf(d, 0.f);
      d = et * d;
      sA[r * kGStride + c] = (r < C && c < C) ? sqrtf(d + 1e-5f) : 0.f;
    }
  }
  __syncthreads();
  if (tid < kC) {  // row sums (dcov.bmm(I_M)): fixed order over j
    float s = 0.f;
    for (int j = 0; j < C; ++j) s += sA[tid * kGStride + j];
    s_rowsum[tid] = s;
  } else if (tid < 2 * kC) {  // column sums (I_M.bmm(dcov)): fixed order over i
    const int c = tid - kC;
    float s = 0.f;
    for (int i = 0; i < C; ++i) s += sA[i * kGStride + c];
    s_colsum[c] = s;
  }
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int j = 0; j < C; ++j) s += s_colsum[j];
    s_total = s;
  }
  __syncthreads();

  const float inv = 1.0f / static_cast<float>(C);
  const float inv2 = 1.0f / static_cast<float>(C * C);
  const float total = s_total;
  if (triu) {
    float* ob = out + static_cast<int64_t>(b) * (C * (C + 1) / 2);
    for (int idx = tid; idx < C * C; idx += kThreads) {
      const int r = idx / C, c = idx - r * C;
      if (c >= r) {
        const float v = sA[r * kGStride + c] - inv * s_rowsum[r] - inv * s_colsum[c] + inv2 * total;
        ob[r * C - (r * (r - 1)) / 2 + (c - r)] = v;
      }
    }
  } else {
    float* ob = out + static_cast<int64_t>(b) * C * C;
    for (int idx = tid; idx < C * C; idx += kThreads) {
      const int r = idx / C, c = idx - r * C;
      ob[idx] = sA[r * kGStride + c] - inv * s_rowsum[r] - inv * s_colsum[c] + inv2 * total;
    }
  }
}

// ---- backward (DeepBDC.set_forward_loss through BdcPool, deepbdc.py:354-378 under autograd) ----
// Recomputes G, D, A exactly as the forward, then
//   dA = dB - rowmean(dB) - colmean(dB) + mean(dB)           (double centring is self-adjoint)
//   dD = dA * e^t / (2 A)  where the raw distance is >= 0 (torch.clamp's subgradient), dt = sum dA * e^t D / (2 A)
//   dG_ij = -2 dD_ij + [i == j] (rowsum_i(dD) + colsum_i(dD)),   dX = (dG + dG^T) X
// One CTA per clip; dX is a [C x C] . [C x M] product done from shared memory in chunks of 64 positions.
constexpr int kTStride = kMC + 1;

__global__ void __launch_bounds__(kThreads)
bdc_bwd_kernel(const float* __restrict__ x, int C, int M, const float* __restrict__ log_temp, int triu,
               const float* __restrict__ grad_out, float* __restrict__ grad_x, float* __restrict__ grad_t) {
  extern __shared__ __align__(16) float s_dyn[];      // > 48 KB in total: dynamic shared memory
  float* sX = s_dyn;                                  // forward: sX[m][c]; dX phase: sXT[c][m] (kTStride)
  float* sA = sX + kMC * kXStride;                    // H = dG + dG^T
  float* sB = sA + kC * kGStride;                     // dB, later dD
  __shared__ float s_diag[kC], s_rowsum[kC], s_colsum[kC];
  __shared__ float s_total, s_dt[kThreads / 32];
  static_assert(kC * kTStride <= kMC * kXStride, "transposed chunk must fit in sX");

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int ti = tid >> 4, tj = tid & 15;
  const float* xb = x + static_cast<int64_t>(b) * C * M;

  float g[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) g[i][j] = 0.f;
  for (int m0 = 0; m0 < M; m0 += kMC) {
    __syncthreads();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int slot = tid + kThreads * it;
      const int lm = slot & (kMC - 1);
      const int c4 = slot >> 6;
      const int m = m0 + lm;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < M) {
        const int c = 4 * c4;
        if (c + 0 < C) v.x = __ldg(xb + static_cast<int64_t>(c + 0) * M + m);
        if (c + 1 < C) v.y = __ldg(xb + static_cast<int64_t>(c + 1) * M + m);
        if (c + 2 < C) v.z = __ldg(xb + static_cast<int64_t>(c + 2) * M + m);
        if (c + 3 < C) v.w = __ldg(xb + static_cast<int64_t>(c + 3) * M + m);
      }
      *reinterpret_cast<float4*>(&sX[lm * kXStride + 4 * c4]) = v;
    }
    __syncthreads();
#pragma unroll 8
    for (int lm = 0; lm < kMC; ++lm) {
      const float4 a = *reinterpret_cast<const float4*>(&sX[lm * kXStride + 4 * ti]);
      const float4 c = *reinterpret_cast<const float4*>(&sX[lm * kXStride + 4 * tj]);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        g[i][0] = fmaf(av[i], c.x, g[i][0]);
        g[i][1] = fmaf(av[i], c.y, g[i][1]);
        g[i][2] = fmaf(av[i], c.z, g[i][2]);
        g[i][3] = fmaf(av[i], c.w, g[i][3]);
      }
    }
  }
  if (ti == tj) {
#pragma unroll
    for (int i = 0; i < 4; ++i) s_diag[4 * ti + i] = g[i][i];
  }
  // dB (upper triangle of the incoming gradient, zero elsewhere)
  for (int idx = tid; idx < kC * kC; idx += kThreads) {
    const int r = idx >> 6, c = idx & (kC - 1);
    float v = 0.f;
    if (r < C && c < C) {
      if (triu) {
        if (c >= r) v = grad_out[static_cast<int64_t>(b) * (C * (C + 1) / 2) + r * C - (r * (r - 1)) / 2 + (c - r)];
      } else {
        v = grad_out[static_cast<int64_t>(b) * C * C + r * C + c];
      }
    }
    sB[r * kGStride + c] = v;
  }
  __syncthreads();
  if (tid < kC) {
    float s = 0.f;
    for (int j = 0; j < C; ++j) s += sB[tid * kGStride + j];
    s_rowsum[tid] = s;
  } else if (tid < 2 * kC) {
    const int c = tid - kC;
    float s = 0.f;
    for (int i = 0; i < C; ++i) s += sB[i * kGStride + c];
    s_colsum[c] = s;
  }
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int j = 0; j < C; ++j) s += s_colsum[j];
    s_total = s;
  }
  __syncthreads();

  const float et = expf(__ldg(log_temp));
  const float inv = 1.0f / static_cast<float>(C);
  const float inv2 = 1.0f / static_cast<float>(C * C);
  float dd[4][4];
  float dt = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = 4 * ti + i, c = 4 * tj + j;
      const float raw = s_diag[c] + s_diag[r] - 2.f * g[i][j];
      const float ds = et * fmaxf(raw, 0.f);
      const float a = sqrtf(ds + 1e-5f);
      float v = 0.f;
      if (r < C && c < C) {
        const float dA = sB[r * kGStride + c] - inv * s_rowsum[r] - inv * s_colsum[c] + inv2 * s_total;
        const float h = dA / (2.f * a);
        dt = fmaf(h, ds, dt);
        v = raw >= 0.f ? h * et : 0.f;
      }
      dd[i][j] = v;
    }
  }
  __syncthreads();  // every thread has read dB / its sums
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) sB[(4 * ti + i) * kGStride + 4 * tj + j] = dd[i][j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dt += __shfl_xor_sync(0xffffffffu, dt, o);
  if ((tid & 31) == 0) s_dt[tid >> 5] = dt;
  __syncthreads();
  if (tid < kC) {
    float s = 0.f;
    for (int j = 0; j < kC; ++j) s += sB[tid * kGStride + j];
    s_rowsum[tid] = s;
  } else if (tid < 2 * kC) {
    const int c = tid - kC;
    float s = 0.f;
    for (int i = 0; i < kC; ++i) s += sB[i * kGStride + c];
    s_colsum[c] = s;
  }
  if (tid == 0) {
    float s = 0.f;
    for (int k = 0; k < kThreads / 32; ++k) s += s_dt[k];
    grad_t[b] = s;
  }
  __syncthreads();
  // H = dG + dG^T
  for (int idx = tid; idx < kC * kC; idx += kThreads) {
    const int r = idx >> 6, c = idx & (kC - 1);
    float v = -2.f * (sB[r * kGStride + c] + sB[c * kGStride + r]);
    if (r == c) v += 2.f * (s_rowsum[r] + s_colsum[r]);
    sA[r * kGStride + c] = v;
  }
  __syncthreads();

  // dX[c][m] = sum_j H[c][j] X[j][m]; thread -> position lm = tid & 63, channels 16*(tid >> 6) .. +15
  const int lm = tid & (kMC - 1);
  const int cq = tid >> 6;
  float* gxb = grad_x + static_cast<int64_t>(b) * C * M;
  for (int m0 = 0; m0 < M; m0 += kMC) {
    __syncthreads();
    for (int idx = tid; idx < kC * kMC; idx += kThreads) {
      const int c = idx >> 6, l = idx & (kMC - 1);
      sX[c * kTStride + l] = (c < C && m0 + l < M) ? __ldg(xb + static_cast<int64_t>(c) * M + m0 + l) : 0.f;
    }
    __syncthreads();
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    for (int j = 0; j < kC; ++j) {
      const float xv = sX[j * kTStride + lm];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = fmaf(sA[(16 * cq + i) * kGStride + j], xv, acc[i]);
    }
    if (m0 + lm < M) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = 16 * cq + i;
        if (c < C) gxb[static_cast<int64_t>(c) * M + m0 + lm] = acc[i];
      }
    }
  }
}

}  // namespace
}  // namespace afs

extern "C" int afs_bdc_bwd(const float* x, int32_t B, int32_t C, int32_t M, const float* log_temp,
                           int32_t triu, const float* grad_out, float* grad_x, float* grad_log_temp,
                           afs_stream_t stream_) {
  using namespace afs;
  if (x == nullptr || log_temp == nullptr || grad_out == nullptr || grad_x == nullptr ||
      grad_log_temp == nullptr || B < 0 || C < 1 || M < 1)
    return AFS_ERR_INVALID_ARG;
  if (C > kC) return AFS_ERR_UNSUPPORTED;
  if (B == 0) return AFS_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  constexpr size_t smem = (kMC * kXStride + 2 * kC * kGStride) * sizeof(float);
  AFS_CUDA_TRY(cudaFuncSetAttribute(bdc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(smem)));
  bdc_bwd_kernel<<<B, kThreads, smem, stream>>>(x, C, M, log_temp, triu, grad_out, grad_x, grad_log_temp);
  AFS_LAUNCH_CHECK();
  return AFS_OK;
}

namespace afs {
int bdc_fwd_tc(const float* x, int32_t B, int32_t C, int32_t M, const float* log_temp, int32_t triu, float* out,
               cudaStream_t stream);  // bdc_tc.cu
static bool g_bdc_tensor_core = true;
}  // namespace afs

extern "C" int afs_bdc_set_tensor_core(int32_t enable) {
  afs::g_bdc_tensor_core = enable != 0;
  return AFS_OK;
}

extern "C" int afs_bdc_fwd(const float* x, int32_t B, int32_t C, int32_t M, const float* log_temp,
                           int32_t triu, float* out, afs_stream_t stream_) {
  using namespace afs;
  if (x == nullptr || log_temp == nullptr || out == nullptr || B < 0 || C < 1 || M < 1)
    return AFS_ERR_INVALID_ARG;
  if (C > kC) return AFS_ERR_UNSUPPORTED;
  if (B == 0) return AFS_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (g_bdc_tensor_core) {  // C == 64, M % 4 == 0: Gram on tcgen05 (3 x TF32), same epilogue arithmetic
    const int r = bdc_fwd_tc(x, B, C, M, log_temp, triu, out, stream);
    if (r != AFS_ERR_UNSUPPORTED) return r;
  }
  bdc_kernel<<<B, kThreads, 0, stream>>>(x, C, M, log_temp, triu, out);
  AFS_LAUNCH_CHECK();
  return AFS_OK;
}
