// DN4 local-descriptor head for sm_100a (fp32 path: bit-stable top-k indices).
//
// Arithmetic follows DN4Layer.forward (reference libfewshot_core/model/metric/
// dn4.py:52-73): every HW position of a clip's [C, HW] map is a descriptor,
// L2-normalised over C (F.normalize, eps 1e-12); relation = q_hat . s_hat for
// each query descriptor against the S*HW descriptors of a class; top-n_k over
// those, summed over n_k and over the query's HW descriptors.
//
// The reference materialises relation [t, wq, w, hw, s*hw] in HBM (3 MB per
// 5w5s15q episode) through cuBLAS batched GEMMs and an ATen topk.  Here:
//   1. dn4_normalize: one pass writes the normalised descriptors to scratch;
//   2. dn4_main: CTA = (64 query descriptors) x (one class), register-tiled
//      fp32 GEMM (8 rows x 4 columns per lane) from shared-memory k-chunks,
//      warp-shuffle top-k on the accumulators -- the relation tensor never
//      leaves the SM;
//   3. dn4_reduce: fixed-order sum over the HW descriptors of a query + argmax.
// Every reduction has a fixed order, so scores and indices are reproducible.
#include <limits.h>

#include "common.cuh"

namespace afs {
namespace {

constexpr int kThreads = 256;
constexpr int kRows = 64;   // query descriptors per CTA tile (8 per warp)
constexpr int kCols = 128;  // support descriptors per column tile (4 per lane)
constexpr int kKC = 32;     // channels per shared-memory chunk
constexpr int kMaxWay = 32;
constexpr int kMaxK = 8;

__global__ void __launch_bounds__(128)
dn4_normalize_kernel(const float* __restrict__ feat, int64_t n_desc, int C, int HW,
                     float* __restrict__ nf) {
  const int64_t gid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (gid >= n_desc) return;
  const int64_t row = gid / HW;
  const int m = static_cast<int>(gid - row * HW);
  const float* src = feat + row * C * HW + m;
  float* dst = nf + row * C * HW + m;
  float ss = 0.f;
  for (int c = 0; c < C; ++c) {
    const float v = src[static_cast<int64_t>(c) * HW];
    ss = fmaf(v, v, ss);
  }
  const float denom = fmaxf(sqrtf(ss), 1e-12f);  // F.normalize: x / max(||x||, eps)
  for (int c = 0; c < C; ++c) dst[static_cast<int64_t>(c) * HW] = src[static_cast<int64_t>(c) * HW] / denom;
}

__device__ __forceinline__ bool better(float v, int i, float bv, int bi) {
  return v > bv || (v == bv && i < bi);
}

__global__ void __launch_bounds__(kThreads)
dn4_main_kernel(const float* __restrict__ nf, const int32_t* __restrict__ cls_row, int W, int S,
                int C, int HW, int n_k, float* __restrict__ rowsum, int32_t* __restrict__ topk_idx) {
  __shared__ __align__(16) float sQ[kKC][kRows];
  __shared__ __align__(16) float sS[kKC][kCols];
  __shared__ int s_qbase[kMaxWay + 1];
  __shared__ int s_rowfeat[kRows], s_rowm[kRows], s_rowo[kRows];
  __shared__ float s_runv[kThreads / 32][8][kMaxK];
  __shared__ int s_runi[kThreads / 32][8][kMaxK];

  const int e = blockIdx.z;
  const int w = blockIdx.y;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;

  if (tid <= W) {
    const int g = e * W + tid;
    s_qbase[tid] = cls_row[g] - g * S;
  }
  __syncthreads();
  const int out0 = s_qbase[0];
  const int out1 = s_qbase[W];
  const int NS = S * HW;
  const int sup_row0 = cls_row[e * W + w];
  const int n_rows = (out1 - out0) * HW;
  const int n_tiles = (n_rows + kRows - 1) / kRows;
  const int n_ctiles = (NS + kCols - 1) / kCols;

  const int q_r = tid & (kRows - 1);   // this thread's row when filling sQ
  const int q_c0 = tid / kRows;        // 0..3
  const int s_n = tid & (kCols - 1);   // this thread's column when filling sS
  const int s_c0 = tid / kCols;        // 0..1

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();
    if (tid < kRows) {
      const int dr = tile * kRows + tid;
      int featrow = -1, m = 0, o = 0;
      if (dr < n_rows) {
        o = out0 + dr / HW;
        m = dr % HW;
        int cw = 0;
        while (s_qbase[cw + 1] <= o) ++cw;
        featrow = o + (e * W + cw + 1) * S;
      }
      s_rowfeat[tid] = featrow;
      s_rowm[tid] = m;
      s_rowo[tid] = o;
    }
    __syncthreads();
    const int my_featrow = s_rowfeat[q_r];
    const float* q_src = my_featrow >= 0
                             ? nf + static_cast<int64_t>(my_featrow) * C * HW + s_rowm[q_r]
                             : nullptr;

    for (int ct = 0; ct < n_ctiles; ++ct) {
      const int col = ct * kCols + s_n;
      const float* s_src = nullptr;
      if (col < NS) {
        const int s = col / HW;
        const int m = col - s * HW;
        s_src = nf + static_cast<int64_t>(sup_row0 + s) * C * HW + m;
      }
      float acc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

      for (int kc = 0; kc < C; kc += kKC) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kKC / 4; ++i) {
          const int c = q_c0 + 4 * i;
          sQ[c][q_r] = (q_src != nullptr && kc + c < C) ? __ldg(q_src + static_cast<int64_t>(kc + c) * HW) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < kKC / 2; ++i) {
          const int c = s_c0 + 2 * i;
          sS[c][s_n] = (s_src != nullptr && kc + c < C) ? __ldg(s_src + static_cast<int64_t>(kc + c) * HW) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int c = 0; c < kKC; ++c) {
          const float4 s4 = *reinterpret_cast<const float4*>(&sS[c][4 * lane]);
          const float4 qa = *reinterpret_cast<const float4*>(&sQ[c][8 * warp]);
          const float4 qb = *reinterpret_cast<const float4*>(&sQ[c][8 * warp + 4]);
          const float qv[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            acc[i][0] = fmaf(qv[i], s4.x, acc[i][0]);
            acc[i][1] = fmaf(qv[i], s4.y, acc[i][1]);
            acc[i][2] = fmaf(qv[i], s4.z, acc[i][2]);
            acc[i][3] = fmaf(qv[i], s4.w, acc[i][3]);
          }
        }
      }

      // merge this column tile into each row's running top-n_k (descending,
      // lowest column index on ties); lane k < n_k owns running entry k.
      const bool last = (ct == n_ctiles - 1);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float cv[5];
        int ci[5];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int cj = ct * kCols + 4 * lane + j;
          cv[j] = cj < NS ? acc[i][j] : -INFINITY;
          ci[j] = cj < NS ? cj : INT_MAX;
        }
        cv[4] = -INFINITY;
        ci[4] = INT_MAX;
        if (ct > 0 && lane < n_k) {
          cv[4] = s_runv[warp][i][lane];
          ci[4] = s_runi[warp][i][lane];
        }
        float newv = -INFINITY, sum = 0.f;
        int newi = -1;
        for (int k = 0; k < n_k; ++k) {
          float bv = cv[0];
          int bi = ci[0];
#pragma unroll
          for (int j = 1; j < 5; ++j)
            if (better(cv[j], ci[j], bv, bi)) { bv = cv[j]; bi = ci[j]; }
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
          }
#pragma unroll
          for (int j = 0; j < 5; ++j)
            if (ci[j] == bi) cv[j] = -INFINITY;
          if (lane == k) { newv = bv; newi = bi; }
          sum += bv;
        }
        if (!last) {
          __syncwarp();
          if (lane < n_k) {
            s_runv[warp][i][lane] = newv;
            s_runi[warp][i][lane] = newi;
          }
          __syncwarp();
        } else {
          const int r = 8 * warp + i;
          if (s_rowfeat[r] >= 0) {
            const int64_t base = (static_cast<int64_t>(s_rowo[r]) * W + w) * HW + s_rowm[r];
            if (lane == 0) rowsum[base] = sum;
            if (topk_idx != nullptr && lane < n_k) topk_idx[base * n_k + lane] = newi;
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(128)
dn4_reduce_kernel(const float* __restrict__ rowsum, int NQ, int W, int HW,
                  float* __restrict__ score, int32_t* __restrict__ pred) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= NQ) return;
  float best = -INFINITY;
  int best_w = 0;
  for (int w = 0; w < W; ++w) {
    const float* p = rowsum + (static_cast<int64_t>(o) * W + w) * HW;
    float s = 0.f;
    for (int m = 0; m < HW; ++m) s += p[m];
    score[static_cast<int64_t>(o) * W + w] = s;
    if (s > best) { best = s; best_w = w; }
  }
  if (pred != nullptr) pred[o] = best_w;
}

// ---- backward (DN4.set_forward_loss, dn4.py:122-155 under autograd) ----
// score[o,w] = sum_m sum_{k<n_k} <q_hat[o,m,:], s_hat[w, idx[o,w,m,k], :]> with the top-k selection held
// fixed (torch.topk's backward routes the gradient to the selected entries only).
// One warp per query descriptor (o, m): lanes stride over channels; d q_hat is owned (plain store),
// d s_hat is scattered with atomicAdd (several query descriptors select the same support descriptor).
__global__ void __launch_bounds__(256)
dn4_bwd_scatter_kernel(const float* __restrict__ nf, const int32_t* __restrict__ cls_row, int EW, int W, int S,
                       int C, int HW, int n_k, int NQ, const int32_t* __restrict__ topk_idx,
                       const float* __restrict__ grad_score, float* __restrict__ dnf) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (wid >= static_cast<int64_t>(NQ) * HW) return;
  const int o = static_cast<int>(wid / HW);
  const int m = static_cast<int>(wid - static_cast<int64_t>(o) * HW);
  // block g of output row o: cls_row[g] - g*S <= o < cls_row[g+1] - (g+1)*S
  int lo = 0, hi = EW - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (cls_row[mid] - mid * S <= o) lo = mid; else hi = mid - 1;
  }
  const int g = lo;
  const int e = g / W;
  const int64_t qrow = static_cast<int64_t>(o) + static_cast<int64_t>(g + 1) * S;
  const float* qn = nf + qrow * C * HW + m;
  float* dq = dnf + qrow * C * HW + m;
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int c = c0 + lane;
    float acc = 0.f;
    const float qv = c < C ? qn[static_cast<int64_t>(c) * HW] : 0.f;
    for (int w = 0; w < W; ++w) {
      const float gs = grad_score[static_cast<int64_t>(o) * W + w];
      const int64_t sup_row0 = cls_row[e * W + w];
      const int32_t* idx = topk_idx + ((static_cast<int64_t>(o) * W + w) * HW + m) * n_k;
      for (int k = 0; k < n_k; ++k) {
        const int col = idx[k];
        const int s = col / HW;
        const int ms = col - s * HW;
        const int64_t off = ((sup_row0 + s) * C + c) * HW + ms;
        if (c < C) {
          acc = fmaf(gs, nf[off], acc);
          atomicAdd(dnf + off, gs * qv);
        }
      }
    }
    if (c < C) dq[static_cast<int64_t>(c) * HW] = acc;
  }
}

// x_hat = x / max(|x|, eps):  dx = (d - x_hat <x_hat, d>) / |x|   (|x| >= eps),  d / eps otherwise
__global__ void __launch_bounds__(128)
dn4_bwd_normalize_kernel(const float* __restrict__ feat, const float* __restrict__ dnf, int64_t n_desc, int C,
                         int HW, float* __restrict__ grad_feat) {
  const int64_t gid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (gid >= n_desc) return;
  const int64_t row = gid / HW;
  const int m = static_cast<int>(gid - row * HW);
  const float* x = feat + row * C * HW + m;
  const float* d = dnf + row * C * HW + m;
  float* o = grad_feat + row * C * HW + m;
  float ss = 0.f, xd = 0.f;
  for (int c = 0; c < C; ++c) {
    const float v = x[static_cast<int64_t>(c) * HW];
    ss = fmaf(v, v, ss);
    xd = fmaf(v, d[static_cast<int64_t>(c) * HW], xd);
  }
  const float nrm = sqrtf(ss);
  if (nrm >= 1e-12f) {
    const float inv = 1.0f / nrm;
    const float k = xd * inv * inv * inv;  // <x_hat, d> / |x| * (1/|x|) applied to x
    for (int c = 0; c < C; ++c)
      o[static_cast<int64_t>(c) * HW] = d[static_cast<int64_t>(c) * HW] * inv - x[static_cast<int64_t>(c) * HW] * k;
  } else {
    for (int c = 0; c < C; ++c) o[static_cast<int64_t>(c) * HW] = d[static_cast<int64_t>(c) * HW] * 1e12f;
  }
}

size_t align256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace
}  // namespace afs

extern "C" size_t afs_dn4_workspace_bytes(int32_t N, int32_t E, int32_t W, int32_t S, int32_t C,
                                          int32_t HW) {
  if (N <= 0 || E <= 0 || W <= 0 || S <= 0 || C <= 0 || HW <= 0) return 0;
  const int64_t nq = static_cast<int64_t>(N) - static_cast<int64_t>(E) * W * S;
  if (nq < 0) return 0;
  return afs::align256(static_cast<size_t>(N) * C * HW * sizeof(float)) +
         afs::align256(static_cast<size_t>(nq) * W * HW * sizeof(float));
}

extern "C" int afs_dn4_fwd(const float* feat, const int32_t* cls_row, int32_t N, int32_t E,
                           int32_t W, int32_t S, int32_t C, int32_t HW, int32_t n_k, float* score,
                           int32_t* topk_idx, int32_t* pred, void* ws, size_t ws_bytes,
                           afs_stream_t stream_) {
  using namespace afs;
  if (feat == nullptr || cls_row == nullptr || score == nullptr || E < 0 || W < 1 || W > kMaxWay ||
      S < 1 || C < 1 || HW < 1 || N < E * W * S || n_k < 1 || n_k > kMaxK ||
      n_k > S * HW)
    return AFS_ERR_INVALID_ARG;
  const int NQ = N - E * W * S;
  if (E == 0 || NQ == 0) return AFS_OK;
  if (E > 65535) return AFS_ERR_UNSUPPORTED;
  const size_t need = afs_dn4_workspace_bytes(N, E, W, S, C, HW);
  if (ws == nullptr || ws_bytes < need) return AFS_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(ws) % 16 != 0) return AFS_ERR_INVALID_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float* nf = static_cast<float*>(ws);
  float* rowsum = reinterpret_cast<float*>(static_cast<char*>(ws) +
                                           align256(static_cast<size_t>(N) * C * HW * sizeof(float)));

  const int64_t n_desc = static_cast<int64_t>(N) * HW;
  dn4_normalize_kernel<<<static_cast<unsigned>((n_desc + 127) / 128), 128, 0, stream>>>(feat, n_desc, C, HW, nf);
  AFS_LAUNCH_CHECK();

  const int64_t avg_rows = static_cast<int64_t>(NQ) * HW / E;
  int tiles = static_cast<int>((avg_rows + kRows - 1) / kRows);
  if (tiles < 1) tiles = 1;
  dim3 grid(tiles, W, E);
  dn4_main_kernel<<<grid, kThreads, 0, stream>>>(nf, cls_row, W, S, C, HW, n_k, rowsum, topk_idx);
  AFS_LAUNCH_CHECK();

  dn4_reduce_kernel<<<(NQ + 127) / 128, 128, 0, stream>>>(rowsum, NQ, W, HW, score, pred);
  AFS_LAUNCH_CHECK();
  return AFS_OK;
}

extern "C" size_t afs_dn4_bwd_workspace_bytes(int32_t N, int32_t C, int32_t HW) {
  if (N <= 0 || C <= 0 || HW <= 0) return 0;
  return 2 * afs::align256(static_cast<size_t>(N) * C * HW * sizeof(float));
}

extern "C" int afs_dn4_bwd(const float* feat, const int32_t* cls_row, int32_t N, int32_t E, int32_t W,
                           int32_t S, int32_t C, int32_t HW, int32_t n_k, const int32_t* topk_idx,
                           const float* grad_score, float* grad_feat, void* ws, size_t ws_bytes,
                           afs_stream_t stream_) {
  using namespace afs;
  if (feat == nullptr || cls_row == nullptr || topk_idx == nullptr || grad_score == nullptr ||
      grad_feat == nullptr || E < 0 || W < 1 || W > kMaxWay || S < 1 || C < 1 || HW < 1 ||
      N < E * W * S || n_k < 1 || n_k > kMaxK)
    return AFS_ERR_INVALID_ARG;
  if (N == 0) return AFS_OK;
  const size_t need = afs_dn4_bwd_workspace_bytes(N, C, HW);
  if (ws == nullptr || ws_bytes < need) return AFS_ERR_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t half = align256(static_cast<size_t>(N) * C * HW * sizeof(float));
  float* nf = static_cast<float*>(ws);
  float* dnf = reinterpret_cast<float*>(static_cast<char*>(ws) + half);
  const int NQ = N - E * W * S;
  const int64_t n_desc = static_cast<int64_t>(N) * HW;
  dn4_normalize_kernel<<<static_cast<unsigned>((n_desc + 127) / 128), 128, 0, stream>>>(feat, n_desc, C, HW, nf);
  AFS_LAUNCH_CHECK();
  AFS_CUDA_TRY(cudaMemsetAsync(dnf, 0, static_cast<size_t>(N) * C * HW * sizeof(float), stream));
  if (NQ > 0) {
    const int64_t warps = static_cast<int64_t>(NQ) * HW;
    dn4_bwd_scatter_kernel<<<static_cast<unsigned>((warps * 32 + 255) / 256), 256, 0, stream>>>(
        nf, cls_row, E * W, W, S, C, HW, n_k, NQ, topk_idx, grad_score, dnf);
    AFS_LAUNCH_CHECK();
  }
  dn4_bwd_normalize_kernel<<<static_cast<unsigned>((n_desc + 127) / 128), 128, 0, stream>>>(feat, dnf, n_desc, C,
                                                                                           HW, grad_feat);
  AFS_LAUNCH_CHECK();
  return AFS_OK;
}
