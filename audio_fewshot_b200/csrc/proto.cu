// ProtoNet / DeepBDC prototype head for sm_100a.
//
// Arithmetic follows ProtoLayer.forward (reference libfewshot_core/model/metric/
// proto_net.py:49-63) and deepbdc.ProtoLayer.forward (deepbdc.py:34-53):
//   p[e,w,:] = (1/S) sum_s support[e,w,s,:]
//   euclid :  logit = -sum_d (q_d - p_d)^2      (direct difference, fp32)
//   cosine :  logit = <q, p> / (max(|q|,1e-12) max(|p|,1e-12))
//   dot    :  logit = <q, p>
// The reference materialises a [t,wq,w,c] temporary and launches one ATen op
// per step and per episode.  Here a pre-kernel writes the E*W prototypes once
// (support rows are read exactly once), and one CTA per (episode, row split)
// streams every query row from HBM exactly once with 128-bit loads, eight in
// flight per thread, against a SLICE of the prototypes held in shared memory:
// wide features (D = 12 800 for flattened ResNet-12 maps) walk the slices in
// ascending column order with the partial sums kept in registers, so the
// summation order -- and therefore every logit bit -- is the same whatever the
// slice size.  HBM-bound: 4*W*(S+Q)*D + 4*WQ*W bytes per episode.
#include "common.cuh"

namespace afs {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxWay = 32;
constexpr int kSliceBytes = 40 * 1024;  // prototype slice in shared memory: W * DS * 4 bytes <= this (5 CTAs per SM)

__global__ void __launch_bounds__(256) proto_mean_kernel(const float* __restrict__ feat,
                                                         int64_t ld,
                                                         const int32_t* __restrict__ cls_row,
                                                         int EW, int S, int D4,
                                                         float4* __restrict__ protos) {
  const int64_t total = static_cast<int64_t>(EW) * D4;
  const float inv = static_cast<float>(S);
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx / D4);
    const int c = static_cast<int>(idx - static_cast<int64_t>(g) * D4);
    const float* base = feat + static_cast<int64_t>(cls_row[g]) * ld + 4 * c;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < S; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(base + s * ld);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    protos[idx] = make_float4(acc.x / inv, acc.y / inv, acc.z / inv, acc.w / inv);
  }
}

// ||p||^-1 of every prototype (cosine mode), one warp per prototype
__global__ void __launch_bounds__(256) proto_pinv_kernel(const float4* __restrict__ protos, int EW, int D4,
                                                         float* __restrict__ pinv) {
  const int lane = threadIdx.x & 31;
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= EW) return;
  float ss = 0.f;
  for (int c = lane; c < D4; c += 32) {
    const float4 p = protos[static_cast<int64_t>(g) * D4 + c];
    ss = fmaf(p.x, p.x, ss); ss = fmaf(p.y, p.y, ss);
    ss = fmaf(p.z, p.z, ss); ss = fmaf(p.w, p.w, ss);
  }
  ss = warp_sum(ss);
  if (lane == 0) pinv[g] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
}

template <int MODE, int WT, int RW>
__device__ __forceinline__ void proto_accumulate(const float4 (&q)[RW], const float4* __restrict__ proto, int DS4, int c,
                                                 int W, float (&acc)[RW][WT], float (&qq)[RW]) {
  if (MODE == AFS_PROTO_COSINE) {
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      qq[r] = fmaf(q[r].x, q[r].x, qq[r]); qq[r] = fmaf(q[r].y, q[r].y, qq[r]);
      qq[r] = fmaf(q[r].z, q[r].z, qq[r]); qq[r] = fmaf(q[r].w, q[r].w, qq[r]);
    }
  }
#pragma unroll
  for (int w = 0; w < WT; ++w) {
    if (w < W) {
      const float4 p = proto[w * DS4 + c];
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        if (MODE == AFS_PROTO_EUCLIDEAN) {
          const float dx = q[r].x - p.x, dy = q[r].y - p.y;
          const float dz = q[r].z - p.z, dw = q[r].w - p.w;
          acc[r][w] = fmaf(dx, dx, acc[r][w]); acc[r][w] = fmaf(dy, dy, acc[r][w]);
          acc[r][w] = fmaf(dz, dz, acc[r][w]); acc[r][w] = fmaf(dw, dw, acc[r][w]);
        } else {
          acc[r][w] = fmaf(q[r].x, p.x, acc[r][w]); acc[r][w] = fmaf(q[r].y, p.y, acc[r][w]);
          acc[r][w] = fmaf(q[r].z, p.z, acc[r][w]); acc[r][w] = fmaf(q[r].w, p.w, acc[r][w]);
        }
      }
    }
  }
}

template <int MODE, int WT, int RW>
__global__ void __launch_bounds__(kThreads)
proto_fwd_kernel(const float* __restrict__ feat, int64_t ld, const int32_t* __restrict__ cls_row,
                 int W, int S, int D4, int DS4, float* __restrict__ logits, int32_t* __restrict__ pred,
                 const float4* __restrict__ protos_g, const float* __restrict__ pinv_g) {
  extern __shared__ float4 s_proto[];  // [W][DS4]: the current column slice of this episode's prototypes
  __shared__ int s_qbase[kMaxWay + 1];

  const int e = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;

  if (tid <= W) {
    const int g = e * W + tid;
    s_qbase[tid] = cls_row[g] - g * S;
  }
  __syncthreads();
  const float4* pe = protos_g + static_cast<int64_t>(e) * W * D4;
  const int out0 = s_qbase[0];
  const int out1 = s_qbase[W];
  const int ngroups = (out1 - out0 + RW - 1) / RW;
  const int n_slices = (D4 + DS4 - 1) / DS4;
  bool filled = false;

  // every warp of the CTA makes the same number of trips (the slice loop synchronises the CTA)
  for (int g0 = blockIdx.y * kWarps; g0 < ngroups; g0 += gridDim.y * kWarps) {
    const int grp = g0 + warp;
    const bool active = grp < ngroups;
    const int o_base = out0 + grp * RW;
    const float* qptr[RW];
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      int o = active ? o_base + r : out0;
      if (o >= out1) o = out1 - 1;  // tail rows recompute the last row; never stored
      int w = 0;
      while (s_qbase[w + 1] <= o) ++w;
      qptr[r] = feat + static_cast<int64_t>(o + (e * W + w + 1) * S) * ld;
    }

    float acc[RW][WT];
    float qq[RW];
#pragma unroll
    for (int r = 0; r < RW; ++r) {
      qq[r] = 0.f;
#pragma unroll
      for (int w = 0; w < WT; ++w) acc[r][w] = 0.f;
    }

    for (int sl = 0; sl < n_slices; ++sl) {
      const int c0 = sl * DS4;
      const int cn = min(DS4, D4 - c0);
      if (n_slices > 1 || !filled) {
        __syncthreads();  // the previous slice has been consumed
        for (int idx = tid; idx < W * cn; idx += kThreads) {
          const int w = idx / cn, c = idx - w * cn;
          s_proto[w * DS4 + c] = __ldg(pe + static_cast<int64_t>(w) * D4 + c0 + c);
        }
        __syncthreads();
        filled = true;
      }
      if (active) {
        int c = lane;
        for (; c + 32 < cn; c += 64) {  // two column chunks per trip: 2 * RW 128-bit loads in flight per thread
          float4 qa[RW], qb[RW];
#pragma unroll
          for (int r = 0; r < RW; ++r) qa[r] = ldg_stream4(qptr[r] + 4 * (c0 + c));
#pragma unroll
          for (int r = 0; r < RW; ++r) qb[r] = ldg_stream4(qptr[r] + 4 * (c0 + c + 32));
          proto_accumulate<MODE, WT, RW>(qa, s_proto, DS4, c, W, acc, qq);
          proto_accumulate<MODE, WT, RW>(qb, s_proto, DS4, c + 32, W, acc, qq);
        }
        if (c < cn) {
          float4 qa[RW];
#pragma unroll
          for (int r = 0; r < RW; ++r) qa[r] = ldg_stream4(qptr[r] + 4 * (c0 + c));
          proto_accumulate<MODE, WT, RW>(qa, s_proto, DS4, c, W, acc, qq);
        }
      }
    }
    if (!active) continue;

#pragma unroll
    for (int r = 0; r < RW; ++r) {
      if (MODE == AFS_PROTO_COSINE) qq[r] = warp_sum(qq[r]);
#pragma unroll
      for (int w = 0; w < WT; ++w) acc[r][w] = warp_sum(acc[r][w]);
    }

#pragma unroll
    for (int r = 0; r < RW; ++r) {
      const int o = o_base + r;
      if (lane == r && o < out1) {
        float best = -INFINITY;
        int best_w = 0;
        const float qinv = (MODE == AFS_PROTO_COSINE) ? 1.0f / fmaxf(sqrtf(qq[r]), 1e-12f) : 1.f;
#pragma unroll
        for (int w = 0; w < WT; ++w) {
          if (w < W) {
            float v;
            if (MODE == AFS_PROTO_EUCLIDEAN) v = -acc[r][w];
            else if (MODE == AFS_PROTO_COSINE) v = acc[r][w] * qinv * pinv_g[e * W + w];
            else v = acc[r][w];
            logits[static_cast<int64_t>(o) * W + w] = v;
            if (v > best) { best = v; best_w = w; }
          }
        }
        if (pred != nullptr) pred[o] = best_w;
      }
    }
  }
}

template <int MODE, int WT, int RW>
int launch_fwd(const float* feat, int64_t ld, const int32_t* cls_row, int N, int E, int W, int S,
               int D, float* logits, int32_t* pred, const float4* protos_g, const float* pinv_g,
               cudaStream_t stream) {
  auto kern = proto_fwd_kernel<MODE, WT, RW>;
  const int D4 = D / 4;
  int DS4 = kSliceBytes / (16 * W);  // float4 columns of a slice
  DS4 -= DS4 % 64;                   // whole double-chunk trips
  if (DS4 < 64) DS4 = 64;
  if (DS4 > D4) DS4 = D4;
  const size_t smem = static_cast<size_t>(W) * DS4 * sizeof(float4);
  if (smem > 48 * 1024) {
    AFS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  const int nq = N - E * W * S;
  const int avg_groups = (nq / (E > 0 ? E : 1) + RW - 1) / RW;
  int max_split = (avg_groups + kWarps - 1) / kWarps;
  if (max_split < 1) max_split = 1;
  int want = (4 * kNumSMs + E - 1) / E;  // enough CTAs for four per SM
  int nsplit = want < max_split ? want : max_split;
  if (nsplit < 1) nsplit = 1;
  dim3 grid(E, nsplit);
  kern<<<grid, kThreads, smem, stream>>>(feat, ld, cls_row, W, S, D4, DS4, logits, pred, protos_g, pinv_g);
  AFS_LAUNCH_CHECK();
  return AFS_OK;
}

template <int MODE>
int dispatch_way(const float* feat, int64_t ld, const int32_t* cls_row, int N, int E, int W,
                 int S, int D, float* logits, int32_t* pred, const float4* protos_g, const float* pinv_g,
                 cudaStream_t stream) {
  if (W <= 5)
    return launch_fwd<MODE, 5, 4>(feat, ld, cls_row, N, E, W, S, D, logits, pred, protos_g, pinv_g, stream);
  if (W <= 8)
    return launch_fwd<MODE, 8, 2>(feat, ld, cls_row, N, E, W, S, D, logits, pred, protos_g, pinv_g, stream);
  if (W <= 16)
    return launch_fwd<MODE, 16, 1>(feat, ld, cls_row, N, E, W, S, D, logits, pred, protos_g, pinv_g, stream);
  return launch_fwd<MODE, 32, 1>(feat, ld, cls_row, N, E, W, S, D, logits, pred, protos_g, pinv_g, stream);
}

int dispatch_mode(int mode, const float* feat, int64_t ld, const int32_t* cls_row, int N, int E,
                  int W, int S, int D, float* logits, int32_t* pred, const float4* protos_g, const float* pinv_g,
                  cudaStream_t stream) {
  switch (mode) {
    case AFS_PROTO_EUCLIDEAN:
      return dispatch_way<AFS_PROTO_EUCLIDEAN>(feat, ld, cls_row, N, E, W, S, D, logits, pred, protos_g, pinv_g, stream);
    case AFS_PROTO_COSINE:
      return dispatch_way<AFS_PROTO_COSINE>(feat, ld, cls_row, N, E, W, S, D, logits, pred, protos_g, pinv_g, stream);
    case AFS_PROTO_DOT:
      return dispatch_way<AFS_PROTO_DOT>(feat, ld, cls_row, N, E, W, S, D, logits, pred, protos_g, pinv_g, stream);
    default:
      return AFS_ERR_INVALID_ARG;
  }
}

// Backward: one thread per feature column d, one CTA per (episode, column
// slice); prototypes and their gradients live in registers, so nothing is
// reduced across threads and no atomics are needed.
//   euclid: dq = -2 sum_w G_w (q - p_w);  dp_w =  2 sum_o G_ow (q_o - p_w)
//   dot   : dq =    sum_w G_w p_w;        dp_w =    sum_o G_ow q_o
//   d support_{w,s} = dp_w / S
template <int MODE, int WT>
__global__ void __launch_bounds__(128)
proto_bwd_kernel(const float* __restrict__ feat, int64_t ld, const int32_t* __restrict__ cls_row,
                 int W, int S, int D, const float* __restrict__ grad_logits,
                 float* __restrict__ grad_feat, int64_t ldg) {
  const int e = blockIdx.x;
  const int d = blockIdx.y * blockDim.x + threadIdx.x;
  if (d >= D) return;
  float p[WT], dp[WT];
  const float fS = static_cast<float>(S);
#pragma unroll
  for (int w = 0; w < WT; ++w) {
    p[w] = 0.f;
    dp[w] = 0.f;
    if (w < W) {
      const int row0 = cls_row[e * W + w];
      float acc = 0.f;
      for (int s = 0; s < S; ++s) acc += feat[static_cast<int64_t>(row0 + s) * ld + d];
      p[w] = acc / fS;
    }
  }
  for (int wc = 0; wc < W; ++wc) {
    const int g = e * W + wc;
    const int r0 = cls_row[g] + S;
    const int r1 = cls_row[g + 1];
    for (int row = r0; row < r1; ++row) {
      const int o = row - (g + 1) * S;
      const float q = feat[static_cast<int64_t>(row) * ld + d];
      const float* G = grad_logits + static_cast<int64_t>(o) * W;
      float gq = 0.f;
#pragma unroll
      for (int w = 0; w < WT; ++w) {
        if (w < W) {
          const float gw = __ldg(G + w);
          if (MODE == AFS_PROTO_EUCLIDEAN) {
            const float diff = q - p[w];
            gq = fmaf(-2.f * gw, diff, gq);
            dp[w] = fmaf(2.f * gw, diff, dp[w]);
          } else {
            gq = fmaf(gw, p[w], gq);
            dp[w] = fmaf(gw, q, dp[w]);
          }
        }
      }
      grad_feat[static_cast<int64_t>(row) * ldg + d] = gq;
    }
  }
#pragma unroll
  for (int w = 0; w < WT; ++w) {
    if (w < W) {
      const int row0 = cls_row[e * W + w];
      const float v = dp[w] / fS;
      for (int s = 0; s < S; ++s) grad_feat[static_cast<int64_t>(row0 + s) * ldg + d] = v;
    }
  }
}

// ---- cosine mode backward (MetaBaseline, reference meta_baseline.py:305-332 under autograd) ----
// logit_ow = <q_o, p_w> / (nq_o np_w),  n = max(|.|, 1e-12)  (F.normalize).  With c = logit_ow:
//   dq_o  = sum_w G_ow ( p_w / (nq_o np_w) - c q_o / nq_o^2 )      (second term only where |q_o| >= eps)
//   dp_w  = sum_o G_ow ( q_o / (nq_o np_w) - c p_w / np_w^2 )
// Row norms come from a small pre-pass (one warp per row); prototypes are recomputed per column as above.
__global__ void __launch_bounds__(256)
proto_norms_kernel(const float* __restrict__ feat, int64_t ld, const int32_t* __restrict__ cls_row, int EW, int W,
                   int S, int D, int NQ, float* __restrict__ qnorm, float* __restrict__ pnorm) {
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= NQ + EW) return;
  float ss = 0.f;
  if (wid < NQ) {  // query output row o -> feature row
    int lo = 0, hi = EW - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (cls_row[mid] - mid * S <= wid) lo = mid; else hi = mid - 1;
    }
    const float* q = feat + (static_cast<int64_t>(wid) + static_cast<int64_t>(lo + 1) * S) * ld;
    for (int d = lane; d < D; d += 32) ss = fmaf(q[d], q[d], ss);
    ss = warp_sum(ss);
    if (lane == 0) qnorm[wid] = sqrtf(ss);
  } else {
    const int g = wid - NQ;
    const float* s0 = feat + static_cast<int64_t>(cls_row[g]) * ld;
    const float fS = static_cast<float>(S);
    for (int d = lane; d < D; d += 32) {
      float acc = 0.f;
      for (int s = 0; s < S; ++s) acc += s0[static_cast<int64_t>(s) * ld + d];
      const float pv = acc / fS;
      ss = fmaf(pv, pv, ss);
    }
    ss = warp_sum(ss);
    if (lane == 0) pnorm[g] = sqrtf(ss);
  }
}

template <int WT>
__global__ void __launch_bounds__(128)
proto_bwd_cos_kernel(const float* __restrict__ feat, int64_t ld, const int32_t* __restrict__ cls_row, int W, int S,
                     int D, const float* __restrict__ grad_logits, const float* __restrict__ logits,
                     const float* __restrict__ qnorm, const float* __restrict__ pnorm,
                     float* __restrict__ grad_feat, int64_t ldg) {
  const int e = blockIdx.x;
  const int d = blockIdx.y * blockDim.x + threadIdx.x;
  if (d >= D) return;
  float p[WT], dp[WT], ip[WT];
  bool pbig[WT];
  const float fS = static_cast<float>(S);
#pragma unroll
  for (int w = 0; w < WT; ++w) {
    p[w] = 0.f; dp[w] = 0.f; ip[w] = 0.f; pbig[w] = false;
    if (w < W) {
      const int row0 = cls_row[e * W + w];
      float acc = 0.f;
      for (int s = 0; s < S; ++s) acc += feat[static_cast<int64_t>(row0 + s) * ld + d];
      p[w] = acc / fS;
      const float n = __ldg(pnorm + e * W + w);
      pbig[w] = n >= 1e-12f;
      ip[w] = 1.0f / fmaxf(n, 1e-12f);
    }
  }
  for (int wc = 0; wc < W; ++wc) {
    const int g = e * W + wc;
    const int r0 = cls_row[g] + S;
    const int r1 = cls_row[g + 1];
    for (int row = r0; row < r1; ++row) {
      const int o = row - (g + 1) * S;
      const float q = feat[static_cast<int64_t>(row) * ld + d];
      const float nq = __ldg(qnorm + o);
      const bool qbig = nq >= 1e-12f;
      const float iq = 1.0f / fmaxf(nq, 1e-12f);
      const float* G = grad_logits + static_cast<int64_t>(o) * W;
      const float* Cv = logits + static_cast<int64_t>(o) * W;
      float gq = 0.f;
#pragma unroll
      for (int w = 0; w < WT; ++w) {
        if (w < W) {
          const float gw = __ldg(G + w);
          const float c = __ldg(Cv + w);
          const float s = gw * iq * ip[w];
          gq = fmaf(s, p[w], gq);
          if (qbig) gq = fmaf(-gw * c * iq * iq, q, gq);
          dp[w] = fmaf(s, q, dp[w]);
          if (pbig[w]) dp[w] = fmaf(-gw * c * ip[w] * ip[w], p[w], dp[w]);
        }
      }
      grad_feat[static_cast<int64_t>(row) * ldg + d] = gq;
    }
  }
#pragma unroll
  for (int w = 0; w < WT; ++w) {
    if (w < W) {
      const int row0 = cls_row[e * W + w];
      const float v = dp[w] / fS;
      for (int s = 0; s < S; ++s) grad_feat[static_cast<int64_t>(row0 + s) * ldg + d] = v;
    }
  }
}

template <int MODE>
int launch_bwd(const float* feat, int64_t ld, const int32_t* cls_row, int E, int W, int S, int D,
               const float* grad_logits, float* grad_feat, int64_t ldg, cudaStream_t stream) {
  dim3 grid(E, (D + 127) / 128);
  if (W <= 8)
    proto_bwd_kernel<MODE, 8><<<grid, 128, 0, stream>>>(feat, ld, cls_row, W, S, D, grad_logits, grad_feat, ldg);
  else
    proto_bwd_kernel<MODE, 32><<<grid, 128, 0, stream>>>(feat, ld, cls_row, W, S, D, grad_logits, grad_feat, ldg);
  AFS_LAUNCH_CHECK();
  return AFS_OK;
}

bool args_ok(const void* feat, int64_t ld, const void* cls_row, int N, int E, int W, int S, int D) {
  return feat != nullptr && cls_row != nullptr && E >= 0 && W >= 1 && W <= kMaxWay && S >= 1 &&
         D >= 4 && (D % 4) == 0 && ld >= D && (ld % 4) == 0 && N >= E * W * S &&
         (reinterpret_cast<uintptr_t>(feat) % 16) == 0;
}

}  // namespace
}  // namespace afs

extern "C" size_t afs_proto_workspace_bytes(int32_t E, int32_t W, int32_t S, int32_t D) {
  (void)S;
  if (E <= 0 || W <= 0 || D <= 0) return 0;
  // the E*W prototypes [D] and their inverse norms (cosine mode), 16-byte aligned
  const size_t protos = static_cast<size_t>(E) * W * D * sizeof(float);
  return protos + ((static_cast<size_t>(E) * W * sizeof(float) + 15) & ~static_cast<size_t>(15));
}

extern "C" int afs_proto_fwd(const float* feat, int64_t ld_feat, const int32_t* cls_row, int32_t N,
                             int32_t E, int32_t W, int32_t S, int32_t D, int32_t mode,
                             float* logits, int32_t* pred, void* ws, size_t ws_bytes,
                             afs_stream_t stream_) {
  using namespace afs;
  if (!args_ok(feat, ld_feat, cls_row, N, E, W, S, D) || logits == nullptr) return AFS_ERR_INVALID_ARG;
  if (E == 0 || N == E * W * S) return AFS_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t need = afs_proto_workspace_bytes(E, W, S, D);
  if (ws == nullptr || ws_bytes < need) return AFS_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(ws) % 16 != 0) return AFS_ERR_INVALID_ARG;
  float4* protos = static_cast<float4*>(ws);
  float* pinv = reinterpret_cast<float*>(static_cast<char*>(ws) + static_cast<size_t>(E) * W * D * sizeof(float));
  const int64_t total = static_cast<int64_t>(E) * W * (D / 4);
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  proto_mean_kernel<<<blocks, 256, 0, stream>>>(feat, ld_feat, cls_row, E * W, S, D / 4, protos);
  AFS_LAUNCH_CHECK();
  if (mode == AFS_PROTO_COSINE) {
    proto_pinv_kernel<<<(E * W * 32 + 255) / 256, 256, 0, stream>>>(protos, E * W, D / 4, pinv);
    AFS_LAUNCH_CHECK();
  }
  return dispatch_mode(mode, feat, ld_feat, cls_row, N, E, W, S, D, logits, pred, protos, pinv, stream);
}

extern "C" size_t afs_proto_bwd_cos_workspace_bytes(int32_t N, int32_t E, int32_t W, int32_t S) {
  if (N <= 0 || E <= 0 || W <= 0 || S <= 0) return 0;
  const int64_t nq = static_cast<int64_t>(N) - static_cast<int64_t>(E) * W * S;
  return nq < 0 ? 0 : (static_cast<size_t>(nq) + static_cast<size_t>(E) * W) * sizeof(float);
}

extern "C" int afs_proto_bwd_cos(const float* feat, int64_t ld_feat, const int32_t* cls_row, int32_t N, int32_t E,
                                 int32_t W, int32_t S, int32_t D, const float* logits, const float* grad_logits,
                                 float* grad_feat, int64_t ld_grad, void* ws, size_t ws_bytes,
                                 afs_stream_t stream_) {
  using namespace afs;
  if (!args_ok(feat, ld_feat, cls_row, N, E, W, S, D) || logits == nullptr || grad_logits == nullptr ||
      grad_feat == nullptr || ld_grad < D)
    return AFS_ERR_INVALID_ARG;
  if (E == 0) return AFS_OK;
  const int NQ = N - E * W * S;
  if (ws == nullptr || ws_bytes < afs_proto_bwd_cos_workspace_bytes(N, E, W, S)) return AFS_ERR_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float* qnorm = static_cast<float*>(ws);
  float* pnorm = qnorm + NQ;
  const int64_t warps = static_cast<int64_t>(NQ) + static_cast<int64_t>(E) * W;
  proto_norms_kernel<<<static_cast<unsigned>((warps * 32 + 255) / 256), 256, 0, stream>>>(feat, ld_feat, cls_row, E * W, W,
                                                                                         S, D, NQ, qnorm, pnorm);
  AFS_LAUNCH_CHECK();
  dim3 grid(E, (D + 127) / 128);
  if (W <= 8)
    proto_bwd_cos_kernel<8><<<grid, 128, 0, stream>>>(feat, ld_feat, cls_row, W, S, D, grad_logits, logits, qnorm, pnorm,
                                                      grad_feat, ld_grad);
  else
    proto_bwd_cos_kernel<32><<<grid, 128, 0, stream>>>(feat, ld_feat, cls_row, W, S, D, grad_logits, logits, qnorm, pnorm,
                                                       grad_feat, ld_grad);
  AFS_LAUNCH_CHECK();
  return AFS_OK;
}

extern "C" int afs_proto_bwd(const float* feat, int64_t ld_feat, const int32_t* cls_row, int32_t N,
                             int32_t E, int32_t W, int32_t S, int32_t D, int32_t mode,
                             const float* grad_logits, float* grad_feat, int64_t ld_grad,
                             afs_stream_t stream_) {
  using namespace afs;
  if (!args_ok(feat, ld_feat, cls_row, N, E, W, S, D) || grad_logits == nullptr ||
      grad_feat == nullptr || ld_grad < D)
    return AFS_ERR_INVALID_ARG;
  if (E == 0) return AFS_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  switch (mode) {
    case AFS_PROTO_EUCLIDEAN:
      return launch_bwd<AFS_PROTO_EUCLIDEAN>(feat, ld_feat, cls_row, E, W, S, D, grad_logits, grad_feat, ld_grad, stream);
    case AFS_PROTO_DOT:
      return launch_bwd<AFS_PROTO_DOT>(feat, ld_feat, cls_row, E, W, S, D, grad_logits, grad_feat, ld_grad, stream);
    default:
      return AFS_ERR_UNSUPPORTED;
  }
}
