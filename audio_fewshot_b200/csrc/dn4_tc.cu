// DN4 local-descriptor head on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same arithmetic as csrc/dn4.cu (reference DN4Layer.forward, libfewshot_core/model/metric/
// dn4.py:52-73) with the cosine relation computed as a TF32 tensor-core GEMM:
//   relation[128 query descriptors x 128 support descriptors] = Q_hat . S_hat^T   (K = C channels)
// is ONE chain of C/8 `tcgen05.mma.cta_group::1.kind::tf32` instructions issued by a single thread,
// accumulating in tensor memory; the 128 threads of the CTA then read their own accumulator row
// straight out of TMEM (`tcgen05.ld.32x32b`) and keep a running top-n_k in registers, so the relation
// tensor exists neither in HBM (the reference round-trips 3 MB per episode, dn4.py:71-72) nor in
// shared memory.  L2-normalisation (F.normalize, eps 1e-12) is applied while the operand tiles are
// written to shared memory, so there is no normalised copy in HBM either: feat is the only input.
//
// Operand tiles use the canonical K-major, no-swizzle UMMA layout: a tile of R rows (descriptors) by
// C channels is stored as [C/4 chunks][R rows][4 floats]: 8 consecutive rows of one 16-byte chunk form
// a contiguous 128-byte core matrix; SBO (next 8 rows) = 128 B, LBO (next 16-byte K chunk) = 16*R B.
// A thread writes whole 16-byte chunks of its own row (conflict-free 128-bit stores).
//
// Numerics: TF32 operands (10-bit mantissa), fp32 accumulation: scores within ~1e-4 relative of the
// fp32 path; top-k INDICES can differ from the fp32 path at near-ties, which is why csrc/dn4.cu stays
// the parity path and this kernel is the opt-in throughput path (north star: tensor-core path stated
// separately).  C % 8 == 0 and C <= 128 are built (Conv64F maps); wider maps use csrc/dn4.cu.
#include <limits.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "topk.cuh"

namespace afs {
namespace {

using namespace tc;
using namespace topk;

constexpr int kTcThreads = 128;
constexpr int kTcRows = 128;   // UMMA M: query descriptors per tile == TMEM lanes == threads
constexpr int kTcCols = 128;   // UMMA N: support descriptors per column tile == TMEM columns
constexpr int kTcMaxC = 128;
constexpr int kTcMaxWay = 32;

// Instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 (1 << 4), A and B TF32 (2 << 7, 2 << 10),
// both K-major (bits 15, 16 clear), N >> 3 in [17,23), M >> 4 in [24,29).
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((kTcCols >> 3) << 17) | ((kTcRows >> 4) << 24);

// One thread writes one L2-normalised, TF32-rounded descriptor (C channels of feat[row, :, m], channel
// stride HW) into row r of a K-major operand tile; a null src writes zeros.  Loads are issued 16 at a
// time so a row costs C/16 global-memory latencies, not C/4.
__device__ __forceinline__ void fill_row(float* tile, int r, const float* __restrict__ src, int C, int HW) {
  float4* dst = reinterpret_cast<float4*>(tile) + r;
  if (src == nullptr) {
    for (int c4 = 0; c4 < C / 4; ++c4) dst[c4 * kTcRows] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  float ss = 0.f;
  int c4 = 0;
  for (; c4 + 4 <= C / 4; c4 += 4) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float* p = src + static_cast<int64_t>(4 * (c4 + u)) * HW;
      v[u].x = __ldg(p);
      v[u].y = __ldg(p + HW);
      v[u].z = __ldg(p + 2 * HW);
      v[u].w = __ldg(p + 3 * HW);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ss = fmaf(v[u].x, v[u].x, ss); ss = fmaf(v[u].y, v[u].y, ss);
      ss = fmaf(v[u].z, v[u].z, ss); ss = fmaf(v[u].w, v[u].w, ss);
      dst[(c4 + u) * kTcRows] = v[u];
    }
  }
  for (; c4 < C / 4; ++c4) {
    const float* p = src + static_cast<int64_t>(4 * c4) * HW;
    const float4 v = make_float4(__ldg(p), __ldg(p + HW), __ldg(p + 2 * HW), __ldg(p + 3 * HW));
    ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
    dst[c4 * kTcRows] = v;
  }
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize: x / max(||x||, eps)
  for (c4 = 0; c4 < C / 4; ++c4) {
    float4 v = dst[c4 * kTcRows];
    v.x = to_tf32(v.x * inv); v.y = to_tf32(v.y * inv); v.z = to_tf32(v.z * inv); v.w = to_tf32(v.w * inv);
    dst[c4 * kTcRows] = v;
  }
}

template <int NK>
__global__ void __launch_bounds__(kTcThreads)
dn4_tc_kernel(const float* __restrict__ feat, const int32_t* __restrict__ cls_row, int W, int S, int C, int HW,
              float* __restrict__ rowsum, int32_t* __restrict__ topk_idx) {
  extern __shared__ __align__(128) float s_dyn[];
  float* sA = s_dyn;                   // [C/4][128][4]
  float* sB = s_dyn + kTcRows * C;     // [C/4][128][4]
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  __shared__ int s_qbase[kTcMaxWay + 1];

  const int e = blockIdx.z;
  const int tid = threadIdx.x;
  const int warp = tid >> 5;

  if (tid <= W) {
    const int g = e * W + tid;
    s_qbase[tid] = cls_row[g] - g * S;
  }
  if (warp == 0) {  // one warp allocates 128 TMEM columns (a 128 x 128 fp32 accumulator)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                 "r"(static_cast<uint32_t>(kTcCols)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(1u));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem;
  const uint32_t bar = smem_u32(&s_bar);
  uint32_t phase = 0;

  const int out0 = s_qbase[0];
  const int out1 = s_qbase[W];
  const int NS = S * HW;
  const int n_rows = (out1 - out0) * HW;
  const int n_tiles = (n_rows + kTcRows - 1) / kTcRows;
  const int n_ctiles = (NS + kTcCols - 1) / kTcCols;
  const uint32_t lbo = 16u * kTcRows, sbo = 128u;
  const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
  // this thread's accumulator row: TMEM lane == tid (warp w may only touch lanes 32w .. 32w+31)
  const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // ---- this thread's query descriptor (row tid of the tile)
    const int dr = tile * kTcRows + tid;
    const bool live = dr < n_rows;
    int o = 0, m = 0;
    const float* q_src = nullptr;
    if (live) {
      o = out0 + dr / HW;
      m = dr - (dr / HW) * HW;
      int cw = 0;
      while (s_qbase[cw + 1] <= o) ++cw;
      q_src = feat + (static_cast<int64_t>(o) + static_cast<int64_t>(e * W + cw + 1) * S) * C * HW + m;
    }
    fill_row(sA, tid, q_src, C, HW);

    for (int w = 0; w < W; ++w) {
      const int64_t sup_row0 = cls_row[e * W + w];
      float tv[NK];
      int ti[NK];
#pragma unroll
      for (int k = 0; k < NK; ++k) { tv[k] = -INFINITY; ti[k] = INT_MAX; }

      for (int ct = 0; ct < n_ctiles; ++ct) {
        const int col = ct * kTcCols + tid;
        const float* s_src = nullptr;
        if (col < NS) {
          const int s = col / HW;
          s_src = feat + (sup_row0 + s) * C * HW + (col - s * HW);
        }
        fill_row(sB, tid, s_src, C, HW);
        // generic-proxy stores -> visible to the tensor core's async proxy; earlier TMEM reads ordered
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          for (int k8 = 0; k8 < C / 8; ++k8) {
            const uint64_t da = desc_kmajor_noswizzle(a_addr + static_cast<uint32_t>(k8) * 2u * lbo, lbo, sbo);
            const uint64_t db = desc_kmajor_noswizzle(b_addr + static_cast<uint32_t>(k8) * 2u * lbo, lbo, sbo);
            const uint32_t accumulate = k8 > 0 ? 1u : 0u;
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
                "}\n" ::"r"(tmem_base), "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
                : "memory");
          }
          // arrives on the mbarrier when every MMA above has completed (implies fence::before_thread_sync)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                       : "memory");
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // ---- epilogue: my row of the 128 x 128 relation tile -> running top-NK (descending, lowest column on ties)
        uint32_t tk[NK];
#pragma unroll
        for (int k = 0; k < NK; ++k) tk[k] = 0u;  // below every real key (keys of finite values are > 0)
        const int valid = NS - ct * kTcCols;   // columns of this tile that exist (>= kTcCols for full tiles)
#pragma unroll 1
        for (int c0 = 0; c0 < kTcCols; c0 += 32) {
          uint32_t v[32];
          if (c0 >= valid) break;  // tile-uniform: nothing left in this column tile
          tmem_ld32(t_row + static_cast<uint32_t>(c0), v);
          if (c0 + 32 <= valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) topk_push<NK>(tk, topk_key(v[j], c0 + j));
          } else {
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {  // boundary chunk: whole groups of 8 are skipped by a uniform branch
              if (c0 + 8 * g8 < valid) {
#pragma unroll
                for (int j = 8 * g8; j < 8 * g8 + 8; ++j)
                  if (c0 + j < valid) topk_push<NK>(tk, topk_key(v[j], c0 + j));
              }
            }
          }
        }
#pragma unroll
        for (int k = 0; k < NK; ++k)
          if (tk[k] != 0u) topk_merge<NK>(tv, ti, topk_key_value(tk[k]), ct * kTcCols + 127 - static_cast<int>(tk[k] & 127u));
      }
      if (live) {
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < NK; ++k) sum += tv[k];
        const int64_t base = (static_cast<int64_t>(o) * W + w) * HW + m;
        rowsum[base] = sum;
        if (topk_idx != nullptr) {
#pragma unroll
          for (int k = 0; k < NK; ++k) topk_idx[base * NK + k] = ti[k];
        }
      }
    }
    // the next tile rewrites sA: every MMA that read it has completed (barrier waited above)
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(kTcCols)));
  }
}

// sum over the HW descriptors of a query (fixed order) + argmax; same as csrc/dn4.cu
__global__ void __launch_bounds__(128)
dn4_tc_reduce_kernel(const float* __restrict__ rowsum, int NQ, int W, int HW, float* __restrict__ score,
                     int32_t* __restrict__ pred) {
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

template <int NK>
cudaError_t launch_tc(dim3 grid, size_t smem, cudaStream_t stream, const float* feat, const int32_t* cls_row, int W,
                      int S, int C, int HW, float* rowsum, int32_t* topk_idx) {
  cudaError_t e = cudaFuncSetAttribute(dn4_tc_kernel<NK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  dn4_tc_kernel<NK><<<grid, kTcThreads, smem, stream>>>(feat, cls_row, W, S, C, HW, rowsum, topk_idx);
  return cudaSuccess;
}

}  // namespace
}  // namespace afs

extern "C" size_t afs_dn4_tc_workspace_bytes(int32_t N, int32_t E, int32_t W, int32_t S, int32_t HW) {
  if (N <= 0 || E <= 0 || W <= 0 || S <= 0 || HW <= 0) return 0;
  const int64_t nq = static_cast<int64_t>(N) - static_cast<int64_t>(E) * W * S;
  return nq <= 0 ? 0 : static_cast<size_t>(nq) * W * HW * sizeof(float);
}

extern "C" int afs_dn4_fwd_tc(const float* feat, const int32_t* cls_row, int32_t N, int32_t E, int32_t W,
                              int32_t S, int32_t C, int32_t HW, int32_t n_k, float* score, int32_t* topk_idx,
                              int32_t* pred, void* ws, size_t ws_bytes, afs_stream_t stream_) {
  using namespace afs;
  if (feat == nullptr || cls_row == nullptr || score == nullptr || E < 0 || W < 1 || W > kTcMaxWay || S < 1 ||
      C < 1 || HW < 1 || N < E * W * S || n_k < 1 || n_k > 8 || n_k > S * HW)
    return AFS_ERR_INVALID_ARG;
  if ((C & 7) != 0 || C > kTcMaxC) return AFS_ERR_UNSUPPORTED;  // wider maps: afs_dn4_fwd (fp32 path)
  const int NQ = N - E * W * S;
  if (E == 0 || NQ == 0) return AFS_OK;
  if (E > 65535) return AFS_ERR_UNSUPPORTED;
  if (ws == nullptr || ws_bytes < afs_dn4_tc_workspace_bytes(N, E, W, S, HW)) return AFS_ERR_WORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float* rowsum = static_cast<float*>(ws);
  const int64_t avg_rows = static_cast<int64_t>(NQ) * HW / E;
  int tiles = static_cast<int>((avg_rows + kTcRows - 1) / kTcRows);
  if (tiles < 1) tiles = 1;
  const dim3 grid(tiles, 1, E);
  const size_t smem = 2 * static_cast<size_t>(kTcRows) * C * sizeof(float);
  cudaError_t err = cudaSuccess;
  switch (n_k) {
    case 1: err = launch_tc<1>(grid, smem, stream, feat, cls_row, W, S, C, HW, rowsum, topk_idx); break;
    case 2: err = launch_tc<2>(grid, smem, stream, feat, cls_row, W, S, C, HW, rowsum, topk_idx); break;
    case 3: err = launch_tc<3>(grid, smem, stream, feat, cls_row, W, S, C, HW, rowsum, topk_idx); break;
    case 4: err = launch_tc<4>(grid, smem, stream, feat, cls_row, W, S, C, HW, rowsum, topk_idx); break;
    case 5: err = launch_tc<5>(grid, smem, stream, feat, cls_row, W, S, C, HW, rowsum, topk_idx); break;
    case 6: err = launch_tc<6>(grid, smem, stream, feat, cls_row, W, S, C, HW, rowsum, topk_idx); break;
    case 7: err = launch_tc<7>(grid, smem, stream, feat, cls_row, W, S, C, HW, rowsum, topk_idx); break;
    default: err = launch_tc<8>(grid, smem, stream, feat, cls_row, W, S, C, HW, rowsum, topk_idx); break;
  }
  if (err != cudaSuccess) return cuda_fail(err);
  AFS_LAUNCH_CHECK();
  dn4_tc_reduce_kernel<<<(NQ + 127) / 128, 128, 0, stream>>>(rowsum, NQ, W, HW, score, pred);
  AFS_LAUNCH_CHECK();
  return AFS_OK;
}
