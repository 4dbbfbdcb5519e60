// First Conv64F block on the tensor cores: Conv2d(1->64, 3x3, pad 1) + BatchNorm(eval) + ReLU/LeakyReLU +
// MaxPool2d(3,3) as tcgen05 TF32 MMAs with the max-pool done on the accumulators in tensor memory.  sm_100a.
//
// Same operator as csrc/conv1.cu (reference libfewshot_core/model/backbone/conv_four.py:61-66,101-103);
// conv1.cu is the exact-fp32 SIMT kernel and is FMA-pipe bound (9 conv positions x 9 taps x 64 channels per
// pooled pixel: 0.37 ms per 800 clips, 69 % of the fp32 FMA peak).  The reference runs its convolutions in
// TF32 by default (torch.backends.cudnn.allow_tf32 = True), and under that setting this kernel is used.
//
// Formulation: a tile is 128 POOLED pixels = 128 accumulator rows = 128 TMEM lanes = 128 threads.  For each of
// the 9 positions d of the pooling window one GEMM  D_d[128 x 64] = P_d[128 x 16] . Wf^T[16 x 64]  (K = 9 taps
// zero-padded to 16) gives the conv output at that window position for all 64 channels; the max over d is an
// elementwise max over nine accumulators that all sit in the SAME lane, so the pooling needs no cross-thread
// traffic at all.  The three windows of one window row share a shared-memory buffer and 192 TMEM columns; the
// im2col rows P_d are written by the thread that owns the pooled pixel straight from its 5x5 register patch
// (K-major no-swizzle UMMA layout, conflict-free 128-bit stores), double-buffered so that building the next
// window row overlaps the MMAs of the current one.  Epilogue: tcgen05.ld, running max, + folded shift,
// activation, 256-byte NHWC row store.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace afs {
namespace {

using namespace tc;

constexpr int kPix1 = 128;                 // pooled pixels per tile == threads
constexpr int kCh1 = 64;                   // output channels == UMMA N
constexpr uint32_t kWinBytes = 4u * kPix1 * 16u;      // one window's P_d: [4 chunks][128 rows][16 B] = 8 KB
constexpr uint32_t kBufBytes = 3u * kWinBytes;        // one window row: 24 KB
constexpr uint32_t kWBytes = 4u * kCh1 * 16u;         // Wf: [4 chunks][64 rows][16 B] = 4 KB

struct Conv1TcParams {
  float w[kCh1 * 9];  // folded weights [c][ky][kx]
  float shift[kCh1];
};

__global__ void __launch_bounds__(kPix1)
conv1_tc_kernel(const float* __restrict__ x, int64_t total_pix, int H, int Wd, int PH, int PW, float slope,
                float* __restrict__ out, const __grid_constant__ Conv1TcParams prm) {
  extern __shared__ __align__(128) uint8_t s_buf[];  // [2][kBufBytes] im2col | [kWBytes] weights
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const uint32_t a_base = smem_u32(s_buf);
  const uint32_t w_base = a_base + 2u * kBufBytes;

  constexpr uint32_t kTmemCols1 = 256;  // 3 accumulators x 64 columns, power of two
  // folded weights -> K-major operand [chunk][channel][4 taps], TF32-rounded; taps 9..15 are zero
  for (int i = tid; i < 4 * kCh1; i += kPix1) {
    const int chunk = i / kCh1, c = i - chunk * kCh1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (chunk == 0) v = make_float4(to_tf32(prm.w[c * 9 + 0]), to_tf32(prm.w[c * 9 + 1]), to_tf32(prm.w[c * 9 + 2]), to_tf32(prm.w[c * 9 + 3]));
    if (chunk == 1) v = make_float4(to_tf32(prm.w[c * 9 + 4]), to_tf32(prm.w[c * 9 + 5]), to_tf32(prm.w[c * 9 + 6]), to_tf32(prm.w[c * 9 + 7]));
    if (chunk == 2) v.x = to_tf32(prm.w[c * 9 + 8]);
    reinterpret_cast<float4*>(s_buf + 2u * kBufBytes)[i] = v;
  }
  // the all-zero fourth chunk of every window tile never changes
  for (int i = tid; i < 2 * 3 * kPix1; i += kPix1) {
    const int win = i / kPix1, r = i - win * kPix1;
    reinterpret_cast<float4*>(s_buf + win * kWinBytes)[3 * kPix1 + r] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (warp == 0) tmem_alloc(&s_tmem, kTmemCols1);
  if (tid == 0) {
    mbar_init(smem_u32(&s_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = s_tmem;
  const uint32_t bar = smem_u32(&s_bar);
  const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  constexpr uint32_t kIdesc = idesc_tf32(kPix1, kCh1);

  const int64_t n_tiles = (total_pix + kPix1 - 1) / kPix1;
  const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t n_steps = 3 * my_tiles;

  float p[5][5];   // this thread's input patch (tile being built)
  float pn[5][5];  // patch of the NEXT tile, loaded two steps ahead so its global latency is off the critical path
  float best[kCh1];

  auto load_patch = [&](int64_t tile_iter) {
    const int64_t tile = blockIdx.x + tile_iter * gridDim.x;
    int64_t pix = tile * kPix1 + tid;
    if (pix >= total_pix) pix = total_pix - 1;  // tail rows recompute the last pixel; never stored
    const int per = PH * PW;
    const int64_t n = pix / per;
    const int r = static_cast<int>(pix - n * per);
    const int ph = r / PW, pw = r - ph * PW;
    const float* img = x + n * static_cast<int64_t>(H) * Wd;
    const int y0 = 3 * ph - 1, x0 = 3 * pw - 1;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int yy = y0 + i;
      const bool yin = (yy >= 0) && (yy < H);
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int xx = x0 + j;
        pn[i][j] = (yin && xx >= 0 && xx < Wd) ? __ldg(img + static_cast<int64_t>(yy) * Wd + xx) : 0.f;
      }
    }
  };

  // writes the three im2col rows of window row g (positions (g,0..2)) of this thread's pooled pixel
  auto build = [&](int64_t step) {
    const int g = static_cast<int>(step % 3);
    if (g == 0) {  // the patch was requested two steps ago (or in the prologue)
#pragma unroll
      for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) p[i][j] = to_tf32(pn[i][j]);
    }
    if (g == 1 && step / 3 + 1 < my_tiles) load_patch(step / 3 + 1);
    float4* dst = reinterpret_cast<float4*>(s_buf + (step & 1) * kBufBytes) + tid;
#pragma unroll
    for (int gg = 0; gg < 3; ++gg) {
      if (gg == g) {  // resolved at compile time inside each unrolled copy: p[] keeps static indices
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          float4* row = dst + dx * (4 * kPix1);
          row[0 * kPix1] = make_float4(p[gg][dx], p[gg][dx + 1], p[gg][dx + 2], p[gg + 1][dx]);
          row[1 * kPix1] = make_float4(p[gg + 1][dx + 1], p[gg + 1][dx + 2], p[gg + 2][dx], p[gg + 2][dx + 1]);
          row[2 * kPix1] = make_float4(p[gg + 2][dx + 2], 0.f, 0.f, 0.f);
        }
      }
    }
  };

  uint32_t phase = 0;
  if (n_steps > 0) {
    load_patch(0);
    build(0);
  }
  for (int64_t step = 0; step < n_steps; ++step) {
    fence_async_smem();   // im2col rows of this step -> visible to the tensor core
    fence_before();       // the previous step's TMEM reads are ordered before the barrier
    __syncthreads();
    if (tid == 0) {
      fence_after();
      const uint32_t buf = a_base + static_cast<uint32_t>(step & 1) * kBufBytes;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
        for (int k8 = 0; k8 < 2; ++k8) {
          const uint64_t da = desc_kmajor_noswizzle(buf + dx * kWinBytes + k8 * 2u * (16u * kPix1), 16u * kPix1, 128u);
          const uint64_t db = desc_kmajor_noswizzle(w_base + k8 * 2u * (16u * kCh1), 16u * kCh1, 128u);
          mma_tf32(tmem_base + dx * kCh1, da, db, kIdesc, k8 > 0);
        }
      }
      commit(bar);
    }
    if (step + 1 < n_steps) build(step + 1);  // overlaps the MMAs just issued (other buffer)
    mbar_wait(bar, phase);
    phase ^= 1u;
    fence_after();

    const int g = static_cast<int>(step % 3);
#pragma unroll
    for (int half = 0; half < kCh1 / 32; ++half) {  // 32 channels at a time: three windows x 32 columns, loads batched
      uint32_t v0[32], v1[32];
      tmem_ld32_nowait(t_row + static_cast<uint32_t>(0 * kCh1 + half * 32), v0);
      tmem_ld32_nowait(t_row + static_cast<uint32_t>(1 * kCh1 + half * 32), v1);
      tmem_wait_ld();
      if (g == 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) best[half * 32 + j] = fmaxf(__uint_as_float(v0[j]), __uint_as_float(v1[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          best[half * 32 + j] = fmaxf(best[half * 32 + j], fmaxf(__uint_as_float(v0[j]), __uint_as_float(v1[j])));
      }
      tmem_ld32(t_row + static_cast<uint32_t>(2 * kCh1 + half * 32), v0);
#pragma unroll
      for (int j = 0; j < 32; ++j) best[half * 32 + j] = fmaxf(best[half * 32 + j], __uint_as_float(v0[j]));
    }
    if (g == 2) {
      const int64_t tile = blockIdx.x + (step / 3) * gridDim.x;
      const int64_t pix = tile * kPix1 + tid;
      if (pix < total_pix) {
        float4* o = reinterpret_cast<float4*>(out + pix * kCh1);
#pragma unroll
        for (int c4 = 0; c4 < kCh1 / 4; ++c4) {
          float4 r;
          r.x = best[4 * c4 + 0] + prm.shift[4 * c4 + 0];
          r.y = best[4 * c4 + 1] + prm.shift[4 * c4 + 1];
          r.z = best[4 * c4 + 2] + prm.shift[4 * c4 + 2];
          r.w = best[4 * c4 + 3] + prm.shift[4 * c4 + 3];
          r.x = r.x > 0.f ? r.x : r.x * slope;
          r.y = r.y > 0.f ? r.y : r.y * slope;
          r.z = r.z > 0.f ? r.z : r.z * slope;
          r.w = r.w > 0.f ? r.w : r.w * slope;
          o[c4] = r;
        }
      }
    }
  }

  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols1);
}

}  // namespace
}  // namespace afs

extern "C" int afs_conv1_bn_act_pool3_fwd_tf32(const float* x, int32_t N, int32_t H, int32_t Wd,
                                               const float* w_folded_host, const float* shift_host, int32_t C,
                                               float negative_slope, float* out, afs_stream_t stream_) {
  using namespace afs;
  if (x == nullptr || w_folded_host == nullptr || shift_host == nullptr || out == nullptr || N < 0 || H < 3 ||
      Wd < 3 || negative_slope < 0.f)
    return AFS_ERR_INVALID_ARG;
  if (C != kCh1) return AFS_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(out) & 15) != 0) return AFS_ERR_INVALID_ARG;
  if (N == 0) return AFS_OK;
  const int PH = H / 3, PW = Wd / 3;
  const int64_t total = static_cast<int64_t>(N) * PH * PW;
  const int64_t n_tiles = (total + kPix1 - 1) / kPix1;
  Conv1TcParams prm;
  for (int i = 0; i < kCh1 * 9; ++i) prm.w[i] = w_folded_host[i];
  for (int i = 0; i < kCh1; ++i) prm.shift[i] = shift_host[i];
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t smem = 2 * kBufBytes + kWBytes;
  int64_t blocks = n_tiles < 2 * kNumSMs ? n_tiles : 2 * kNumSMs;  // persistent: 2 CTAs per SM (256 TMEM columns each)
  AFS_CUDA_TRY(cudaFuncSetAttribute(conv1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  conv1_tc_kernel<<<static_cast<unsigned>(blocks), kPix1, smem, stream>>>(x, total, H, Wd, PH, PW, negative_slope, out, prm);
  AFS_LAUNCH_CHECK();
  return AFS_OK;
}
