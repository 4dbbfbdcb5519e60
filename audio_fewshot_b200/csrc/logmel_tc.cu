// Tensor-core engine of the fused waveform -> log-mel front-end (sm_100a): the 1024-point real DFT of every frame
// as a FOUR-STEP 32 x 32 decomposition whose two DFT stages are tcgen05 GEMMs with fp32 accumulators in TMEM.
//
//   n = n1 + 32 n2,  k = k1 + 32 k2,  W_N = exp(-2 pi i / N),  xw = window * frame
//   step 1 (GEMM):  Y[n1, k1] = sum_n2 xw[n1 + 32 n2] W_32^(n2 k1)          real input: k1 = 0..16 is enough,
//                   rows (frame, n1) x K = n2 (32) x N = 32 columns {Y0, Y16, Re/Im Y1..15}
//   step 2 (SIMT):  U[n1, k1] = Y[n1, k1] W_1024^(n1 k1) / 32               in the accumulator's owner thread
//   step 3 (GEMM):  X[k1 + 32 k2] = sum_n1 U[n1, k1] W_32^(n1 k2)           rows (frame, k1 = 0..15) x K = (Re|Im, n1)
//                   (64) x N = (k2, Re|Im) (64); row k1 holds bins k1 + 32 k2 and, mirrored, 1024 - k1 - 32 k2
//   step 3s (GEMM): X[16 + 32 k2] = sum_n1 Y[n1, 16]/32 W_64^(n1 (2 k2 + 1))  the k1 = 16 rows (real before the
//                   twiddle), one row per frame x K = n1 (32) x N = 32
//   then |X|^2 -> banded mel projection -> log -> (x - mean)/std as in the FFT engine (logmel.cu).
//
// Precision: operands are fp16 PAIRS (hi, lo = x - hi), every product is hi*hi + hi*lo + lo*hi accumulated in fp32
// (the dropped lo*lo term is 2^-22 relative).  Samples of a chunk are scaled by a power of two so that the largest
// windowed sample sits in [2^13, 2^14): hi and lo are then normal fp16 numbers whatever the signal level, and the
// factor is undone exactly on the mel energies.  Measured on B200 (tools/tc_dft_probe.cu): step-1 GEMM error
// 4e-7 of max|Y| (rms 6e-8) -- the level of an fp32 FFT.
//
// Why a GEMM: the radix-8 FFT engine is instruction-issue bound (~1 000 warp-instructions per frame); here the
// 2 x 10^5 multiply-adds per frame of the two DFT stages run on the tensor pipe and the SIMT lanes are left with
// operand conversion, the twiddle, |X|^2 and the mel projection.
//
// One persistent CTA per SM, 25 warps in five roles that work on DIFFERENT chunks (8 frames of one clip) at the same
// time, chained by mbarriers (every operand buffer and accumulator is double-buffered):
//   LD   warps 0-3    thread = 8-sample group of 4 frames (a tile): samples come from a 3-deep shared-memory ring that
//                     one thread fills two chunks ahead with cp.async.bulk (chunks that need reflect padding, the
//                     Philox augmentation or lack 16-byte alignment are loaded directly, as the FFT engine does);
//                     window, tile max -> power-of-two scale, fp16 hi/lo split, 16-byte stores into the A1 operand
//                     [frame][32 sample rows of 32][SWIZZLE_64B]: frame memory order IS the MN-major operand
//   MMA  warp 24      one thread: step-1 GEMMs of chunk i (2 tiles x 3 terms x 2 K-steps of M128 N32 K16), then the
//                     step-3 / 3s GEMMs of chunk i-1 (3 x 4 of M128 N64 K16 + 3 x 2 of N32); tcgen05.commit hands
//                     accumulators to the epilogues and operand buffers back to their producers
//   E1   warps 4-11   warp = frame, lane = n1: tcgen05.ld 32 columns, twiddle, fp16 split, 16-byte stores into the A3
//                     operand (MN-major: 8 consecutive k1 of one n1 are one store), Y16 into the A3s operand
//   E3   warps 12-15  thread = row (frame, k1): tcgen05.ld 64 columns, |X|^2 into the power buffer [frame][528]
//   MEL  warps 16-23  thread = half of a mel filter's band x 8 frames (two lanes per filter, halves added with one
//                     shuffle), log, normalise, [128 x 8] tile, store along time
#include <cuda_fp16.h>
#include <math.h>

#include <new>
#include <vector>

#include "common.cuh"
#include "logmel_core.cuh"
#include "logmel_fft.cuh"
#include "logmel_plan.cuh"
#include "tc_common.cuh"

namespace afs {
namespace logmel {
namespace {

using namespace tc;

constexpr int kWarpsLD = 4, kWarpsE1 = 8, kWarpsE3 = 4, kWarpsMEL = 8;
constexpr int kWarpE1 = kWarpsLD, kWarpE3 = kWarpE1 + kWarpsE1, kWarpMEL = kWarpE3 + kWarpsE3, kWarpMMA = kWarpMEL + kWarpsMEL;
constexpr int kTcThreads = 32 * (kWarpMMA + 1);  // 800
constexpr int kTcFrames = 8;         // frames per chunk
constexpr int kTcChunksPerItem = 5;  // consecutive chunks of one clip per work item (partial output sectors meet in L2)
constexpr int kPS = 528;             // power row stride in floats: == 16 mod 32, the |X|^2 stores are conflict-free
constexpr int kTileS = kTcFrames + 1;
constexpr int kRing = 16;            // per-chunk scale exponents travel LD -> MEL through a ring deeper than the pipeline
constexpr int kStages = 3;           // raw-sample ring
constexpr uint32_t kStageBytes = 18432;  // 7 * 512 + 1024 fp32 samples: a hop-512 chunk
constexpr int kMaxMelW = 1500;       // floats of the MEL role's weight table (128 slaney mels: ~1 200)

// shared-memory map, byte offsets from a 1024-aligned base
constexpr uint32_t kA1 = 0;                                     // 2 tile stages: hi [4 frames][2048 B] | lo
constexpr uint32_t kA1Stage = 16384, kA1Lo = 8192;
constexpr uint32_t kA3 = kA1 + 2 * kA1Stage;                    // x 2: hi [8 K groups][16 row groups][128 B] | lo
constexpr uint32_t kA3Stage = 32768, kA3Lo = 16384;
constexpr uint32_t kA3s = kA3 + 2 * kA3Stage;                   // x 2: hi [4 K groups][8 rows][16 B] | lo
constexpr uint32_t kA3sStage = 1024, kA3sLo = 512;
constexpr uint32_t kB1 = kA3s + 2 * kA3sStage;                  // hi 2048 | lo 2048   [4 K groups][32 rows][16 B]
constexpr uint32_t kB3 = kB1 + 4096;                            // hi 8192 | lo 8192   [8 K groups][64 rows][16 B]
constexpr uint32_t kB3s = kB3 + 16384;                          // hi 2048 | lo 2048   [4 K groups][32 rows][16 B]
constexpr uint32_t kTw = kB3s + 4096;                           // float2 [16][32]
constexpr uint32_t kConstBytes = kTw + 4096 - kB1;              // 28 672: one contiguous image built by the host
constexpr uint32_t kPow = kTw + 4096;                           // x 2: float [8][528]
constexpr uint32_t kPowStage = kPS * kTcFrames * 4;
constexpr uint32_t kTile = kPow + 2 * kPowStage;                // float [128][9]
constexpr uint32_t kRaw = kTile + kMaxMels * kTileS * 4;        // x 3: raw samples of a chunk
constexpr uint32_t kMBand = kRaw + kStages * kStageBytes;       // int2 [256]: first bin / count, weight offset
constexpr uint32_t kMW = kMBand + 256 * 8;                      // float [mw]
static_assert(kMW + kMaxMelW * 4 + 1024 + 512 <= 232448, "shared memory budget (227 KB, incl. alignment slack and statics)");

// tensor memory columns: accumulators x 2
constexpr uint32_t kAcc1 = 0, kAcc1Stage = 64;                  // 2 tiles x 32 columns
constexpr uint32_t kAcc3 = 128, kAcc3Stage = 96;                // step 3: 64 columns, step 3s: 32
constexpr uint32_t kTmemCols = 512;

// mbarriers
enum Bar {
  kA1Full = 0, kA1Empty = 2, kAcc1Full = 4, kAcc1Empty = 6, kA3Full = 8, kA3Empty = 10, kAcc3Full = 12, kAcc3Empty = 14,
  kPFull = 16, kPEmpty = 18, kRawFull = 20, kEsFull = 20 + kStages, kNumBars = 20 + kStages + kRing
};

struct TcArgs {
  const uint8_t* consts;  // kConstBytes
  const int2* mband;      // [256]
  const float* mw;        // [mw_count]
  int mw_count;
  int chunks8;            // ceil(T / 8)
  int groups;             // ceil(chunks8 / kTcChunksPerItem)
  int64_t total_bytes;    // B * L * sizeof(sample): bulk copies never read past it
};

struct TcTables {  // behind afs_logmel_plan::d_tc
  void* d_consts;
  int2* d_mband;
  float* d_mw;
  int mw_count;
};

// 8 consecutive samples s .. s + 7 of the (reflect-padded, optionally augmented) clip
template <bool AUG, typename S>
__device__ __forceinline__ void fetch8(const S* __restrict__ x, int64_t s, int64_t L, const AugState& a, float (&v)[8]) {
  const bool interior = s >= 0 && s + 8 <= L;
  if constexpr (!AUG) {
    if (interior) {
      const S* q = x + s;
      const uintptr_t addr = reinterpret_cast<uintptr_t>(q);
      if constexpr (sizeof(S) == 4) {
        if ((addr & 15u) == 0) {
          const float4 a0 = __ldg(reinterpret_cast<const float4*>(q)), a1 = __ldg(reinterpret_cast<const float4*>(q) + 1);
          v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
        } else if ((addr & 7u) == 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 t = ld_pair(q + 2 * i, a.pcm_scale);
            v[2 * i] = t.x; v[2 * i + 1] = t.y;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = ld_sample(q + i, a.pcm_scale);
        }
      } else {
        if ((addr & 3u) == 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 t = ld_pair(q + 2 * i, a.pcm_scale);
            v[2 * i] = t.x; v[2 * i + 1] = t.y;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = ld_sample(q + i, a.pcm_scale);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = ld_sample(x + reflect_index(s + i, L), a.pcm_scale);
    }
  } else {
    if (interior && (s & 1) == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 t = aug_pair(x, s + 2 * i, L, a);
        v[2 * i] = t.x; v[2 * i + 1] = t.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = aug_sample(x, reflect_index(s + i, L), L, a);
    }
  }
}

// (a, b) -> packed fp16 hi pair and lo pair (lo = value - hi, exact in fp32 before its own rounding)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// named barriers (bar.sync id, threads): 0 is __syncthreads
__device__ __forceinline__ void role_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// 8 consecutive samples of the raw ring (byte address `a`, aligned to `align` bytes at least)
template <typename S>
__device__ __forceinline__ void lds8(const uint8_t* a, int align, float pcm_scale, float (&v)[8]) {
  if constexpr (sizeof(S) == 4) {
    if (align >= 16) {
      const float4 x0 = *reinterpret_cast<const float4*>(a), x1 = *reinterpret_cast<const float4*>(a + 16);
      v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w; v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
    } else if (align >= 8) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 t = *reinterpret_cast<const float2*>(a + 8 * i);
        v[2 * i] = t.x; v[2 * i + 1] = t.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const float*>(a + 4 * i);
    }
  } else {
    if (align >= 16) {
      const uint4 x = *reinterpret_cast<const uint4*>(a);
      const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[2 * i] = __fmul_rn(static_cast<float>(static_cast<int16_t>(w[i] & 0xffffu)), pcm_scale);
        v[2 * i + 1] = __fmul_rn(static_cast<float>(static_cast<int16_t>(w[i] >> 16)), pcm_scale);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __fmul_rn(static_cast<float>(*reinterpret_cast<const int16_t*>(a + 2 * i)), pcm_scale);
    }
  }
}

// Development instrumentation (-DAFS_TC_PROFILE): per-role cycles spent waiting on mbarriers and working, CTA 0.
#ifdef AFS_TC_PROFILE
__device__ long long g_tc_prof[96];
#define TCP_DECL long long tp_w_ = 0, tp_k_ = 0, tp_t_ = clock64(), tp_s_[4] = {0, 0, 0, 0}
#define TCP_SEG(k) { const long long n_ = clock64(); tp_s_[k] += n_ - tp_t_; tp_k_ += n_ - tp_t_; tp_t_ = n_; }
#define TCP_STORE_SEG(slot, cond) if (blockIdx.x == 0 && (cond)) { for (int k_ = 0; k_ < 4; ++k_) g_tc_prof[(slot) + k_] = tp_s_[k_]; }
#define TCP_WAITED() { const long long n_ = clock64(); tp_w_ += n_ - tp_t_; tp_t_ = n_; }
#define TCP_WORKED() { const long long n_ = clock64(); tp_k_ += n_ - tp_t_; tp_t_ = n_; }
#define TCP_STORE(slot, cond, cnt) if (blockIdx.x == 0 && (cond)) { g_tc_prof[slot] = tp_w_; g_tc_prof[(slot) + 1] = tp_k_; g_tc_prof[(slot) + 2] = (cnt); }
#else
#define TCP_DECL
#define TCP_SEG(k)
#define TCP_STORE_SEG(slot, cond)
#define TCP_WAITED()
#define TCP_WORKED()
#define TCP_STORE(slot, cond, cnt)
#endif

template <bool AUG, typename S>
__global__ void __launch_bounds__(kTcThreads, 1) logmel_tc_kernel(const Params p, const TcArgs q) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_bars[kNumBars];
  __shared__ uint32_t s_tmem;
  __shared__ float s_red[2][kWarpsLD];
  __shared__ int s_es[kRing][2];  // scale exponents of the two 4-frame tiles of a chunk
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sb = smem_u32(sm);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const uint32_t bars = smem_u32(s_bars);
#define BAR(id) (bars + 8u * static_cast<uint32_t>(id))

  if (warp == 0) tmem_alloc(&s_tmem, kTmemCols);
  if (tid == 32) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(BAR(kA1Full + s), 32 * kWarpsLD);
      mbar_init(BAR(kA1Empty + s), 1);
      mbar_init(BAR(kAcc1Full + s), 1);
      mbar_init(BAR(kAcc1Empty + s), 32 * kWarpsE1);
      mbar_init(BAR(kA3Full + s), 32 * kWarpsE1);
      mbar_init(BAR(kA3Empty + s), 1);
      mbar_init(BAR(kAcc3Full + s), 1);
      mbar_init(BAR(kAcc3Empty + s), 32 * kWarpsE3);
      mbar_init(BAR(kPFull + s), 32 * kWarpsE3);
      mbar_init(BAR(kPEmpty + s), 32 * kWarpsMEL);
    }
    for (int s = 0; s < kStages; ++s) mbar_init(BAR(kRawFull + s), 1);
    for (int s = 0; s < kRing; ++s) mbar_init(BAR(kEsFull + s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < static_cast<int>(kConstBytes / 16); i += kTcThreads)
    reinterpret_cast<uint4*>(sm + kB1)[i] = __ldg(reinterpret_cast<const uint4*>(q.consts) + i);
  {
    int2* s_mband = reinterpret_cast<int2*>(sm + kMBand);
    float* s_mw = reinterpret_cast<float*>(sm + kMW);
    for (int i = tid; i < 256; i += kTcThreads) s_mband[i] = q.mband[i];
    for (int i = tid; i < q.mw_count; i += kTcThreads) s_mw[i] = q.mw[i];
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = s_tmem;
  const int n_items = p.B * q.groups;

  // Every role walks the same chunk sequence: items blockIdx.x, + gridDim.x, ...; chunks of an item in order.
  // `it` counts chunks: double-buffered stage = it & 1, barrier parity = (it >> 1) & 1.
  if (warp < kWarpE1) {
    // =========================================================== LD: waveform -> A1 operand
    const int c8 = tid & 127;  // 8-sample group of the frame: n = 8 c8 .. 8 c8 + 7
    float win[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) win[i] = __ldg(p.window + 8 * c8 + i);
    uint32_t sw_off = 16u * c8;
    sw_off ^= ((sw_off >> 7) & 3u) << 4;  // SWIZZLE_64B: 16-byte chunk index ^= bits 7..8 of the (1024-aligned) address
    // raw ring: a chunk whose 8 frames lie inside the clip and whose first sample is 16-byte aligned is fetched by ONE
    // bulk copy two chunks ahead; anything else (clip edges, augmentation) is loaded directly
    const int span = 7 * p.hop + kNfft;                                       // samples of a chunk
    const uint32_t span_bytes = (static_cast<uint32_t>(span) * sizeof(S) + 15u) & ~15u;
    const bool ring_ok = !AUG && static_cast<int64_t>(span) * static_cast<int64_t>(sizeof(S)) <= static_cast<int64_t>(kStageBytes);
    const int lds_align = (p.hop * static_cast<int>(sizeof(S))) % 16 == 0 ? 16 : ((p.hop * static_cast<int>(sizeof(S))) % 8 == 0 ? 8 : static_cast<int>(sizeof(S)));
    auto staged = [&](int item, int chunk) -> bool {  // same answer for the issuing thread and the consumers
      if (!ring_ok) return false;
      const int clip = item / q.groups;
      const int t0 = chunk * kTcFrames;
      const int64_t first = static_cast<int64_t>(t0) * p.hop - p.pad;
      if (t0 + kTcFrames > p.T || first < 0 || first + span > p.L) return false;
      const int64_t byte0 = (static_cast<int64_t>(clip) * p.L + first) * static_cast<int64_t>(sizeof(S));
      return ((reinterpret_cast<uintptr_t>(p.wav) + static_cast<uintptr_t>(byte0)) & 15u) == 0 && byte0 + span_bytes <= q.total_bytes;
    };
    auto issue = [&](int item, int chunk, uint32_t n) {  // thread 0: fill ring slot n % 3 for the n-th chunk of this CTA
      const uint32_t slot = n % kStages;
      if (staged(item, chunk)) {
        const int clip = item / q.groups;
        const int64_t first = static_cast<int64_t>(chunk) * kTcFrames * p.hop - p.pad;
        const S* src = static_cast<const S*>(p.wav) + static_cast<int64_t>(clip) * p.L + first;
        fence_async_smem();
        mbar_expect_tx(BAR(kRawFull + slot), span_bytes);
        bulk_g2s(sb + kRaw + slot * kStageBytes, src, span_bytes, BAR(kRawFull + slot));
      } else {
        mbar_arrive(BAR(kRawFull + slot));  // nothing to copy: complete the phase
      }
    };
    // look-ahead cursor of the issuing thread
    int la_item = blockIdx.x, la_chunk = 0, la_end = 0;
    uint32_t la_n = 0;
    auto la_open = [&]() {
      if (la_item < n_items) {
        const int grp = la_item % q.groups;
        la_chunk = grp * kTcChunksPerItem;
        la_end = min((grp + 1) * kTcChunksPerItem, q.chunks8);
      }
    };
    auto la_step = [&]() {  // issue the chunk under the cursor and advance
      if (la_item >= n_items) return;
      issue(la_item, la_chunk, la_n);
      ++la_n;
      if (++la_chunk >= la_end) {
        la_item += gridDim.x;
        la_open();
      }
    };
    if (tid == 0) {
      la_open();
      la_step();
      la_step();
    }
    TCP_DECL;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int clip = item / q.groups, grp = item - clip * q.groups;
      const S* __restrict__ x = static_cast<const S*>(p.wav) + static_cast<int64_t>(clip) * p.L;
      AugState aug;
      aug.pcm_scale = p.pcm_scale;
      if (AUG) init_clip_aug(aug, p, clip);
      const int chunk_end = min((grp + 1) * kTcChunksPerItem, q.chunks8);
      for (int chunk = grp * kTcChunksPerItem; chunk < chunk_end; ++chunk, ++it) {
        const uint32_t par = it & 1u;  // tile stage tau is used once per chunk
        const int t0 = chunk * kTcFrames;
        const bool from_ring = staged(item, chunk);
        const uint32_t slot = it % kStages;
        TCP_WORKED();
        mbar_wait_warp_sleep(BAR(kRawFull + slot), (it / kStages) & 1u, lane);
        TCP_WAITED();
        const uint8_t* raw = sm + kRaw + slot * kStageBytes + static_cast<uint32_t>(8 * c8) * sizeof(S);
#pragma unroll 1
        for (int tau = 0; tau < 2; ++tau) {
          float v[4][8];
          float mx = 0.f;
          if (from_ring) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              lds8<S>(raw + static_cast<uint32_t>((4 * tau + j) * p.hop) * sizeof(S), lds_align, p.pcm_scale, v[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int fi = min(t0 + 4 * tau + j, p.T - 1);  // frames past the end repeat the last one (never stored)
              fetch8<AUG, S>(x, static_cast<int64_t>(fi) * p.hop - p.pad + 8 * c8, p.L, aug, v[j]);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              v[j][i] *= win[i];
              mx = fmaxf(mx, fabsf(v[j][i]));
            }
          }
          mx = warp_max(mx);
          if (lane == 0) s_red[tau][warp] = mx;
          TCP_SEG(0);
          role_bar(1, 32 * kWarpsLD);
          TCP_SEG(1);
          // every LD thread is past the previous chunk here: its ring slot may be refilled (two chunks ahead)
          if (tau == 0 && tid == 0) la_step();
#pragma unroll
          for (int i = 0; i < kWarpsLD; ++i) mx = fmaxf(mx, s_red[tau][i]);
          // mx = m 2^e with m in [1, 2): scale 2^(13 - e) puts the largest sample in [2^13, 2^14)
          int es = 0;
          if (mx > 0.f) es = 13 - (static_cast<int>((__float_as_uint(mx) >> 23) & 255u) - 127);
          es = max(-50, min(60, es));
          const float scale = __uint_as_float(static_cast<uint32_t>(es + 127) << 23);
          if (tid == 0) s_es[it & (kRing - 1)][tau] = es;
          TCP_WORKED();
          mbar_wait_warp_sleep(BAR(kA1Empty + tau), par ^ 1u, lane);  // the step-1 MMAs that read this tile buffer completed
          TCP_WAITED();
          uint8_t* a1 = sm + kA1 + tau * kA1Stage;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 hi, lo;
            split2(v[j][0] * scale, v[j][1] * scale, hi.x, lo.x);
            split2(v[j][2] * scale, v[j][3] * scale, hi.y, lo.y);
            split2(v[j][4] * scale, v[j][5] * scale, hi.z, lo.z);
            split2(v[j][6] * scale, v[j][7] * scale, hi.w, lo.w);
            const uint32_t off = static_cast<uint32_t>(j) * 2048u + sw_off;
            *reinterpret_cast<uint4*>(a1 + off) = hi;
            *reinterpret_cast<uint4*>(a1 + kA1Lo + off) = lo;
          }
          TCP_SEG(2);
          fence_async_smem();
          mbar_arrive(BAR(kA1Full + tau));
          TCP_SEG(3);
        }
        if (tid == 0) mbar_arrive(BAR(kEsFull + (it & (kRing - 1))));  // releases both s_es entries to the MEL warps
      }
    }
    TCP_WORKED();
    TCP_STORE(0, tid == 32, it);
    TCP_STORE_SEG(64, tid == 32);
  } else if (warp == kWarpMMA) {
    // =========================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t kIdesc1 = idesc_f16(128, 32, 1, 0);
      constexpr uint32_t kIdesc3 = idesc_f16(128, 64, 1, 0);
      constexpr uint32_t kIdesc3s = idesc_f16(128, 32, 0, 0);
      // descriptors differ from these bases only in the start-address field: + (byte offset >> 4) on the low word
      const uint64_t dA1 = make_desc(sb + kA1, 2048, 512, 4), dB1 = make_desc(sb + kB1, 512, 128, 0);
      const uint64_t dA3 = make_desc(sb + kA3, 2048, 128, 0), dB3 = make_desc(sb + kB3, 1024, 128, 0);
      const uint64_t dA3s = make_desc(sb + kA3s, 128, 128, 0), dB3s = make_desc(sb + kB3s, 512, 128, 0);
      TCP_DECL;
      auto step3 = [&](uint32_t j) {  // step-3 and step-3s GEMMs of chunk j
        const uint32_t st = j & 1u, par = (j >> 1) & 1u;
        TCP_WORKED();
        mbar_wait_sleep(BAR(kA3Full + st), par);
        mbar_wait_sleep(BAR(kAcc3Empty + st), par ^ 1u);
        TCP_WAITED();
        fence_after();
        const uint64_t a3 = dA3 + ((st * kA3Stage) >> 4), a3s = dA3s + ((st * kA3sStage) >> 4);
        const uint32_t d = tmem + kAcc3 + st * kAcc3Stage;
#pragma unroll
        for (int term = 0; term < 3; ++term) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma_f16(d, a3 + (((term == 2 ? kA3Lo : 0u) + ks * 4096u) >> 4), dB3 + (((term == 1 ? 8192u : 0u) + ks * 2048u) >> 4),
                    kIdesc3, term > 0 || ks > 0);
        }
#pragma unroll
        for (int term = 0; term < 3; ++term) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            mma_f16(d + 64u, a3s + (((term == 2 ? kA3sLo : 0u) + ks * 256u) >> 4), dB3s + (((term == 1 ? 2048u : 0u) + ks * 1024u) >> 4),
                    kIdesc3s, term > 0 || ks > 0);
        }
        commit(BAR(kA3Empty + st));
        commit(BAR(kAcc3Full + st));
      };
      uint32_t it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int grp = item % q.groups;
        const int chunk_end = min((grp + 1) * kTcChunksPerItem, q.chunks8);
        for (int chunk = grp * kTcChunksPerItem; chunk < chunk_end; ++chunk, ++it) {
          const uint32_t st = it & 1u, par = (it >> 1) & 1u;
          TCP_WORKED();
          mbar_wait_sleep(BAR(kAcc1Empty + st), par ^ 1u);
          TCP_WAITED();
#pragma unroll
          for (int tau = 0; tau < 2; ++tau) {
            TCP_WORKED();
            mbar_wait_sleep(BAR(kA1Full + tau), it & 1u);
            TCP_WAITED();
            fence_after();
            const uint64_t a1 = dA1 + ((tau * kA1Stage) >> 4);
#pragma unroll
            for (int term = 0; term < 3; ++term) {
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)
                mma_f16(tmem + kAcc1 + st * kAcc1Stage + 32u * tau, a1 + (((term == 2 ? kA1Lo : 0u) + ks * 1024u) >> 4),
                        dB1 + (((term == 1 ? 2048u : 0u) + ks * 1024u) >> 4), kIdesc1, term > 0 || ks > 0);
            }
            commit(BAR(kA1Empty + tau));
          }
          commit(BAR(kAcc1Full + st));
          if (it > 0) step3(it - 1);
        }
      }
      if (it > 0) step3(it - 1);
      TCP_WORKED();
      TCP_STORE(8, true, it);
    }
  } else if (warp < kWarpE3) {
    // =========================================================== E1: acc1 -> twiddle -> A3 / A3s operands
    const int w = warp - kWarpE1;        // frame of the chunk; TMEM lane quadrant = warp & 3 = w & 3, tile = w >> 2
    const uint32_t t_lane = (static_cast<uint32_t>((warp & 3) * 32) << 16) + 32u * static_cast<uint32_t>(w >> 2);
    const float2* tw = reinterpret_cast<const float2*>(sm + kTw);
    const uint32_t row_off = static_cast<uint32_t>(lane >> 3) * 2048u + static_cast<uint32_t>(2 * w) * 128u +
                             static_cast<uint32_t>(lane & 7) * 16u;
    const uint32_t s_off = static_cast<uint32_t>(lane >> 3) * 128u + static_cast<uint32_t>(w) * 16u +
                           static_cast<uint32_t>(lane & 7) * 2u;
    TCP_DECL;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int grp = item % q.groups;
      const int n_chunks = min((grp + 1) * kTcChunksPerItem, q.chunks8) - grp * kTcChunksPerItem;
      for (int c = 0; c < n_chunks; ++c, ++it) {
        const uint32_t st = it & 1u, par = (it >> 1) & 1u;
        TCP_WORKED();
        mbar_wait_warp_sleep(BAR(kAcc1Full + st), par, lane);
        TCP_WAITED();
        fence_after();
        uint32_t y[32];
        tmem_ld32(tmem + t_lane + kAcc1 + st * kAcc1Stage, y);
        fence_before();
        mbar_arrive(BAR(kAcc1Empty + st));  // the accumulator is in registers
        TCP_WORKED();
        mbar_wait_warp_sleep(BAR(kA3Empty + st), par ^ 1u, lane);  // the step-3 MMAs that read this buffer have completed
        TCP_WAITED();
        uint8_t* a3 = sm + kA3 + st * kA3Stage;
#pragma unroll
        for (int g = 0; g < 2; ++g) {  // k1 = 8 g .. 8 g + 7
          float ure[8], uim[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int k1 = 8 * g + i;
            if (k1 == 0) {
              ure[i] = __uint_as_float(y[0]) * 0.03125f;
              uim[i] = 0.f;
            } else {
              const float2 u = c_mul(make_float2(__uint_as_float(y[2 * k1]), __uint_as_float(y[2 * k1 + 1])), tw[k1 * 32 + lane]);
              ure[i] = u.x;
              uim[i] = u.y;
            }
          }
#pragma unroll
          for (int part = 0; part < 2; ++part) {
            const float* u = part == 0 ? ure : uim;
            uint4 hi, lo;
            split2(u[0], u[1], hi.x, lo.x);
            split2(u[2], u[3], hi.y, lo.y);
            split2(u[4], u[5], hi.z, lo.z);
            split2(u[6], u[7], hi.w, lo.w);
            const uint32_t off = row_off + static_cast<uint32_t>(part) * 8192u + static_cast<uint32_t>(g) * 128u;
            *reinterpret_cast<uint4*>(a3 + off) = hi;
            *reinterpret_cast<uint4*>(a3 + kA3Lo + off) = lo;
          }
        }
        {  // Y16 / 32: row = frame, K = n1 of the step-3s operand (K-major)
          const float vv = __uint_as_float(y[1]) * 0.03125f;
          const __half h = __float2half_rn(vv);
          const __half l = __float2half_rn(vv - __half2float(h));
          uint8_t* a3s = sm + kA3s + st * kA3sStage;
          *reinterpret_cast<__half*>(a3s + s_off) = h;
          *reinterpret_cast<__half*>(a3s + kA3sLo + s_off) = l;
        }
        fence_async_smem();
        mbar_arrive(BAR(kA3Full + st));
      }
    }
    TCP_WORKED();
    TCP_STORE(16, w == 0 && lane == 0, it);
  } else if (warp < kWarpMEL) {
    // =========================================================== E3: acc3 -> |X|^2 -> power buffer
    const int quad = warp & 3;
    const uint32_t t_lane = static_cast<uint32_t>(quad * 32) << 16;
    const int f = 2 * quad + (lane >> 4), k1 = lane & 15;
    TCP_DECL;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int grp = item % q.groups;
      const int n_chunks = min((grp + 1) * kTcChunksPerItem, q.chunks8) - grp * kTcChunksPerItem;
      for (int c = 0; c < n_chunks; ++c, ++it) {
        const uint32_t st = it & 1u, par = (it >> 1) & 1u;
        TCP_WORKED();
        mbar_wait_warp_sleep(BAR(kAcc3Full + st), par, lane);
        fence_after();
        mbar_wait_warp_sleep(BAR(kPEmpty + st), par ^ 1u, lane);  // the MEL warps are done with this power buffer
        TCP_WAITED();
        float* s_pow = reinterpret_cast<float*>(sm + kPow + st * kPowStage);
        float* prow = s_pow + f * kPS;
        const uint32_t acc = tmem + t_lane + kAcc3 + st * kAcc3Stage;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t xr[32];
          tmem_ld32(acc + 32u * half, xr);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float re0 = __uint_as_float(xr[4 * j]), re1 = __uint_as_float(xr[4 * j + 1]);
            const float im0 = __uint_as_float(xr[4 * j + 2]), im1 = __uint_as_float(xr[4 * j + 3]);
            const float p0 = re0 * re0 + im0 * im0, p1 = re1 * re1 + im1 * im1;
            if (half == 0) {  // k2 = 2j, 2j + 1: bins k1 + 32 k2
              prow[k1 + 64 * j] = p0;
              prow[k1 + 64 * j + 32] = p1;
            } else {  // k2 = 16 + 2j, 17 + 2j: bins 1024 - k1 - 32 k2 (k1 = 0: only bin 512 is new)
              if (k1 > 0) {
                prow[512 - k1 - 64 * j] = p0;
                prow[480 - k1 - 64 * j] = p1;
              } else if (j == 0) {
                prow[512] = p0;
              }
            }
          }
        }
        if (quad == 0) {  // step 3s: rows = frames (lanes 0..7 of quadrant 0), bins 16 + 32 k2
          uint32_t xs[32];
          tmem_ld32(acc + 64u, xs);
          if (lane < kTcFrames) {
            float* pr = s_pow + lane * kPS;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float re0 = __uint_as_float(xs[4 * j]), re1 = __uint_as_float(xs[4 * j + 1]);
              const float im0 = __uint_as_float(xs[4 * j + 2]), im1 = __uint_as_float(xs[4 * j + 3]);
              pr[16 + 64 * j] = re0 * re0 + im0 * im0;
              pr[48 + 64 * j] = re1 * re1 + im1 * im1;
            }
          }
        }
        fence_before();
        mbar_arrive(BAR(kAcc3Empty + st));
        mbar_arrive(BAR(kPFull + st));
      }
    }
    TCP_WORKED();
    TCP_STORE(24, quad == 0 && lane == 0, it);
  } else {
    // =========================================================== MEL: power -> mel -> log -> tile -> global
    const int mt = tid - 32 * kWarpMEL;  // 0..255: filter mt >> 1, half mt & 1 of its band
    const int mel = mt >> 1, mh = mt & 1;
    const int2 mb = reinterpret_cast<const int2*>(sm + kMBand)[mt];
    const int m_lo = mb.x & 0xffff, m_cnt = mb.x >> 16;
    const float* m_w = reinterpret_cast<const float*>(sm + kMW) + mb.y;
    float mel_scale = 0.f, mel_shift = 0.f;
    if (mel < p.n_mels) {
      const float sd = p.stdv[mel];
      mel_scale = p.log_mult * 0.30102999566398120f / sd;
      mel_shift = -p.mean[mel] / sd;
    }
    float* s_tile = reinterpret_cast<float*>(sm + kTile);
    TCP_DECL;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int clip = item / q.groups, grp = item - clip * q.groups;
      const int chunk_end = min((grp + 1) * kTcChunksPerItem, q.chunks8);
      for (int chunk = grp * kTcChunksPerItem; chunk < chunk_end; ++chunk, ++it) {
        const uint32_t st = it & 1u, par = (it >> 1) & 1u;
        const int t0 = chunk * kTcFrames;
        const int nfr = min(kTcFrames, p.T - t0);
        TCP_WORKED();
        mbar_wait_warp_sleep(BAR(kEsFull + (it & (kRing - 1))), (it / kRing) & 1u, lane);
        mbar_wait_warp_sleep(BAR(kPFull + st), par, lane);
        TCP_WAITED();
        const float* pw = reinterpret_cast<const float*>(sm + kPow + st * kPowStage) + m_lo;
        float2 a01 = make_float2(0.f, 0.f), a23 = a01, a45 = a01, a67 = a01;
#pragma unroll 2
        for (int i = 0; i < m_cnt; ++i) {
          const float w = m_w[32 * i];
          const float2 ww = make_float2(w, w);
          a01 = p_fma(ww, make_float2(pw[i], pw[kPS + i]), a01);
          a23 = p_fma(ww, make_float2(pw[2 * kPS + i], pw[3 * kPS + i]), a23);
          a45 = p_fma(ww, make_float2(pw[4 * kPS + i], pw[5 * kPS + i]), a45);
          a67 = p_fma(ww, make_float2(pw[6 * kPS + i], pw[7 * kPS + i]), a67);
        }
        TCP_SEG(0);
        mbar_arrive(BAR(kPEmpty + st));  // power buffer consumed
        // the two halves of the band: lane pair (2j, 2j + 1); the even lane finishes frames 0-3, the odd one 4-7
        float o0 = mh ? a45.x : a01.x, o1 = mh ? a45.y : a01.y, o2 = mh ? a67.x : a23.x, o3 = mh ? a67.y : a23.y;
        const float g0 = mh ? a01.x : a45.x, g1 = mh ? a01.y : a45.y, g2 = mh ? a23.x : a67.x, g3 = mh ? a23.y : a67.y;
        const float r0 = __shfl_xor_sync(0xffffffffu, g0, 1), r1 = __shfl_xor_sync(0xffffffffu, g1, 1);
        const float r2 = __shfl_xor_sync(0xffffffffu, g2, 1), r3 = __shfl_xor_sync(0xffffffffu, g3, 1);
        // fixed order: half 0 + half 1
        o0 = mh ? r0 + o0 : o0 + r0; o1 = mh ? r1 + o1 : o1 + r1; o2 = mh ? r2 + o2 : o2 + r2; o3 = mh ? r3 + o3 : o3 + r3;
        const int es = s_es[it & (kRing - 1)][mh];  // frames 4 mh .. 4 mh + 3 are tile mh
        const float rescale = __uint_as_float(static_cast<uint32_t>(10 - 2 * es + 127) << 23);  // 32^2 / scale^2
        TCP_SEG(1);
        role_bar(2, 32 * kWarpsMEL);  // the previous chunk's tile has been stored
        TCP_SEG(2);
        if (mel < p.n_mels) {
          float* tr = s_tile + mel * kTileS + 4 * mh;
          tr[0] = norm_db(o0 * rescale, p.log_eps, mel_scale, mel_shift);
          tr[1] = norm_db(o1 * rescale, p.log_eps, mel_scale, mel_shift);
          tr[2] = norm_db(o2 * rescale, p.log_eps, mel_scale, mel_shift);
          tr[3] = norm_db(o3 * rescale, p.log_eps, mel_scale, mel_shift);
        }
        role_bar(2, 32 * kWarpsMEL);  // tile complete
        TCP_SEG(2);
        float* o = p.out + static_cast<int64_t>(clip) * p.n_mels * p.T + t0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int e = mt + 256 * r;
          const int m = e >> 3, fl = e & 7;
          if (m < p.n_mels && fl < nfr) o[static_cast<int64_t>(m) * p.T + fl] = s_tile[m * kTileS + fl];
        }
      }
    }
    TCP_WORKED();
    TCP_SEG(3);
    TCP_STORE(32 + 4 * (mt >> 5), (mt & 31) == 0, it);
    TCP_STORE_SEG(68, mt == 0);
    TCP_STORE_SEG(72, mt == 224);
  }
#undef BAR

  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

void split_half(double v, __half& hi, __half& lo) {
  hi = __float2half_rn(static_cast<float>(v));
  lo = __float2half_rn(static_cast<float>(v - static_cast<double>(__half2float(hi))));
}

}  // namespace

int tc_tables_create(afs_logmel_plan* plan, const float* fb_host) {
  plan->d_tc = nullptr;
  std::vector<uint8_t> img(kConstBytes, 0);
  const double two_pi = 6.283185307179586476925286766559;
  __half* b1h = reinterpret_cast<__half*>(img.data());
  __half* b1l = b1h + 1024;
  __half* b3h = reinterpret_cast<__half*>(img.data() + 4096);
  __half* b3l = b3h + 4096;
  __half* bsh = reinterpret_cast<__half*>(img.data() + 4096 + 16384);
  __half* bsl = bsh + 1024;
  float2* tw = reinterpret_cast<float2*>(img.data() + 4096 + 16384 + 4096);
  // B1 [K = n2][N = c]: c = 0 -> Y0, c = 1 -> Y16, c = 2 k1 -> Re Y[k1], c = 2 k1 + 1 -> Im Y[k1]
  for (int n2 = 0; n2 < 32; ++n2)
    for (int c = 0; c < 32; ++c) {
      double v;
      if (c == 0) v = 1.0;
      else if (c == 1) v = (n2 & 1) ? -1.0 : 1.0;
      else {
        const double a = two_pi * ((n2 * (c >> 1)) % 32) / 32.0;
        v = (c & 1) ? -sin(a) : cos(a);
      }
      const int idx = (n2 >> 3) * 256 + c * 8 + (n2 & 7);
      split_half(v, b1h[idx], b1l[idx]);
    }
  // B3 [K = part * 32 + n1][N = col], col = 4 (k2 >> 1) + 2 * (re: 0 | im: 1) + (k2 & 1)
  //   Re X = sum Ure cos + Uim sin,  Im X = sum -Ure sin + Uim cos   (theta = 2 pi n1 k2 / 32)
  for (int kk = 0; kk < 64; ++kk)
    for (int k2 = 0; k2 < 32; ++k2)
      for (int im = 0; im < 2; ++im) {
        const int n1 = kk & 31, part = kk >> 5;
        const double th = two_pi * ((n1 * k2) % 32) / 32.0;
        const double v = im == 0 ? (part == 0 ? cos(th) : sin(th)) : (part == 0 ? -sin(th) : cos(th));
        const int col = 4 * (k2 >> 1) + 2 * im + (k2 & 1);
        const int idx = (kk >> 3) * 512 + col * 8 + (kk & 7);
        split_half(v, b3h[idx], b3l[idx]);
      }
  // B3s [K = n1][N = col]: X[16 + 32 k2] = sum_n1 V[n1] exp(-2 pi i n1 (2 k2 + 1) / 64), k2 = 0..15
  for (int n1 = 0; n1 < 32; ++n1)
    for (int k2 = 0; k2 < 16; ++k2)
      for (int im = 0; im < 2; ++im) {
        const double th = two_pi * ((n1 * (2 * k2 + 1)) % 64) / 64.0;
        const double v = im == 0 ? cos(th) : -sin(th);
        const int col = 4 * (k2 >> 1) + 2 * im + (k2 & 1);
        const int idx = (n1 >> 3) * 256 + col * 8 + (n1 & 7);
        split_half(v, bsh[idx], bsl[idx]);
      }
  // twiddles W_1024^(n1 k1) / 32
  for (int k1 = 0; k1 < 16; ++k1)
    for (int n1 = 0; n1 < 32; ++n1) {
      const double a = two_pi * ((n1 * k1) % 1024) / 1024.0;
      tw[k1 * 32 + n1] = make_float2(static_cast<float>(cos(a) / 32.0), static_cast<float>(-sin(a) / 32.0));
    }
  // MEL role table: thread t = (filter t >> 1, half t & 1) owns half of the filter's band; its weight i sits at
  // mw[off + 32 i] ([warp][i][lane], zero padded to the longest half of the warp: conflict-free weight loads)
  std::vector<int> b0;
  std::vector<float> w0;
  pack_mel_bands(fb_host, plan->cfg.n_mels, b0, w0);
  std::vector<int2> mband(256, make_int2(0, 0));
  std::vector<float> mw;
  for (int warp = 0; warp < 8; ++warp) {
    int start[32], cnt[32], src[32], rows = 0;
    for (int lane = 0; lane < 32; ++lane) {
      const int t = warp * 32 + lane, m = t >> 1, h = t & 1;
      start[lane] = cnt[lane] = src[lane] = 0;
      if (m < plan->cfg.n_mels) {
        const int lo = b0[m], len = b0[kMaxMels + m], n0 = (len + 1) / 2;
        start[lane] = h == 0 ? lo : lo + n0;
        cnt[lane] = h == 0 ? n0 : len - n0;
        src[lane] = b0[2 * kMaxMels + m] + (h == 0 ? 0 : n0);
      }
      if (cnt[lane] > rows) rows = cnt[lane];
    }
    const int base = static_cast<int>(mw.size());
    mw.resize(mw.size() + static_cast<size_t>(rows) * 32, 0.f);
    for (int lane = 0; lane < 32; ++lane) {
      for (int i = 0; i < cnt[lane]; ++i) mw[base + i * 32 + lane] = w0[src[lane] + i];
      mband[warp * 32 + lane] = make_int2(start[lane] | (cnt[lane] << 16), base + lane);
    }
  }
  if (mw.empty()) mw.push_back(0.f);
  if (mw.size() > static_cast<size_t>(kMaxMelW)) return AFS_OK;  // very wide filters: the plan runs the FFT engine only

  TcTables* t = new (std::nothrow) TcTables();
  if (t == nullptr) return AFS_ERR_CUDA;
  t->d_consts = nullptr; t->d_mband = nullptr; t->d_mw = nullptr;
  t->mw_count = static_cast<int>(mw.size());
  cudaError_t e = cudaMalloc(&t->d_consts, kConstBytes);
  if (e == cudaSuccess) e = cudaMalloc(&t->d_mband, 256 * sizeof(int2));
  if (e == cudaSuccess) e = cudaMalloc(&t->d_mw, mw.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(t->d_consts, img.data(), kConstBytes, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(t->d_mband, mband.data(), 256 * sizeof(int2), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(t->d_mw, mw.data(), mw.size() * sizeof(float), cudaMemcpyHostToDevice);
  plan->d_tc = t;
  if (e != cudaSuccess) {
    tc_tables_destroy(plan);
    return AFS_ERR_CUDA;
  }
  return AFS_OK;
}

void tc_tables_destroy(afs_logmel_plan* plan) {
  TcTables* t = static_cast<TcTables*>(plan->d_tc);
  if (t != nullptr) {
    cudaFree(t->d_consts);
    cudaFree(t->d_mband);
    cudaFree(t->d_mw);
    delete t;
  }
  plan->d_tc = nullptr;
}

template <typename S>
int tc_launch(const afs_logmel_plan* plan, const Params& p, bool aug, cudaStream_t stream) {
  const TcTables* t = static_cast<const TcTables*>(plan->d_tc);
  if (t == nullptr) return AFS_ERR_UNSUPPORTED;
  TcArgs q;
  q.consts = static_cast<const uint8_t*>(t->d_consts);
  q.mband = t->d_mband;
  q.mw = t->d_mw;
  q.mw_count = t->mw_count;
  q.chunks8 = (p.T + kTcFrames - 1) / kTcFrames;
  q.groups = (q.chunks8 + kTcChunksPerItem - 1) / kTcChunksPerItem;
  q.total_bytes = static_cast<int64_t>(p.B) * p.L * static_cast<int64_t>(sizeof(S));
  const int64_t items = static_cast<int64_t>(p.B) * q.groups;
  if (items > 0x7fffffffLL) return AFS_ERR_UNSUPPORTED;
  const size_t smem = kMW + static_cast<size_t>(t->mw_count) * 4 + 1024;
  const unsigned grid = static_cast<unsigned>(items < kNumSMs ? items : kNumSMs);  // persistent: one CTA per SM
  if (aug) {
    AFS_CUDA_TRY(cudaFuncSetAttribute(logmel_tc_kernel<true, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    logmel_tc_kernel<true, S><<<grid, kTcThreads, smem, stream>>>(p, q);
  } else {
    AFS_CUDA_TRY(cudaFuncSetAttribute(logmel_tc_kernel<false, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    logmel_tc_kernel<false, S><<<grid, kTcThreads, smem, stream>>>(p, q);
  }
  AFS_LAUNCH_CHECK();
  return AFS_OK;
}

template int tc_launch<float>(const afs_logmel_plan*, const Params&, bool, cudaStream_t);
template int tc_launch<int16_t>(const afs_logmel_plan*, const Params&, bool, cudaStream_t);

}  // namespace logmel
}  // namespace afs

#ifdef AFS_TC_PROFILE
extern "C" int afs_logmel_tc_profile_read(long long* host64) {
  return cudaMemcpyFromSymbol(host64, afs::logmel::g_tc_prof, 96 * sizeof(long long)) == cudaSuccess ? AFS_OK : AFS_ERR_CUDA;
}
#endif
