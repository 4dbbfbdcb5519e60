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
// CTA = 256 threads, 2 CTAs per SM (one runs SIMT phases while the other's MMAs are in flight); a chunk is 8 frames
// of one clip:
//   P0/P1  thread (frame f, 8-sample group c): load (reflect pad / Philox augmentation / PCM16 as the FFT engine),
//          window, chunk max -> power-of-two scale, split to fp16 hi/lo, 16-byte stores into the A1 operand
//          [frame][32 sample rows of 32][SWIZZLE_64B]: frame memory order IS the MN-major operand (M = (f, n1)).
//   P2     one thread: 2 tiles x 3 terms x 2 K-steps of tcgen05.mma M128 N32 K16, commit -> mbarrier
//   P3     warp = frame, lane = n1: tcgen05.ld 32 columns, twiddle, fp16 split, 16-byte stores into the A3 operand
//          (MN-major: 8 consecutive k1 of one n1 are one store), Y16 into the A3s operand
//   P4     one thread: 3 x 4 MMAs M128 N64 K16 (step 3) + 3 x 2 MMAs N32 (step 3s), commit
//   P5     warp (lane quadrant, column half): tcgen05.ld, |X|^2 into the power buffer [frame][528]
//   P6/P7  banded mel projection (the FFT engine's ELL table and start shifts), log, normalise, [128 x 8] tile,
//          store along time.
#include <cuda_fp16.h>
#include <math.h>

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

constexpr int kTcThreads = 256;
constexpr int kTcFrames = 8;         // frames per chunk
constexpr int kTcChunksPerItem = 5;  // consecutive chunks of one clip per work item (partial output sectors meet in L2)
constexpr int kPS = 528;             // power row stride in floats: == 16 mod 32, the |X|^2 stores are conflict-free
constexpr int kTileS = kTcFrames + 1;

// shared-memory map, byte offsets from a 1024-aligned base
constexpr uint32_t kA1Hi = 0, kA1Lo = 16384;                    // [8 frames][2048 B]
constexpr uint32_t kPow = 0;                                    // aliases A1 (dead once the step-1 MMAs completed)
constexpr uint32_t kTile = kPS * kTcFrames * 4;                 // out tile [128][9] floats, ends at 21 504
constexpr uint32_t kA3Hi = 32768, kA3Lo = 49152;                // [8 K groups][16 row groups][128 B]
constexpr uint32_t kA3sHi = 65536, kA3sLo = 66048;              // [4 K groups][8 rows][16 B]
constexpr uint32_t kB1 = 66560;                                 // hi 2048 | lo 2048   [4 K groups][32 rows][16 B]
constexpr uint32_t kB3 = kB1 + 4096;                            // hi 8192 | lo 8192   [8 K groups][64 rows][16 B]
constexpr uint32_t kB3s = kB3 + 16384;                          // hi 2048 | lo 2048   [4 K groups][32 rows][16 B]
constexpr uint32_t kTw = kB3s + 4096;                           // float2 [16][32]
constexpr uint32_t kConstBytes = kTw + 4096 - kB1;              // 28 672: one contiguous image built by the host
constexpr uint32_t kBand = kTw + 4096;                          // int [3][128]
constexpr uint32_t kWts = kBand + 3 * kMaxMels * 4;             // float [nnz]
static_assert(kTile + kMaxMels * kTileS * 4 <= kA3Hi, "power buffer and out tile must fit in the A1 region");

constexpr uint32_t kAcc1 = 0, kAcc3 = 64, kAcc3s = 128, kTmemCols = 256;

struct TcArgs {
  const uint8_t* consts;  // kConstBytes
  float* dbg;             // nullable: item 0 / chunk 0 dumps acc1 [256][32], then the power buffer [8][528]
  int chunks8;            // ceil(T / 8)
  int groups;             // ceil(chunks8 / kTcChunksPerItem)
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

template <bool AUG, typename S>
__global__ void __launch_bounds__(kTcThreads, 2) logmel_tc_kernel(const Params p, const TcArgs q) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  __shared__ float s_red[8];
  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sb = smem_u32(sm);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  if (warp == 0) tmem_alloc(&s_tmem, kTmemCols);
  if (tid == 32) {
    mbar_init(smem_u32(&s_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < static_cast<int>(kConstBytes / 16); i += kTcThreads)
    reinterpret_cast<uint4*>(sm + kB1)[i] = __ldg(reinterpret_cast<const uint4*>(q.consts) + i);
  {
    int* s_band = reinterpret_cast<int*>(sm + kBand);
    float* s_w = reinterpret_cast<float*>(sm + kWts);
    for (int i = tid; i < 3 * kMaxMels; i += kTcThreads) s_band[i] = p.band[i];
    for (int i = tid; i < p.nnz; i += kTcThreads) s_w[i] = p.weights[i];
  }
  // loader role: 8-sample group c8 of frames fr0, fr0 + 2, fr0 + 4, fr0 + 6
  const int c8 = tid & 127, fr0 = tid >> 7;
  float win[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) win[i] = __ldg(p.window + 8 * c8 + i);
  uint32_t sw_off = 16u * c8;
  sw_off ^= ((sw_off >> 7) & 3u) << 4;  // SWIZZLE_64B: 16-byte chunk index ^= bits 7..8 of the (1024-aligned) address
  // mel role: 64-thread groups; group g = (pass, batch of 4 frames)
  const int t2 = tid & 63, mg = tid >> 6;
  const int mel_batch = mg & 1;
  const int mel = (mg >> 1) == 0 ? (t2 < p.n_mels ? t2 : -1) : (p.n_mels - 1 - t2 >= kGroup ? p.n_mels - 1 - t2 : -1);
  float mel_scale = 0.f, mel_shift = 0.f;
  if (mel >= 0) {
    const float sd = p.stdv[mel];
    mel_scale = p.log_mult * 0.30102999566398120f / sd;
    mel_shift = -p.mean[mel] / sd;
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t bar = smem_u32(&s_bar);
  uint32_t phase = 0;

  constexpr uint32_t kIdesc1 = idesc_f16(128, 32, 1, 0);
  constexpr uint32_t kIdesc3 = idesc_f16(128, 64, 1, 0);
  constexpr uint32_t kIdesc3s = idesc_f16(128, 32, 0, 0);

  const int n_items = p.B * q.groups;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int clip = item / q.groups;
    const int grp = item - clip * q.groups;
    const S* __restrict__ x = static_cast<const S*>(p.wav) + static_cast<int64_t>(clip) * p.L;
    AugState aug;
    aug.pcm_scale = p.pcm_scale;
    if (AUG) init_clip_aug(aug, p, clip);
    const int chunk_end = min((grp + 1) * kTcChunksPerItem, q.chunks8);

    for (int chunk = grp * kTcChunksPerItem; chunk < chunk_end; ++chunk) {
      const int t0 = chunk * kTcFrames;
      const int nfr = min(kTcFrames, p.T - t0);

      // ---- P0: load, window, chunk maximum
      float v[4][8];
      float mx = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int fi = min(t0 + fr0 + 2 * j, p.T - 1);  // frames past the end repeat the last one (never stored)
        fetch8<AUG, S>(x, static_cast<int64_t>(fi) * p.hop - p.pad + 8 * c8, p.L, aug, v[j]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          v[j][i] *= win[i];
          mx = fmaxf(mx, fabsf(v[j][i]));
        }
      }
      mx = warp_max(mx);
      if (lane == 0) s_red[warp] = mx;
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 8; ++i) mx = fmaxf(mx, s_red[i]);
      // mx = m 2^e with m in [1, 2): scale 2^(13 - e) puts it in [2^13, 2^14)
      int es = 0;
      if (mx > 0.f) es = 13 - (static_cast<int>((__float_as_uint(mx) >> 23) & 255u) - 127);
      es = max(-50, min(60, es));
      const float scale = __uint_as_float(static_cast<uint32_t>(es + 127) << 23);
      const float rescale = __uint_as_float(static_cast<uint32_t>(10 - 2 * es + 127) << 23);  // |X|^2 * 32^2 / scale^2

      // ---- P1: fp16 hi/lo split into the step-1 operand
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 hi, lo;
        split2(v[j][0] * scale, v[j][1] * scale, hi.x, lo.x);
        split2(v[j][2] * scale, v[j][3] * scale, hi.y, lo.y);
        split2(v[j][4] * scale, v[j][5] * scale, hi.z, lo.z);
        split2(v[j][6] * scale, v[j][7] * scale, hi.w, lo.w);
        const uint32_t off = static_cast<uint32_t>(fr0 + 2 * j) * 2048u + sw_off;
        *reinterpret_cast<uint4*>(sm + kA1Hi + off) = hi;
        *reinterpret_cast<uint4*>(sm + kA1Lo + off) = lo;
      }
      fence_async_smem();
      fence_before();
      __syncthreads();

      // ---- P2: step-1 GEMMs
      if (tid == 0) {
        fence_after();
#pragma unroll
        for (int tau = 0; tau < 2; ++tau) {
#pragma unroll
          for (int term = 0; term < 3; ++term) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint32_t a = sb + (term == 2 ? kA1Lo : kA1Hi) + tau * 8192u + ks * 1024u;
              const uint32_t b = sb + kB1 + (term == 1 ? 2048u : 0u) + ks * 1024u;
              mma_f16(tmem + kAcc1 + 32u * tau, make_desc(a, 2048, 512, 4), make_desc(b, 512, 128, 0), kIdesc1,
                      term > 0 || ks > 0);
            }
          }
        }
        commit(bar);
      }
      mbar_wait(bar, phase);
      phase ^= 1u;
      fence_after();

      // ---- P3: warp = frame, lane = n1: twiddle, split, step-3 operands
      {
        uint32_t y[32];
        tmem_ld32(tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + kAcc1 + 32u * (warp >> 2), y);
        if (q.dbg != nullptr && item == 0 && chunk == 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i) q.dbg[tid * 32 + i] = __uint_as_float(y[i]);
        }
        const float2* tw = reinterpret_cast<const float2*>(sm + kTw);
        float ure[16], uim[16];
        ure[0] = __uint_as_float(y[0]) * 0.03125f;
        uim[0] = 0.f;
#pragma unroll
        for (int k1 = 1; k1 < 16; ++k1) {
          const float2 u = c_mul(make_float2(__uint_as_float(y[2 * k1]), __uint_as_float(y[2 * k1 + 1])), tw[k1 * 32 + lane]);
          ure[k1] = u.x;
          uim[k1] = u.y;
        }
        const uint32_t row_off = static_cast<uint32_t>(lane >> 3) * 2048u + static_cast<uint32_t>(2 * warp) * 128u +
                                 static_cast<uint32_t>(lane & 7) * 16u;
#pragma unroll
        for (int part = 0; part < 2; ++part) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const float* u = part == 0 ? ure : uim;
            uint4 hi, lo;
            split2(u[8 * g + 0], u[8 * g + 1], hi.x, lo.x);
            split2(u[8 * g + 2], u[8 * g + 3], hi.y, lo.y);
            split2(u[8 * g + 4], u[8 * g + 5], hi.z, lo.z);
            split2(u[8 * g + 6], u[8 * g + 7], hi.w, lo.w);
            const uint32_t off = row_off + static_cast<uint32_t>(part) * 8192u + static_cast<uint32_t>(g) * 128u;
            *reinterpret_cast<uint4*>(sm + kA3Hi + off) = hi;
            *reinterpret_cast<uint4*>(sm + kA3Lo + off) = lo;
          }
        }
        {  // Y16 / 32: row = frame, K = n1 of the step-3s operand (K-major)
          const float vv = __uint_as_float(y[1]) * 0.03125f;
          const __half h = __float2half_rn(vv);
          const __half l = __float2half_rn(vv - __half2float(h));
          const uint32_t off = static_cast<uint32_t>(lane >> 3) * 128u + static_cast<uint32_t>(warp) * 16u +
                               static_cast<uint32_t>(lane & 7) * 2u;
          *reinterpret_cast<__half*>(sm + kA3sHi + off) = h;
          *reinterpret_cast<__half*>(sm + kA3sLo + off) = l;
        }
      }
      fence_async_smem();
      fence_before();
      __syncthreads();

      // ---- P4: step-3 and step-3s GEMMs
      if (tid == 0) {
        fence_after();
#pragma unroll
        for (int term = 0; term < 3; ++term) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t a = sb + (term == 2 ? kA3Lo : kA3Hi) + ks * 4096u;
            const uint32_t b = sb + kB3 + (term == 1 ? 8192u : 0u) + ks * 2048u;
            mma_f16(tmem + kAcc3, make_desc(a, 2048, 128, 0), make_desc(b, 1024, 128, 0), kIdesc3, term > 0 || ks > 0);
          }
        }
#pragma unroll
        for (int term = 0; term < 3; ++term) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint32_t a = sb + (term == 2 ? kA3sLo : kA3sHi) + ks * 256u;
            const uint32_t b = sb + kB3s + (term == 1 ? 2048u : 0u) + ks * 1024u;
            mma_f16(tmem + kAcc3s, make_desc(a, 128, 128, 0), make_desc(b, 512, 128, 0), kIdesc3s, term > 0 || ks > 0);
          }
        }
        commit(bar);
      }
      mbar_wait(bar, phase);
      phase ^= 1u;
      fence_after();

      // ---- P5: |X|^2 into the power buffer
      float* s_pow = reinterpret_cast<float*>(sm + kPow);
      {
        const int quad = warp & 3, half = warp >> 2;
        uint32_t xr[32];
        tmem_ld32(tmem + (static_cast<uint32_t>(quad * 32) << 16) + kAcc3 + 32u * half, xr);
        const int f = 2 * quad + (lane >> 4), k1 = lane & 15;
        float* prow = s_pow + f * kPS;
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
        if (warp == 4) {  // step 3s: rows = frames (lanes 0..7 of quadrant 0), bins 16 + 32 k2
          uint32_t xs[32];
          tmem_ld32(tmem + kAcc3s, xs);
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
      }
      fence_before();
      __syncthreads();
      if (q.dbg != nullptr && item == 0 && chunk == 0) {
        for (int i = tid; i < kTcFrames * kPS; i += kTcThreads) q.dbg[256 * 32 + i] = s_pow[i] * rescale;
      }

      // ---- P6: banded mel projection of 4 frames per thread, log, normalise
      float* s_tile = reinterpret_cast<float*>(sm + kTile);
      if (mel >= 0) {
        const int* s_band = reinterpret_cast<const int*>(sm + kBand);
        const float* s_w = reinterpret_cast<const float*>(sm + kWts);
        float acc[4];
        mel_dot_batch_p<kPS>(s_pow + mel_batch * 4 * kPS, s_w + s_band[2 * kMaxMels + mel], kEllStride, s_band[mel],
                             s_band[kMaxMels + mel], acc);
#pragma unroll
        for (int f = 0; f < 4; ++f)
          s_tile[mel * kTileS + mel_batch * 4 + f] = norm_db(acc[f] * rescale, p.log_eps, mel_scale, mel_shift);
      }
      __syncthreads();

      // ---- P7: store along time (8 frames = 32 bytes per mel row)
      {
        float* o = p.out + static_cast<int64_t>(clip) * p.n_mels * p.T + t0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int e = tid + kTcThreads * r;
          const int m = e >> 3, fl = e & 7;
          if (m < p.n_mels && fl < nfr) o[static_cast<int64_t>(m) * p.T + fl] = s_tile[m * kTileS + fl];
        }
      }
      __syncthreads();  // the A1 region (power buffer, tile) is free for the next chunk
    }
  }

  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

void split_half(double v, __half& hi, __half& lo) {
  hi = __float2half_rn(static_cast<float>(v));
  lo = __float2half_rn(static_cast<float>(v - static_cast<double>(__half2float(hi))));
}

}  // namespace

int tc_tables_create(afs_logmel_plan* plan, const float* /*window_host*/) {
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
  void* d = nullptr;
  if (cudaMalloc(&d, kConstBytes) != cudaSuccess) return AFS_ERR_CUDA;
  if (cudaMemcpy(d, img.data(), kConstBytes, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(d);
    return AFS_ERR_CUDA;
  }
  plan->d_tc = d;
  return AFS_OK;
}

void tc_tables_destroy(afs_logmel_plan* plan) {
  if (plan->d_tc != nullptr) cudaFree(plan->d_tc);
  plan->d_tc = nullptr;
}

static float* g_tc_debug = nullptr;  // development hook (afs_logmel_tc_debug_buffer)

template <typename S>
int tc_launch(const afs_logmel_plan* plan, const Params& p, bool aug, cudaStream_t stream) {
  if (plan->d_tc == nullptr) return AFS_ERR_UNSUPPORTED;
  TcArgs q;
  q.consts = static_cast<const uint8_t*>(plan->d_tc);
  q.dbg = g_tc_debug;
  q.chunks8 = (p.T + kTcFrames - 1) / kTcFrames;
  q.groups = (q.chunks8 + kTcChunksPerItem - 1) / kTcChunksPerItem;
  const int64_t items = static_cast<int64_t>(p.B) * q.groups;
  if (items > 0x7fffffffLL) return AFS_ERR_UNSUPPORTED;
  const size_t smem = kWts + static_cast<size_t>(plan->nnz) * 4 + 1024;
  const unsigned grid = static_cast<unsigned>(items < 2 * kNumSMs ? items : 2 * kNumSMs);
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

extern "C" int afs_logmel_tc_debug_buffer(float* device_buffer) {
  afs::logmel::g_tc_debug = device_buffer;
  return AFS_OK;
}
