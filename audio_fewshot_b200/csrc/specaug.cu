// Spectrogram-domain augmentations of the reference, fused: de-normalise -> augment -> re-normalise
// in ONE pass over each [H, W] plane (one CTA per plane, the plane lives in shared memory).
//
// Reference: libfewshot_core/audio_augmentations.py -- augment_spectrogram :531-604 wraps one of
//   random_cutout :56-103, apply_linear_filteraugment :467-528, background_noise_suppression :106-158,
//   adaptive_noise_profile_matching :388-464, temporal_median_background_subtraction :161-209,
//   spectral_contrast_enhancement :212-266, foreground_energy_normalization :269-325,
//   wiener_like_filtering :328-385
// between denormalize_spectrogram :16-33 and normalize_spectrogram :36-53.  There the quantile-based
// ones loop in Python over every (batch, channel) plane and call torch.quantile (a full sort) on it;
// the OOD test-time-augmentation loop (test.py:382-420) runs that 10x per query.
// Here torch.quantile's order statistics are found by an exact 3-pass radix select (11-bit digits, one shared
// histogram filled with warp-aggregated atomics, two-level bin scan) on the IEEE bit patterns in shared memory (no
// sort), the second order statistic of an interpolated quantile by one counting pass, the per-row quantiles of the
// background subtraction bit by bit on register-resident rows (32 ballot rounds per row and warp); torch's fp32 rank
// arithmetic and lerp are reproduced, and the plane is read from and written to HBM exactly once with 128-bit
// accesses (2 * 4 * H * W bytes per plane).  1 024 threads per plane: round 2's first version (256 threads, per-warp
// 256-bin histograms with per-lane atomics, a one-thread bin walk, W comparisons per element for the row quantiles)
// ran 0.19 / 0.40 / 0.93 ms per 800 planes for cutout / noise suppression / background subtraction; this one 0.046 /
// 0.17 / 0.30 ms.
#include "common.cuh"

namespace afs {
namespace {

constexpr int kThreads = 1024;  // one CTA per plane and SM (the plane and its keys fill 160 KB of shared memory)
constexpr int kWarps = kThreads / 32;
constexpr int kBins = 2048;      // 11-bit digits: three passes (11 + 11 + 10 bits) instead of four 8-bit ones

__device__ __forceinline__ float block_sum(float v, float* s_red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < kWarps; ++i) t += s_red[i];
  return t;
}

// Exact k-th smallest (0-based) of keys[0..n) -- non-negative floats as uint32 -- by MSB-first radix select.  All
// threads return the same value.  One shared 2048-bin histogram per pass, filled with warp-aggregated atomics (the
// magnitudes of a plane share their leading bits: per-lane atomics on a handful of bins serialise 32 ways); the bin that
// holds rank k is found by a two-level scan (warp w scans bins 64 w .. 64 w + 63 with shuffles, warp 0 scans the 32
// warp totals) instead of one thread walking the bins.
__device__ uint32_t radix_select(const uint32_t* keys, int n, int k, int* s_hist, int* s_pick) {
  static_assert(kWarps * 64 == kBins, "one warp per 64 bins");
  uint32_t prefix = 0, mask = 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int* s_wsum = s_pick + 4;  // [kWarps] warp totals
  const int n_up = (n + 31) & ~31;
#pragma unroll 1
  for (int pass = 0; pass < 3; ++pass) {
    const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
    const uint32_t dmask = pass == 2 ? 1023u : 2047u;
    __syncthreads();
    for (int i = tid; i < kBins; i += kThreads) s_hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < n_up; i += kThreads) {  // warp-uniform trip count: the ballot below needs every lane
      const uint32_t key = i < n ? keys[i] : 0u;
      const bool in = i < n && (key & mask) == prefix;
      const uint32_t digit = (key >> shift) & dmask;
      const unsigned act = __ballot_sync(0xffffffffu, in);
      if (in) {
        const unsigned peers = __match_any_sync(act, digit);
        if (lane == __ffs(peers) - 1) atomicAdd(&s_hist[digit], __popc(peers));
      }
    }
    __syncthreads();
    // inclusive scan of this warp's 64 bins (two per lane), warp total to s_wsum
    const int2 h = *reinterpret_cast<const int2*>(&s_hist[64 * warp + 2 * lane]);
    int incl = h.x + h.y;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {  // which warp's bins hold rank k, and the rank inside them
      const int tot = s_wsum[lane];
      int cum = tot;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, cum, d);
        if (lane >= d) cum += t;
      }
      if (cum - tot <= k && k < cum) {
        s_pick[2] = lane;
        s_pick[3] = k - (cum - tot);
      }
    }
    __syncthreads();
    if (warp == s_pick[2]) {
      const int kk = s_pick[3];
      const int excl = incl - h.x - h.y;
      if (excl <= kk && kk < excl + h.x) {
        s_pick[0] = 64 * warp + 2 * lane;
        s_pick[1] = kk - excl;
      } else if (excl + h.x <= kk && kk < incl) {
        s_pick[0] = 64 * warp + 2 * lane + 1;
        s_pick[1] = kk - excl - h.x;
      }
    }
    __syncthreads();
    prefix |= static_cast<uint32_t>(s_pick[0]) << shift;
    mask |= dmask << shift;
    k = s_pick[1];
  }
  return prefix;
}

// The next order statistic after rank `lo` (value v_lo), i.e. rank lo + 1, in ONE pass instead of a second select: if
// more than lo + 1 keys are <= v_lo the value repeats, otherwise it is the smallest key above v_lo.
__device__ uint32_t next_order_stat(const uint32_t* keys, int n, uint32_t v_lo, int lo, int* s_pick) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int cnt = 0;
  uint32_t mn = 0xffffffffu;
  for (int i = tid; i < n; i += kThreads) {
    const uint32_t key = keys[i];
    cnt += key <= v_lo ? 1 : 0;
    mn = key > v_lo ? min(mn, key) : mn;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
  }
  int* s_c = s_pick + 4;                                    // [kWarps] counts
  uint32_t* s_m = reinterpret_cast<uint32_t*>(s_pick + 4 + kWarps);  // [kWarps] minima
  __syncthreads();
  if (lane == 0) {
    s_c[warp] = cnt;
    s_m[warp] = mn;
  }
  __syncthreads();
  int c = 0;
  uint32_t m = 0xffffffffu;
  for (int w = 0; w < kWarps; ++w) {
    c += s_c[w];
    m = min(m, s_m[w]);
  }
  return c > lo + 1 ? v_lo : m;
}

// torch.quantile(v, q) with linear interpolation, for a fp32 input: rank = q * (n - 1) in fp32,
// lerp(v[floor], v[ceil], frac) with ATen's two-sided lerp formula.
__device__ float quantile_abs(const uint32_t* keys, int n, float q, int* s_hist, int* s_pick) {
  const float rank = __fmul_rn(q, static_cast<float>(n - 1));
  const int lo = static_cast<int>(floorf(rank));
  const int hi = static_cast<int>(ceilf(rank));
  const float w = rank - static_cast<float>(lo);
  const uint32_t ka = radix_select(keys, n, lo, s_hist, s_pick);
  const float a = __uint_as_float(ka);
  if (hi == lo) return a;
  const float b = __uint_as_float(next_order_stat(keys, n, ka, lo, s_pick));  // hi == lo + 1
  const float d = b - a;
  return w < 0.5f ? __fmaf_rn(w, d, a) : b - d * (1.0f - w);
}

// Order statistics lo and hi (hi == lo or lo + 1) of a row of W <= 32 * PER floats by a warp: the row is held in
// registers as order-preserving integer keys and rank lo is found bit by bit (32 rounds of PER ballots) -- exact.
template <int PER>
__device__ __forceinline__ void row_order_stats(const float* row, int W, int lo, int hi, int lane, float& a, float& b) {
  uint32_t kk[PER];
  bool alive[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = lane + 32 * j;
    const uint32_t u = i < W ? __float_as_uint(row[i]) : 0u;
    kk[j] = i < W ? ((u & 0x80000000u) ? ~u : (u | 0x80000000u)) : 0xffffffffu;  // ascending as unsigned
    alive[j] = i < W;
  }
  int k = lo;
  uint32_t key = 0;
#pragma unroll 1
  for (int bit = 31; bit >= 0; --bit) {
    int c = 0;
    bool zero[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      zero[j] = alive[j] && ((kk[j] >> bit) & 1u) == 0u;
      c += __popc(__ballot_sync(0xffffffffu, zero[j]));
    }
    const bool take_zero = k < c;  // the k-th smallest of the candidates has this bit clear
    if (!take_zero) {
      k -= c;
      key |= 1u << bit;
    }
#pragma unroll
    for (int j = 0; j < PER; ++j) alive[j] = take_zero ? zero[j] : (alive[j] && !zero[j]);
  }
  // rank lo + 1: the same value if it repeats, else the smallest key above it
  int c_le = 0;
  uint32_t mn = 0xffffffffu;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const bool valid = lane + 32 * j < W;
    c_le += __popc(__ballot_sync(0xffffffffu, valid && kk[j] <= key));
    mn = (valid && kk[j] > key) ? min(mn, kk[j]) : mn;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
  const uint32_t key_hi = (hi == lo || c_le > lo + 1) ? key : mn;
  a = __uint_as_float((key & 0x80000000u) ? (key ^ 0x80000000u) : ~key);
  b = __uint_as_float((key_hi & 0x80000000u) ? (key_hi ^ 0x80000000u) : ~key_hi);
}

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(kThreads, 1)
spec_augment_kernel(const float* __restrict__ in, int H, int W, float mean, float stdv, const afs_specaug_cfg cfg,
                    const float* __restrict__ curve, float* __restrict__ out) {
  extern __shared__ __align__(16) float s_x[];           // [H*W] de-normalised plane
  uint32_t* s_key = reinterpret_cast<uint32_t*>(s_x + H * W);  // [H*W] |x| (or |z|) bit patterns
  float* s_col = reinterpret_cast<float*>(s_key + H * W);       // [W + 16] per-frame scratch
  __shared__ __align__(8) int s_hist[kBins];
  __shared__ int s_pick[4 + 2 * kWarps];  // picks, then per-warp scratch of radix_select / next_order_stat
  __shared__ float s_red[kWarps];

  const int n = H * W;
  const int tid = threadIdx.x;
  const float* src = in + static_cast<int64_t>(blockIdx.x) * n;
  float* dst = out + static_cast<int64_t>(blockIdx.x) * n;
  const int type = cfg.type;

  // 128-bit accesses when the plane allows it (H * W % 4 == 0 and a 16-byte aligned batch: every plane then is)
  const bool vec = (n & 3) == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (vec) {
    const float4* src4 = reinterpret_cast<const float4*>(src);
    for (int i = tid; i < n / 4; i += kThreads) {
      const float4 v = __ldg(src4 + i);
      float4 x;
      x.x = __fadd_rn(__fmul_rn(v.x, stdv), mean);  // denormalize_spectrogram :33
      x.y = __fadd_rn(__fmul_rn(v.y, stdv), mean);
      x.z = __fadd_rn(__fmul_rn(v.z, stdv), mean);
      x.w = __fadd_rn(__fmul_rn(v.w, stdv), mean);
      reinterpret_cast<float4*>(s_x)[i] = x;
      reinterpret_cast<uint4*>(s_key)[i] = make_uint4(__float_as_uint(fabsf(x.x)), __float_as_uint(fabsf(x.y)),
                                                      __float_as_uint(fabsf(x.z)), __float_as_uint(fabsf(x.w)));
    }
  } else {
    for (int i = tid; i < n; i += kThreads) {
      const float x = __fadd_rn(__fmul_rn(src[i], stdv), mean);  // denormalize_spectrogram :33
      s_x[i] = x;
      s_key[i] = __float_as_uint(fabsf(x));
    }
  }
  __syncthreads();

  if (type == AFS_AUG_CUTOUT) {
    // only the rectangles are touched (later rectangles overwrite earlier ones with the same fill: order is irrelevant)
    for (int k = 0; k < cfg.n_rect; ++k) {
      const int r0 = max(cfg.rect[k][0], 0), c0 = max(cfg.rect[k][1], 0);
      const int r1 = min(cfg.rect[k][0] + cfg.rect[k][2], H), c1 = min(cfg.rect[k][1] + cfg.rect[k][3], W);
      const int hh = r1 - r0, ww = c1 - c0;
      if (hh <= 0 || ww <= 0) continue;
      for (int idx = tid; idx < hh * ww; idx += kThreads) {
        const int rr = idx / ww;
        s_x[(r0 + rr) * W + c0 + (idx - rr * ww)] = cfg.fill;
      }
    }
  } else if (type == AFS_AUG_LINEAR_FILTER) {
    for (int i = tid; i < n; i += kThreads) s_x[i] = s_x[i] * curve[i / W];
  } else if (type == AFS_AUG_NOISE_SUPPRESSION) {
    const float thr = quantile_abs(s_key, n, cfg.p0, s_hist, s_pick);
    const float den = thr * 0.1f + 1e-8f;
    for (int i = tid; i < n; i += kThreads) {
      const float x = s_x[i];
      const float m = sigmoidf((fabsf(x) - thr) / den);
      s_x[i] = x * (1.0f - cfg.p1 * (1.0f - m));
    }
  } else if (type == AFS_AUG_NOISE_MATCHING) {
    // per-frame noise floor: min over frequency of |x|, optionally box-smoothed with reflect padding
    for (int c = tid; c < W; c += kThreads) {
      float m = INFINITY;
      for (int r = 0; r < H; ++r) m = fminf(m, fabsf(s_x[r * W + c]));
      s_col[c] = m;
    }
    __syncthreads();
    const int win = cfg.i0;
    float part = 0.f;
    if (win > 1 && W > win) {
      const int pad = win / 2;
      const int out_w = W + 2 * pad - win + 1;  // == W for odd windows
      const float kw = 1.0f / static_cast<float>(win);
      for (int c = tid; c < out_w; c += kThreads) {
        float s = 0.f;
        for (int j = 0; j < win; ++j) {
          int idx = c + j - pad;
          if (idx < 0) idx = -idx;
          if (idx >= W) idx = 2 * (W - 1) - idx;
          s = fmaf(s_col[idx], kw, s);
        }
        part += s;
      }
      part = block_sum(part, s_red) / static_cast<float>(out_w);
    } else {
      for (int c = tid; c < W; c += kThreads) part += s_col[c];
      part = block_sum(part, s_red) / static_cast<float>(W);
    }
    const float cur = part;
    float scale = 1.0f;
    if (cur > 1e-8f) scale = fminf(fmaxf(cfg.p0 / (cur + 1e-8f), 0.5f), 2.0f);
    const float thr = quantile_abs(s_key, n, 0.3f, s_hist, s_pick);
    const float den = thr * 0.1f + 1e-8f;
    for (int i = tid; i < n; i += kThreads) {
      const float x = s_x[i];
      const float m = sigmoidf((fabsf(x) - thr) / den);
      s_x[i] = x * (m + (1.0f - m) * scale);
    }
  } else if (type == AFS_AUG_BACKGROUND_SUBTRACTION) {
    // per frequency row: quantile along time, one warp per row.  Rows of up to 256 frames are held in registers as
    // order-preserving integer keys and the order statistic is found bit by bit (32 ballot rounds) -- exact, like the
    // rank counting it replaces (which costs W comparisons per element and remains the path for longer rows).
    const float rank = __fmul_rn(cfg.p0, static_cast<float>(W - 1));
    const int lo = static_cast<int>(floorf(rank)), hi = static_cast<int>(ceilf(rank));
    const float w = rank - static_cast<float>(lo);
    const int lane = tid & 31, warp = tid >> 5;
    for (int r = warp; r < H; r += kWarps) {
      const float* row = s_x + r * W;
      float a = 0.f, b = 0.f;
      if (W <= 160) {
        row_order_stats<5>(row, W, lo, hi, lane, a, b);
      } else if (W <= 256) {
        row_order_stats<8>(row, W, lo, hi, lane, a, b);
      } else {
        for (int i = lane; i < W; i += 32) {
          const float v = row[i];
          int rk = 0;
          for (int j = 0; j < W; ++j) {
            const float u = row[j];
            rk += (u < v || (u == v && j < i)) ? 1 : 0;
          }
          if (rk == lo) a = v;
          if (rk == hi) b = v;
        }
        // exactly one lane found each; combine (values may be negative: use add of zeros elsewhere)
        a = warp_sum(a);
        b = warp_sum(b);
      }
      const float d = b - a;
      const float bg = hi == lo ? a : (w < 0.5f ? __fmaf_rn(w, d, a) : b - d * (1.0f - w));
      __syncwarp();
      for (int i = lane; i < W; i += 32) {
        const float v = row[i] - bg;
        s_key[r * W + i] = __float_as_uint(fmaxf(v, 0.0f));  // staged: rows still being ranked read s_x
      }
    }
    __syncthreads();
    for (int i = tid; i < n; i += kThreads) s_x[i] = __uint_as_float(s_key[i]);
  } else if (type == AFS_AUG_CONTRAST) {
    float part = 0.f;
    for (int i = tid; i < n; i += kThreads) part += s_x[i];
    const float m = block_sum(part, s_red) / static_cast<float>(n);
    for (int i = tid; i < n; i += kThreads) {
      const float z = m + (s_x[i] - m) * cfg.p0;
      s_x[i] = z;
      s_key[i] = __float_as_uint(fabsf(z));
    }
    __syncthreads();
    if (cfg.p1 < 1.0f) {
      const float mx = quantile_abs(s_key, n, cfg.p1, s_hist, s_pick);
      for (int i = tid; i < n; i += kThreads) s_x[i] = fminf(fmaxf(s_x[i], -mx), mx);
    }
  } else if (type == AFS_AUG_FOREGROUND_NORM) {
    const float thr = quantile_abs(s_key, n, cfg.p0, s_hist, s_pick);
    float cnt = 0.f, sum = 0.f;
    for (int i = tid; i < n; i += kThreads) {
      const float x = s_x[i];
      if (fabsf(x) >= thr) { cnt += 1.f; sum += x; }
    }
    cnt = block_sum(cnt, s_red);
    sum = block_sum(sum, s_red);
    if (cnt > 0.f) {
      const float fm = sum / cnt;
      float ss = 0.f;
      for (int i = tid; i < n; i += kThreads) {
        const float x = s_x[i];
        if (fabsf(x) >= thr) { const float d = x - fm; ss = fmaf(d, d, ss); }
      }
      ss = block_sum(ss, s_red);
      const float sd = sqrtf(ss / (cnt - 1.0f)) + 1e-8f;  // torch.std: unbiased (NaN for a single value, as torch)
      for (int i = tid; i < n; i += kThreads) s_x[i] = (s_x[i] - fm) / sd;
    }
  } else if (type == AFS_AUG_WIENER) {
    const float ne = quantile_abs(s_key, n, cfg.p0, s_hist, s_pick) + 1e-8f;
    for (int i = tid; i < n; i += kThreads) {
      const float x = s_x[i];
      const float snr = fabsf(x) / ne;
      s_x[i] = x * (snr / (snr + 1.0f) * cfg.p1);
    }
  }
  __syncthreads();
  if (vec) {
    float4* dst4 = reinterpret_cast<float4*>(dst);
    for (int i = tid; i < n / 4; i += kThreads) {
      const float4 x = reinterpret_cast<const float4*>(s_x)[i];
      dst4[i] = make_float4((x.x - mean) / stdv, (x.y - mean) / stdv, (x.z - mean) / stdv, (x.w - mean) / stdv);
    }
  } else {
    for (int i = tid; i < n; i += kThreads) dst[i] = (s_x[i] - mean) / stdv;  // normalize_spectrogram :53
  }
}

}  // namespace
}  // namespace afs

extern "C" int afs_spec_augment(const float* in, int32_t planes, int32_t H, int32_t W, float mean, float std,
                                const afs_specaug_cfg* cfg, const float* filter_curve, float* out,
                                afs_stream_t stream_) {
  using namespace afs;
  if (in == nullptr || out == nullptr || cfg == nullptr || planes < 0 || H < 1 || W < 1 || std == 0.f)
    return AFS_ERR_INVALID_ARG;
  if (cfg->type < AFS_AUG_CUTOUT || cfg->type > AFS_AUG_WIENER) return AFS_ERR_INVALID_ARG;
  if (cfg->type == AFS_AUG_LINEAR_FILTER && filter_curve == nullptr) return AFS_ERR_INVALID_ARG;
  if (cfg->type == AFS_AUG_CUTOUT && (cfg->n_rect < 0 || cfg->n_rect > 8)) return AFS_ERR_INVALID_ARG;
  if (planes == 0) return AFS_OK;
  const size_t smem = (2 * static_cast<size_t>(H) * W + W + 16) * sizeof(float);
  if (smem > 200 * 1024) return AFS_ERR_UNSUPPORTED;  // plane must fit in shared memory
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  AFS_CUDA_TRY(cudaFuncSetAttribute(spec_augment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(smem)));
  spec_augment_kernel<<<planes, kThreads, smem, stream>>>(in, H, W, mean, std, *cfg, filter_curve, out);
  AFS_LAUNCH_CHECK();
  return AFS_OK;
}
