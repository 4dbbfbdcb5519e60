// Per-thread phases of the fused waveform -> log-mel kernel (logmel.cu).
//
// A frame of n_fft = 1024 real samples is transformed by ONE group of 64
// threads: z[n] = x[2n] + i x[2n+1] (512 complex points), a 512-point complex
// FFT as three register-resident radix-8 passes (512 = 8*8*8) with two
// shared-memory exchanges, then the real-input split that yields the 513
// power bins, then the banded mel projection, log and normalisation.
//
// Every phase is a __host__ __device__ function of (thread index in group,
// group-private shared buffers), so the exact index arithmetic the kernel runs
// is also executed on the CPU by tests/emul/logmel_emul.cu (64 threads looped
// per phase).  Index derivation, with W_n = exp(-2 pi i / n):
//   n = j + 64 r,  k = q + 8 p:   Z[q + 8p] = sum_j W_64^{jp} ( W_512^{jq} sum_r z[j+64r] W_8^{rq} )
//   j = j0 + 8 j1, p = p0 + 8 p1: U_q[p0 + 8 p1] = sum_j0 W_8^{j0 p1} ( W_64^{j0 p0} sum_j1 u_q[j0+8j1] W_8^{j1 p0} )
//   phase A: thread j      holds r = 0..7   -> q;   twiddle W_512^{jq}
//   phase B: thread q+8*j0 holds j1 = 0..7  -> p0;  twiddle W_64^{j0 p0}
//   phase C: thread q+8*p0 holds j0 = 0..7  -> p1;  Z[t + 64 p1], natural order
//   phase D: X[k] = E + W_1024^k O,  X[512-k] = conj(E - W_1024^k O),
//            E = (Z[k] + conj Z[512-k])/2,  O = (Z[k] - conj Z[512-k])/(2i)
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define AFS_HD __host__ __device__ __forceinline__
#else
#define AFS_HD inline
#endif

namespace afs {
namespace logmel {

constexpr int kNfft = 1024;
constexpr int kHalf = 512;        // complex FFT length
constexpr int kBins = 513;        // n_fft/2 + 1
constexpr int kGroup = 64;        // threads per frame
constexpr int kMaxMels = 128;
constexpr int kE1Stride = 68;     // exchange-1 row stride (== 4 mod 32: conflict-free reads)
constexpr int kBufA = 2 * 8 * kE1Stride;  // floats: re/im of exchange 1; reused (natural order) by exchange 3
constexpr int kBufB = 2 * kHalf;          // floats: re/im of exchange 2; reused by the 513 power bins
static_assert(kBufA >= 2 * kHalf, "exchange 3 must fit in buffer A");
static_assert(kBufB >= kBins, "power bins must fit in buffer B");

struct cpx {
  float re, im;
};

AFS_HD cpx cmul(cpx a, cpx b) {
  cpx r;
  r.re = a.re * b.re - a.im * b.im;
  r.im = a.re * b.im + a.im * b.re;
  return r;
}

// In-place forward 8-point DFT, natural order out: a[q] = sum_r a[r] W_8^{rq}.
AFS_HD void dft8(cpx (&a)[8]) {
  const float h = 0.70710678118654752440f;
  cpx b0, b1, b2, b3, b4, b5, b6, b7;
  b0.re = a[0].re + a[4].re; b0.im = a[0].im + a[4].im;
  b4.re = a[0].re - a[4].re; b4.im = a[0].im - a[4].im;
  b1.re = a[1].re + a[5].re; b1.im = a[1].im + a[5].im;
  b5.re = a[1].re - a[5].re; b5.im = a[1].im - a[5].im;
  b2.re = a[2].re + a[6].re; b2.im = a[2].im + a[6].im;
  b6.re = a[2].re - a[6].re; b6.im = a[2].im - a[6].im;
  b3.re = a[3].re + a[7].re; b3.im = a[3].im + a[7].im;
  b7.re = a[3].re - a[7].re; b7.im = a[3].im - a[7].im;
  // odd branch twiddles: W8^1 = (1-i)/sqrt2, W8^2 = -i, W8^3 = (-1-i)/sqrt2
  cpx t;
  t.re = (b5.re + b5.im) * h; t.im = (b5.im - b5.re) * h; b5 = t;
  t.re = b6.im; t.im = -b6.re; b6 = t;
  t.re = (b7.im - b7.re) * h; t.im = -(b7.re + b7.im) * h; b7 = t;
  // even outputs: DFT4(b0,b1,b2,b3)
  cpx d0, d1, d2, d3;
  d0.re = b0.re + b2.re; d0.im = b0.im + b2.im;
  d2.re = b0.re - b2.re; d2.im = b0.im - b2.im;
  d1.re = b1.re + b3.re; d1.im = b1.im + b3.im;
  d3.re = b1.im - b3.im; d3.im = -(b1.re - b3.re);  // (b1 - b3) * (-i)
  a[0].re = d0.re + d1.re; a[0].im = d0.im + d1.im;
  a[4].re = d0.re - d1.re; a[4].im = d0.im - d1.im;
  a[2].re = d2.re + d3.re; a[2].im = d2.im + d3.im;
  a[6].re = d2.re - d3.re; a[6].im = d2.im - d3.im;
  // odd outputs: DFT4(b4,b5,b6,b7)
  d0.re = b4.re + b6.re; d0.im = b4.im + b6.im;
  d2.re = b4.re - b6.re; d2.im = b4.im - b6.im;
  d1.re = b5.re + b7.re; d1.im = b5.im + b7.im;
  d3.re = b5.im - b7.im; d3.im = -(b5.re - b7.re);
  a[1].re = d0.re + d1.re; a[1].im = d0.im + d1.im;
  a[5].re = d0.re - d1.re; a[5].im = d0.im - d1.im;
  a[3].re = d2.re + d3.re; a[3].im = d2.im + d3.im;
  a[7].re = d2.re - d3.re; a[7].im = d2.im - d3.im;
}

// exchange-2 slot of element (q, j0, p0): written by thread q+8*j0 (p0 = register
// index), read by thread q+8*p0 (j0 = register index); both patterns touch 32
// distinct banks per warp.  Dense: a bijection onto [0, 512).
AFS_HD int e2_slot(int q, int j0, int p0) {
  return q + 8 * (((j0 & 3) + (p0 & 3)) & 3) + 32 * ((j0 >> 2) + 2 * (p0 >> 2) + 4 * (j0 & 3));
}

// Per-thread twiddle BASES, held in registers for the whole kernel; the powers each phase needs are rebuilt per
// frame with complex multiplications (the FMA pipe has headroom, the shared-memory pipe is the kernel's limiter):
//   a  = W_512^j          phase A uses a^q,  q = 1..7
//   b  = W_64^(t >> 3)    phase B uses b^p0, p0 = 1..7
//   d  = W_1024^t         phase D uses d * W_16^m, m = 0..3
struct ThreadTw {
  cpx a, b, d;
};

// w^1 .. w^7 with 6 complex multiplications of depth <= 3 (a few ulp of error, far below the FFT's own rounding)
AFS_HD void powers7(cpx w, cpx (&p)[8]) {
  p[1] = w;
  p[2] = cmul(w, w);
  p[3] = cmul(p[2], w);
  p[4] = cmul(p[2], p[2]);
  p[5] = cmul(p[4], w);
  p[6] = cmul(p[4], p[2]);
  p[7] = cmul(p[4], p[3]);
}

// tw1024[k] = (cos(2 pi k/1024), -sin(2 pi k/1024)), k in [0, 1024)
AFS_HD void load_thread_tw(ThreadTw& tw, int t, const float2* tw1024) {
  const float2 wa = tw1024[(2 * t) & 1023], wb = tw1024[(16 * (t >> 3)) & 1023], wd = tw1024[t];
  tw.a.re = wa.x; tw.a.im = wa.y;
  tw.b.re = wb.x; tw.b.im = wb.y;
  tw.d.re = wd.x; tw.d.im = wd.y;
}

// Phase A. in: z[r] = windowed (x[2n], x[2n+1]), n = j + 64 r.  out: exchange 1.
AFS_HD void phase_a(int j, cpx (&z)[8], const ThreadTw& tw, float* bufA) {
  dft8(z);
  float* re = bufA;
  float* im = bufA + 8 * kE1Stride;
  re[j] = z[0].re;
  im[j] = z[0].im;
  cpx pw[8];
  powers7(tw.a, pw);
#pragma unroll
  for (int q = 1; q < 8; ++q) {
    const cpx v = cmul(z[q], pw[q]);
    re[q * kE1Stride + j] = v.re;
    im[q * kE1Stride + j] = v.im;
  }
}

// Phase B. thread t = q + 8*j0.
AFS_HD void phase_b(int t, const ThreadTw& tw, const float* bufA, float* bufB) {
  const int q = t & 7, j0 = t >> 3;
  const float* re = bufA;
  const float* im = bufA + 8 * kE1Stride;
  cpx v[8];
#pragma unroll
  for (int j1 = 0; j1 < 8; ++j1) {
    v[j1].re = re[q * kE1Stride + j0 + 8 * j1];
    v[j1].im = im[q * kE1Stride + j0 + 8 * j1];
  }
  dft8(v);
  float* ore = bufB;
  float* oim = bufB + kHalf;
  cpx pw[8];
  powers7(tw.b, pw);
#pragma unroll
  for (int p0 = 0; p0 < 8; ++p0) {
    cpx w = v[0];
    if (p0 > 0) w = cmul(v[p0], pw[p0]);
    const int s = e2_slot(q, j0, p0);
    ore[s] = w.re;
    oim[s] = w.im;
  }
}

// Phase C. thread t = q + 8*p0; leaves Z[t + 64 p1] in natural order in buffer A.
AFS_HD void phase_c(int t, const float* bufB, float* bufA) {
  const int q = t & 7, p0 = t >> 3;
  const float* re = bufB;
  const float* im = bufB + kHalf;
  cpx v[8];
#pragma unroll
  for (int j0 = 0; j0 < 8; ++j0) {
    const int s = e2_slot(q, j0, p0);
    v[j0].re = re[s];
    v[j0].im = im[s];
  }
  dft8(v);
  float* ore = bufA;
  float* oim = bufA + kHalf;
#pragma unroll
  for (int p1 = 0; p1 < 8; ++p1) {
    ore[t + 64 * p1] = v[p1].re;
    oim[t + 64 * p1] = v[p1].im;
  }
}

// Phase D. thread u handles k = u + 64 m (m = 0..3) and its mirror 512 - k;
// thread 0 also writes the self-paired bin 256.  Power spectrum into buffer B.
AFS_HD void phase_d(int u, const ThreadTw& tw, const float* bufA, float* power) {
  // W_16^m = exp(-2 pi i m / 16), m = 0..3
  const float w16re[4] = {1.0f, 0.92387953251128673848f, 0.70710678118654752440f, 0.38268343236508978178f};
  const float w16im[4] = {0.0f, -0.38268343236508978178f, -0.70710678118654752440f, -0.92387953251128673848f};
  const float* re = bufA;
  const float* im = bufA + kHalf;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int k = u + 64 * m;
    const int kb = (kHalf - k) & (kHalf - 1);
    const float ar = re[k], ai = im[k];
    const float br = re[kb], bi = im[kb];
    // 2E = A + conj(B); 2O = (A - conj(B)) / i
    const float er = ar + br, ei = ai - bi;
    const float orr = ai + bi, oi = br - ar;
    cpx o2; o2.re = orr; o2.im = oi;
    cpx td = tw.d;  // W_1024^(u + 64 m) = W_1024^u * W_16^m
    if (m > 0) {
      cpx c; c.re = w16re[m]; c.im = w16im[m];
      td = cmul(tw.d, c);
    }
    const cpx t2 = cmul(o2, td);
    const float xr = er + t2.re, xi = ei + t2.im;
    const float yr = er - t2.re, yi = ei - t2.im;
    power[k] = 0.25f * (xr * xr + xi * xi);
    power[kHalf - k] = 0.25f * (yr * yr + yi * yi);
  }
  if (u == 0) {
    const float ar = re[256], ai = im[256];
    power[256] = ar * ar + ai * ai;
  }
}

// Reflect padding index (torch.stft center=True, pad_mode="reflect"); needs pad < L.
AFS_HD int64_t reflect_index(int64_t idx, int64_t L) {
  if (idx < 0) idx = -idx;
  if (idx >= L) idx = 2 * (L - 1) - idx;
  return idx;
}

// Banded mel projection of one filter: sum_i w[i * wstride] * P[lo+i].
AFS_HD float mel_dot(const float* power, const float* weights, int wstride, int lo, int len) {
  float acc = 0.f;
  for (int i = 0; i < len; ++i) acc += weights[i * wstride] * power[lo + i];
  return acc;
}

// The same projection for kMelBatch frames at once (their power spectra kPStride floats apart): one
// weight load feeds kMelBatch FMAs.  Summation order per frame is identical to mel_dot.
constexpr int kMelBatch = 4;
constexpr int kPStride = 516;  // 513 power bins, padded to a multiple of 4 floats
AFS_HD void mel_dot_batch(const float* power, const float* weights, int wstride, int lo, int len,
                          float (&acc)[kMelBatch]) {
#pragma unroll
  for (int f = 0; f < kMelBatch; ++f) acc[f] = 0.f;
  for (int i = 0; i < len; ++i) {
    const float w = weights[i * wstride];
#pragma unroll
    for (int f = 0; f < kMelBatch; ++f) acc[f] += w * power[f * kPStride + lo + i];
  }
}

// log_mult * log10(e) = (log_mult * log10(2)) * log2(e); on the device log2 is the MUFU.LG2 approximation
// (abs error 2^-22 near 1, 2 ulp elsewhere: far inside the 1e-4 dB tolerance), on the host log2f.
AFS_HD float fast_log2(float x) {
#if defined(__CUDA_ARCH__)
  return __log2f(x);
#else
  return log2f(x);
#endif
}

// normalised log-mel value: (log_mult*log10(e + eps) - mean) / std, with scale = log_mult*log10(2)/std and
// shift = -mean/std precomputed per mel bin.
AFS_HD float norm_db(float e, float eps, float scale, float shift) {
  return fast_log2(e + eps) * scale + shift;
}

// Host-side packing of a dense [kBins, n_mels] filterbank into per-filter spans:
// band[m] = first non-zero bin, band[kMaxMels+m] = span length, band[2*kMaxMels+m] = offset
// into `weights`.  Returns the number of packed weights.
template <typename IntVec, typename FloatVec>
inline int pack_mel_bands(const float* fb, int n_mels, IntVec& band, FloatVec& weights) {
  band.assign(3 * kMaxMels, 0);
  weights.clear();
  for (int m = 0; m < n_mels; ++m) {
    int lo = -1, hi = -1;
    for (int k = 0; k < kBins; ++k) {
      if (fb[static_cast<size_t>(k) * n_mels + m] != 0.f) {
        if (lo < 0) lo = k;
        hi = k;
      }
    }
    const int len = lo < 0 ? 0 : hi - lo + 1;
    band[m] = lo < 0 ? 0 : lo;
    band[kMaxMels + m] = len;
    band[2 * kMaxMels + m] = static_cast<int>(weights.size());
    for (int i = 0; i < len; ++i) weights.push_back(fb[static_cast<size_t>(lo + i) * n_mels + m]);
  }
  return static_cast<int>(weights.size());
}

// Shared-memory wavefronts of one warp-pass of the mel projection for one frame: in iteration i lane l reads power
// bin start[l] + i when i < n[l] (all lanes execute iteration i in the same instruction); an instruction costs as
// many wavefronts as the largest number of DISTINCT words that fall on one of the 32 banks.
inline int mel_read_wavefronts(const int (&start)[32], const int (&n)[32]) {
  int total = 0;
  for (int i = 0;; ++i) {
    int words[32][32], cnt[32] = {0}, worst = 0;
    bool any = false;
    for (int l = 0; l < 32; ++l) {
      if (i >= n[l]) continue;
      any = true;
      const int a = start[l] + i, b = a & 31;
      bool seen = false;
      for (int j = 0; j < cnt[b]; ++j) seen = seen || words[b][j] == a;
      if (!seen) words[b][cnt[b]++] = a;
      if (cnt[b] > worst) worst = cnt[b];
    }
    if (!any) break;
    total += worst;
  }
  return total;
}

// The kernel's weight table (ELL layout): thread t of a 64-thread frame group owns filter t (pass 0) and filter
// n_mels-1-t when that is >= 64 (pass 1).  The 32 lanes of a warp read weight i of their filter in the same
// instruction, so the table is stored [warp-pass][i][lane] (zero padded to the longest filter of the warp-pass):
// every weight load is one conflict-free wavefront.
// The POWER reads of that instruction are bins start_l + i with a different start_l per lane; the upper slaney
// filters start 6-12 bins apart, which as it stands puts up to 4 lanes on one bank (102 wavefronts per frame for 43
// instructions).  A filter shorter than the longest one of its warp-pass has slack: it may begin r iterations
// late, i.e. read from bin lo - r with r leading zero weights, which moves its bank by -r in EVERY iteration
// without touching the summation order of its non-zero terms.  The shifts are chosen here, once per plan, by a
// deterministic local search on mel_read_wavefronts (47 wavefronts per frame for the 128-mel slaney bank).
//   band[m] = first bin READ (lo - r), band[kMaxMels+m] = iterations (r + span length),
//   band[2*kMaxMels+m] = index of the filter's table entry 0; consecutive entries are kEllStride apart.
constexpr int kEllStride = 32;
template <typename IntVec, typename FloatVec>
inline int pack_mel_ell(const float* fb, int n_mels, IntVec& band, FloatVec& weights) {
  IntVec b0;
  FloatVec w0;
  pack_mel_bands(fb, n_mels, b0, w0);
  band.assign(3 * kMaxMels, 0);
  weights.clear();
  for (int pass = 0; pass < 2; ++pass) {
    for (int warp = 0; warp < kGroup / 32; ++warp) {
      int filt[32], lo[32], len[32], maxlen = 0;
      for (int lane = 0; lane < 32; ++lane) {
        const int t = warp * 32 + lane;
        int m = pass == 0 ? (t < n_mels ? t : -1) : (n_mels - 1 - t >= kGroup ? n_mels - 1 - t : -1);
        filt[lane] = m;
        lo[lane] = m >= 0 ? b0[m] : 0;
        len[lane] = m >= 0 ? b0[kMaxMels + m] : 0;
        if (len[lane] > maxlen) maxlen = len[lane];
      }
      // start shifts: r <= maxlen - len (the table does not grow) and r <= lo (the first bin read exists)
      int shift[32] = {0}, start[32], n[32];
      for (int lane = 0; lane < 32; ++lane) { start[lane] = lo[lane]; n[lane] = len[lane]; }
      int best = mel_read_wavefronts(start, n);
      uint32_t rng = 0x9E3779B9u + static_cast<uint32_t>(pass * 2 + warp);
      for (int it = 0; it < 6000 && best > maxlen; ++it) {
        rng = rng * 1664525u + 1013904223u;
        const int lane = static_cast<int>((rng >> 8) & 31u);
        if (len[lane] == 0) continue;
        int room = maxlen - len[lane];
        if (lo[lane] < room) room = lo[lane];
        if (room == 0) continue;
        rng = rng * 1664525u + 1013904223u;
        const int cand = static_cast<int>((rng >> 8) % static_cast<uint32_t>(room + 1));
        const int old = shift[lane];
        if (cand == old) continue;
        start[lane] = lo[lane] - cand; n[lane] = len[lane] + cand;
        const int c = mel_read_wavefronts(start, n);
        if (c <= best) {  // sideways moves are accepted: the landscape is full of plateaus
          best = c; shift[lane] = cand;
        } else {
          start[lane] = lo[lane] - old; n[lane] = len[lane] + old;
        }
      }
      const int base = static_cast<int>(weights.size());
      weights.resize(weights.size() + static_cast<size_t>(maxlen) * kEllStride, 0.f);
      for (int lane = 0; lane < 32; ++lane) {
        const int m = filt[lane];
        if (m < 0) continue;
        band[m] = lo[lane] - shift[lane];
        band[kMaxMels + m] = len[lane] + shift[lane];
        band[2 * kMaxMels + m] = base + lane;
        for (int i = 0; i < len[lane]; ++i)
          weights[base + (shift[lane] + i) * kEllStride + lane] = w0[b0[2 * kMaxMels + m] + i];
      }
    }
  }
  return static_cast<int>(weights.size());
}

}  // namespace logmel
}  // namespace afs
