// FFT phases of the fused log-mel kernel (logmel.cu) on packed-f32x2 arithmetic.
//
// sm_100a has FADD2 / FMUL2 / FFMA2 on aligned register pairs, with operand modifiers for a half swap (.LO_HI), a
// one-lane negation (.NP) and a scalar broadcast (Rn.F32) -- so with a complex number held as the pair (re, im):
//   complex add / sub            = 1 instruction (2 scalar)
//   multiplication by -i         = free (swap + one-lane negate fold into the consumer)
//   complex multiplication       = 2 instructions (FMUL2 + FFMA2; 4 scalar)
//   radix-8 butterfly            = 27 instructions (54 scalar)      [static counts: tools/probe_packed_fft.cu]
// and every shared-memory exchange moves (re, im) as one 64-bit word: half the LDS/STS instructions for the same
// wavefronts.  Measured on B200 against the scalar phases this file replaced in round 2: 0.936 -> 0.798 ms per
// 3 200 clips (profiles/r02_logmel_variants.txt), same tolerance against the float64 spec.
//
// Layouts (float2 units; a 64-bit access is served per half-warp, 16 lanes x 8 B = one 128 B wavefront when the 16
// words fall into 16 different 8-byte banks -- checked for every access pattern by tests/emul):
//   exchange 1: element (q, j) at q * 66 + j             (66 == 2 mod 16: the read side q + 8 j0 is conflict-free)
//   exchange 2: element (q, j0, p0) at q + 8 ((j0 ^ p0) & 1) + 16 (j0 + 8 (p0 >> 1))     (dense bijection on [0, 512))
//   exchange 3: Z[k] at k (natural order)
// Index derivation: logmel_core.cuh.
#pragma once
#include "logmel_core.cuh"

namespace afs {
namespace logmel {

#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
#define AFS_F32X2 1
#else
#define AFS_F32X2 0
#endif

AFS_HD float2 p_add(float2 a, float2 b) {
#if AFS_F32X2
  return __fadd2_rn(a, b);
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
AFS_HD float2 p_sub(float2 a, float2 b) { return p_add(a, make_float2(-b.x, -b.y)); }
AFS_HD float2 p_mul(float2 a, float2 b) {
#if AFS_F32X2
  return __fmul2_rn(a, b);
#else
  return make_float2(a.x * b.x, a.y * b.y);
#endif
}
AFS_HD float2 p_fma(float2 a, float2 b, float2 c) {
#if AFS_F32X2
  return __ffma2_rn(a, b, c);
#else
  return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
AFS_HD float2 c_negi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)
AFS_HD float2 c_conj(float2 a) { return make_float2(a.x, -a.y); }
// (ar br - ai bi, ar bi + ai br) = (ar, ar) * b + swap_neg((ai, ai) * b)
AFS_HD float2 c_mul(float2 a, float2 b) {
  const float2 t = p_mul(make_float2(a.y, a.y), b);
  return p_fma(make_float2(a.x, a.x), b, make_float2(-t.y, t.x));
}

// In-place forward 8-point DFT, natural order out: a[q] = sum_r a[r] W_8^{rq}.
AFS_HD void dft8_p(float2 (&a)[8]) {
  const float h = 0.70710678118654752440f;
  const float2 hh = make_float2(h, h);
  const float2 b0 = p_add(a[0], a[4]), b4 = p_sub(a[0], a[4]);
  const float2 b1 = p_add(a[1], a[5]);
  float2 b5 = p_sub(a[1], a[5]);
  const float2 b2 = p_add(a[2], a[6]);
  float2 b6 = p_sub(a[2], a[6]);
  const float2 b3 = p_add(a[3], a[7]);
  float2 b7 = p_sub(a[3], a[7]);
  b5 = p_mul(p_add(b5, c_negi(b5)), hh);  // W8^1 = (1 - i)/sqrt2
  b6 = c_negi(b6);                        // W8^2 = -i
  b7 = p_mul(p_sub(c_negi(b7), b7), hh);  // W8^3 = (-1 - i)/sqrt2
  float2 d0 = p_add(b0, b2), d2 = p_sub(b0, b2), d1 = p_add(b1, b3), d3 = c_negi(p_sub(b1, b3));
  a[0] = p_add(d0, d1); a[4] = p_sub(d0, d1); a[2] = p_add(d2, d3); a[6] = p_sub(d2, d3);
  d0 = p_add(b4, b6); d2 = p_sub(b4, b6); d1 = p_add(b5, b7); d3 = c_negi(p_sub(b5, b7));
  a[1] = p_add(d0, d1); a[5] = p_sub(d0, d1); a[3] = p_add(d2, d3); a[7] = p_sub(d2, d3);
}

AFS_HD void powers7_p(float2 w, float2 (&p)[8]) {
  p[1] = w;
  p[2] = c_mul(w, w);
  p[3] = c_mul(p[2], w);
  p[4] = c_mul(p[2], p[2]);
  p[5] = c_mul(p[4], w);
  p[6] = c_mul(p[4], p[2]);
  p[7] = c_mul(p[4], p[3]);
}

constexpr int kE1StrideP = 66;  // float2 units
static_assert(2 * 8 * kE1StrideP <= kBufA, "packed exchange 1 must fit in buffer A");
AFS_HD int e1p_slot(int q, int j) { return q * kE1StrideP + j; }
AFS_HD int e2p_slot(int q, int j0, int p0) { return q + 8 * ((j0 ^ p0) & 1) + 16 * (j0 + 8 * (p0 >> 1)); }

// Phase A. in: z[r] = windowed (x[2n], x[2n+1]) as (re, im), n = j + 64 r.
AFS_HD void phase_a_p(int j, float2 (&z)[8], const ThreadTw& tw, float2* bufA) {
  dft8_p(z);
  bufA[e1p_slot(0, j)] = z[0];
  float2 pw[8];
  powers7_p(make_float2(tw.a.re, tw.a.im), pw);
#pragma unroll
  for (int q = 1; q < 8; ++q) bufA[e1p_slot(q, j)] = c_mul(z[q], pw[q]);
}

// Phase B. thread t = q + 8*j0.
AFS_HD void phase_b_p(int t, const ThreadTw& tw, const float2* bufA, float2* bufB) {
  const int q = t & 7, j0 = t >> 3;
  float2 v[8];
#pragma unroll
  for (int j1 = 0; j1 < 8; ++j1) v[j1] = bufA[e1p_slot(q, j0 + 8 * j1)];
  dft8_p(v);
  float2 pw[8];
  powers7_p(make_float2(tw.b.re, tw.b.im), pw);
  bufB[e2p_slot(q, j0, 0)] = v[0];
#pragma unroll
  for (int p0 = 1; p0 < 8; ++p0) bufB[e2p_slot(q, j0, p0)] = c_mul(v[p0], pw[p0]);
}

// Phase C. thread t = q + 8*p0; leaves Z[t + 64 p1] in natural order in buffer A.
AFS_HD void phase_c_p(int t, const float2* bufB, float2* bufA) {
  const int q = t & 7, p0 = t >> 3;
  float2 v[8];
#pragma unroll
  for (int j0 = 0; j0 < 8; ++j0) v[j0] = bufB[e2p_slot(q, j0, p0)];
  dft8_p(v);
#pragma unroll
  for (int p1 = 0; p1 < 8; ++p1) bufA[t + 64 * p1] = v[p1];
}

// Phase D. thread u handles k = u + 64 m (m = 0..3) and its mirror 512 - k; thread 0 also the self-paired bin 256.
AFS_HD void phase_d_p(int u, const ThreadTw& tw, const float2* bufA, float* power) {
  const float w16re[4] = {1.0f, 0.92387953251128673848f, 0.70710678118654752440f, 0.38268343236508978178f};
  const float w16im[4] = {0.0f, -0.38268343236508978178f, -0.70710678118654752440f, -0.92387953251128673848f};
  const float2 d = make_float2(tw.d.re, tw.d.im);
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int k = u + 64 * m;
    const int kb = (kHalf - k) & (kHalf - 1);
    const float2 a = bufA[k];
    const float2 cb = c_conj(bufA[kb]);
    const float2 e2 = p_add(a, cb);           // 2E = A + conj(B)
    const float2 o2 = c_negi(p_sub(a, cb));   // 2O = (A - conj(B)) / i
    float2 td = d;                            // W_1024^(u + 64 m) = W_1024^u * W_16^m
    if (m > 0) td = c_mul(d, make_float2(w16re[m], w16im[m]));
    const float2 t2 = c_mul(o2, td);
    const float2 x = p_add(e2, t2), y = p_sub(e2, t2);
    power[k] = 0.25f * (x.x * x.x + x.y * x.y);
    power[kHalf - k] = 0.25f * (y.x * y.x + y.y * y.y);
  }
  if (u == 0) {
    const float2 a = bufA[256];
    power[256] = a.x * a.x + a.y * a.y;
  }
}

// mel_dot_batch with the four frames as two register pairs: one FFMA2 per pair and weight (the weight is a scalar
// broadcast operand, the two power values are the destinations of two LDS); same products and sums as the scalar loop.
template <int PS = kPStride>
AFS_HD void mel_dot_batch_p(const float* power, const float* weights, int wstride, int lo, int len,
                            float (&acc)[kMelBatch]) {
  static_assert(kMelBatch == 4, "two pairs of frames");
  float2 a01 = make_float2(0.f, 0.f), a23 = make_float2(0.f, 0.f);
  const float* p0 = power + lo;
  for (int i = 0; i < len; ++i) {
    const float w = weights[i * wstride];
    const float2 ww = make_float2(w, w);
    a01 = p_fma(ww, make_float2(p0[i], p0[PS + i]), a01);
    a23 = p_fma(ww, make_float2(p0[2 * PS + i], p0[3 * PS + i]), a23);
  }
  acc[0] = a01.x; acc[1] = a01.y; acc[2] = a23.x; acc[3] = a23.y;
}

}  // namespace logmel
}  // namespace afs
