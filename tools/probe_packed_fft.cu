// Static probe (no GPU needed): how many SASS instructions do packed-f32x2 formulations of the log-mel
// kernel's radix-8 butterfly and twiddle products cost against the scalar ones in csrc/logmel_core.cuh?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I audio_fewshot_b200/csrc -cubin \
//        -o /tmp/probe.cubin tools/probe_packed_fft.cu && cuobjdump -sass /tmp/probe.cubin | python tools/sass_count.py
// Each kernel loads 8 complex values, applies ONE formulation, stores them: the difference between two kernels'
// instruction counts is the difference between the formulations (loads/stores/addressing are identical).
#include <cuda_runtime.h>

#ifndef CMUL_FORM
#define CMUL_FORM 1
#endif

#include "logmel_core.cuh"

using namespace afs::logmel;

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, f2(-b.x, -b.y)); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// (A) scalar reference
__global__ void k_dft8_scalar(const float2* in, float2* out) {
  cpx a[8];
  const int t = threadIdx.x + blockIdx.x * blockDim.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float2 v = in[t + 64 * i]; a[i].re = v.x; a[i].im = v.y; }
  dft8(a);
#pragma unroll
  for (int i = 0; i < 8; ++i) out[t + 64 * i] = f2(a[i].re, a[i].im);
}

// (B) lanes = (re, im) of one element: every complex add is one FADD2; a multiplication by -i swaps the halves.
__device__ __forceinline__ float2 mul_negi(float2 x) { return f2(x.y, -x.x); }  // x * (-i)
__device__ __forceinline__ void dft8_reim(float2 (&a)[8]) {
  const float h = 0.70710678118654752440f;
  float2 b0 = add2(a[0], a[4]), b4 = sub2(a[0], a[4]);
  float2 b1 = add2(a[1], a[5]), b5 = sub2(a[1], a[5]);
  float2 b2 = add2(a[2], a[6]), b6 = sub2(a[2], a[6]);
  float2 b3 = add2(a[3], a[7]), b7 = sub2(a[3], a[7]);
  b5 = mul2(add2(b5, mul_negi(b5)), f2(h, h));            // (1-i)/sqrt2 * b5
  b6 = mul_negi(b6);
  b7 = mul2(sub2(mul_negi(b7), b7), f2(h, h));            // (-1-i)/sqrt2 * b7
  float2 d0 = add2(b0, b2), d2 = sub2(b0, b2), d1 = add2(b1, b3), d3 = mul_negi(sub2(b1, b3));
  a[0] = add2(d0, d1); a[4] = sub2(d0, d1); a[2] = add2(d2, d3); a[6] = sub2(d2, d3);
  d0 = add2(b4, b6); d2 = sub2(b4, b6); d1 = add2(b5, b7); d3 = mul_negi(sub2(b5, b7));
  a[1] = add2(d0, d1); a[5] = sub2(d0, d1); a[3] = add2(d2, d3); a[7] = sub2(d2, d3);
}
__global__ void k_dft8_reim(const float2* in, float2* out) {
  float2 a[8];
  const int t = threadIdx.x + blockIdx.x * blockDim.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = in[t + 64 * i];
  dft8_reim(a);
#pragma unroll
  for (int i = 0; i < 8; ++i) out[t + 64 * i] = a[i];
}

// (C) lanes = the even and the odd DFT-4 of the radix-8 (identical structure): stage 1 and the W8 twiddles scalar,
// written straight into the pair halves; stages 2-3 packed, multiplications by -i stay a free re<->im relabelling.
__device__ __forceinline__ void dft8_halves(cpx (&a)[8]) {
  const float h = 0.70710678118654752440f;
  // stage 1 (scalar): lane x = b_k = a_k + a_{k+4}, lane y = b_{k+4} = (a_k - a_{k+4}) * W8^k
  float2 re[4], im[4];
  re[0] = f2(a[0].re + a[4].re, a[0].re - a[4].re); im[0] = f2(a[0].im + a[4].im, a[0].im - a[4].im);
  {
    const float dr = a[1].re - a[5].re, di = a[1].im - a[5].im;
    re[1] = f2(a[1].re + a[5].re, (dr + di) * h); im[1] = f2(a[1].im + a[5].im, (di - dr) * h);
  }
  re[2] = f2(a[2].re + a[6].re, a[2].im - a[6].im); im[2] = f2(a[2].im + a[6].im, -(a[2].re - a[6].re));
  {
    const float dr = a[3].re - a[7].re, di = a[3].im - a[7].im;
    re[3] = f2(a[3].re + a[7].re, (di - dr) * h); im[3] = f2(a[3].im + a[7].im, -(dr + di) * h);
  }
  // stages 2-3 (packed): DFT-4 of (b0..b3 | b4..b7)
  const float2 d0r = add2(re[0], re[2]), d0i = add2(im[0], im[2]);
  const float2 d2r = sub2(re[0], re[2]), d2i = sub2(im[0], im[2]);
  const float2 d1r = add2(re[1], re[3]), d1i = add2(im[1], im[3]);
  const float2 d3r = sub2(im[1], im[3]), d3i = sub2(re[3], re[1]);  // (b1 - b3) * (-i)
  const float2 o0r = add2(d0r, d1r), o0i = add2(d0i, d1i);          // a[0] | a[1]
  const float2 o2r = sub2(d0r, d1r), o2i = sub2(d0i, d1i);          // a[4] | a[5]
  const float2 o1r = add2(d2r, d3r), o1i = add2(d2i, d3i);          // a[2] | a[3]
  const float2 o3r = sub2(d2r, d3r), o3i = sub2(d2i, d3i);          // a[6] | a[7]
  a[0].re = o0r.x; a[0].im = o0i.x; a[1].re = o0r.y; a[1].im = o0i.y;
  a[4].re = o2r.x; a[4].im = o2i.x; a[5].re = o2r.y; a[5].im = o2i.y;
  a[2].re = o1r.x; a[2].im = o1i.x; a[3].re = o1r.y; a[3].im = o1i.y;
  a[6].re = o3r.x; a[6].im = o3i.x; a[7].re = o3r.y; a[7].im = o3i.y;
}
__global__ void k_dft8_halves(const float2* in, float2* out) {
  cpx a[8];
  const int t = threadIdx.x + blockIdx.x * blockDim.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float2 v = in[t + 64 * i]; a[i].re = v.x; a[i].im = v.y; }
  dft8_halves(a);
#pragma unroll
  for (int i = 0; i < 8; ++i) out[t + 64 * i] = f2(a[i].re, a[i].im);
}

// Twiddle products of phase A/B: v[q] *= w^q, q = 1..7, powers rebuilt from w (powers7).
__global__ void k_twiddle_scalar(const float2* in, float2* out, const float2* wtab) {
  cpx a[8];
  const int t = threadIdx.x + blockIdx.x * blockDim.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float2 v = in[t + 64 * i]; a[i].re = v.x; a[i].im = v.y; }
  cpx w; w.re = wtab[t].x; w.im = wtab[t].y;
  cpx pw[8];
  powers7(w, pw);
#pragma unroll
  for (int q = 1; q < 8; ++q) a[q] = cmul(a[q], pw[q]);
#pragma unroll
  for (int i = 0; i < 8; ++i) out[t + 64 * i] = f2(a[i].re, a[i].im);
}

// Packed: elements (q, q+1) in the two lanes, re and im in separate pairs: a complex product of two elements by two
// twiddles is 4 packed instructions (2 per product instead of 4).  Powers: p2 scalar, (p2,p3) = p2 * (1, w) ...
__device__ __forceinline__ void cmul2(float2 ar, float2 ai, float2 br, float2 bi, float2& cr, float2& ci) {
  cr = fma2(ar, br, mul2(f2(-ai.x, -ai.y), bi));
  ci = fma2(ar, bi, mul2(ai, br));
}
__global__ void k_twiddle_packed(const float2* in, float2* out, const float2* wtab) {
  const int t = threadIdx.x + blockIdx.x * blockDim.x;
  float2 vr[4], vi[4];  // (v0,v1), (v2,v3), (v4,v5), (v6,v7)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 e = in[t + 64 * (2 * i)], o = in[t + 64 * (2 * i + 1)];
    vr[i] = f2(e.x, o.x); vi[i] = f2(e.y, o.y);
  }
  const float wr = wtab[t].x, wi = wtab[t].y;
  // p2 = w^2 (scalar), then pairs by packed products with (w2, w2): (p2,p3) = (w,w2)*... built as:
  const float p2r = wr * wr - wi * wi, p2i = 2.f * wr * wi;
  float2 p01r = f2(1.f, wr), p01i = f2(0.f, wi);                       // (w^0, w^1)
  float2 p23r, p23i, p45r, p45i, p67r, p67i;
  cmul2(p01r, p01i, f2(p2r, p2r), f2(p2i, p2i), p23r, p23i);          // (w^2, w^3)
  const float p4r = p2r * p2r - p2i * p2i, p4i = 2.f * p2r * p2i;
  cmul2(p01r, p01i, f2(p4r, p4r), f2(p4i, p4i), p45r, p45i);          // (w^4, w^5)
  cmul2(p23r, p23i, f2(p4r, p4r), f2(p4i, p4i), p67r, p67i);          // (w^6, w^7)
  float2 r, i;
  cmul2(vr[0], vi[0], p01r, p01i, r, i); vr[0] = r; vi[0] = i;
  cmul2(vr[1], vi[1], p23r, p23i, r, i); vr[1] = r; vi[1] = i;
  cmul2(vr[2], vi[2], p45r, p45i, r, i); vr[2] = r; vi[2] = i;
  cmul2(vr[3], vi[3], p67r, p67i, r, i); vr[3] = r; vi[3] = i;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    out[t + 64 * (2 * k)] = f2(vr[k].x, vi[k].x);
    out[t + 64 * (2 * k + 1)] = f2(vr[k].y, vi[k].y);
  }
}

// (D) twiddle products with lanes = (re, im): c = (ar*br - ai*bi, ar*bi + ai*br)
//     = fma2((ar,ar), (br,bi), (ai,ai) * (-bi, br)); the broadcasts and the swap are operand modifiers if the ISA has them.
__device__ __forceinline__ float2 cmul_reim(float2 a, float2 b) {
#if CMUL_FORM == 0
  return fma2(f2(a.x, a.x), b, mul2(f2(a.y, a.y), f2(-b.y, b.x)));
#elif CMUL_FORM == 1
  const float2 t = mul2(f2(a.y, a.y), b);  // (ai br, ai bi)
  return fma2(f2(a.x, a.x), b, f2(-t.y, t.x));
#else
  const float2 t = mul2(f2(a.x, a.x), b);  // (ar br, ar bi)
  const float2 u = mul2(f2(a.y, a.y), b);  // (ai br, ai bi)
  return add2(t, f2(-u.y, u.x));
#endif
}
__global__ void k_twiddle_reim(const float2* in, float2* out, const float2* wtab) {
  float2 a[8];
  const int t = threadIdx.x + blockIdx.x * blockDim.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = in[t + 64 * i];
  const float2 w = wtab[t];
  float2 p[8];
  p[1] = w;
  p[2] = cmul_reim(w, w);
  p[3] = cmul_reim(p[2], w);
  p[4] = cmul_reim(p[2], p[2]);
  p[5] = cmul_reim(p[4], w);
  p[6] = cmul_reim(p[4], p[2]);
  p[7] = cmul_reim(p[4], p[3]);
#pragma unroll
  for (int q = 1; q < 8; ++q) a[q] = cmul_reim(a[q], p[q]);
#pragma unroll
  for (int i = 0; i < 8; ++i) out[t + 64 * i] = a[i];
}
