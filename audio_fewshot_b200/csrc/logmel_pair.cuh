// Phases of the warp-per-frame-pair log-mel engine (logmel_pair.cu): ONE WARP transforms TWO frames at once.
//
// z[n] = w[n] (x_f[n] + i x_{f+1}[n]), n = 0..1023, is transformed by one 1024-point complex FFT held as 32 values
// per lane, decomposed 32 x 32 (n = n1 + 32 n2, k = k1 + 32 k2, W_n = exp(-2 pi i / n)):
//   pass 1, lane n1: Y[n1,k1] = sum_n2 z[n1 + 32 n2] W_32^{n2 k1}   (32-point DFT in registers)
//                    U[n1,k1] = Y[n1,k1] W_1024^{n1 k1}             -> shared memory, slot n1 * 33 + k1
//   pass 2, lane k1: X[k1 + 32 k2] = sum_n1 U[n1,k1] W_32^{n1 k2}   (32-point DFT in registers)
//   split:           x_f, x_{f+1} real  =>  2 F_f[k] = X[k] + conj X[1024-k],  2i F_{f+1}[k] = X[k] - conj X[1024-k];
//                    X[1024 - (k1 + 32 k2)] is value 31 - k2 of lane 32 - k1 (lane 0: value (32 - k2) & 31 of itself),
//                    i.e. the partner is IN THE SAME WARP: the split is 16 complex shuffles, not a third exchange.
// Against the 64-thread radix-8 engine (logmel_fft.cuh: three shared-memory exchanges of 512 complex values per frame)
// this is ONE exchange of 1024 complex values per TWO frames: a third of the exchange traffic per frame, no
// group barriers (a warp only ever synchronises with itself), and 32 independent values per thread for the FMA pipe.
//
// Shared-memory bank model (64-bit accesses are served per half-warp on 16 eight-byte banks): pass-1 writes of one
// instruction are slots n1 * 33 + k1 over lanes n1 -- (n1 + k1) mod 16 distinct; pass-2 reads are n1 * 33 + k1 over lanes
// k1 -- contiguous.  Checked by tests/emul (emul_pair_bank_check).  (128-bit stores of value pairs were measured and
// dropped: ptxas needs four MOVs per store to line the registers up.)
//
// Scaling: the kernel stages the window multiplied by 1/2 (exact), so X is half the textbook transform and
// |X[k] +- conj X[1024-k]|^2 IS the power of the frame -- the 1/4 of the split costs nothing.
#pragma once
#include "logmel_fft.cuh"

namespace afs {
namespace logmel {

constexpr int kPairStride = 33;                // float2 slots per exchange row
constexpr int kPairExch = 32 * kPairStride;    // float2 slots of one warp's exchange buffer

// s * (c + i d) with c, d compile-time constants: both enter as broadcast immediates (no constant register pairs).
AFS_HD float2 c_mul_k(float2 s, float c, float d) {
  const float2 t = p_mul(make_float2(-s.y, s.x), make_float2(d, d));
  return p_fma(s, make_float2(c, c), t);
}

AFS_HD void dft4_p(float2& a0, float2& a1, float2& a2, float2& a3) {
  const float2 s0 = p_add(a0, a2), s1 = p_sub(a0, a2), s2 = p_add(a1, a3), s3 = c_negi(p_sub(a1, a3));
  a0 = p_add(s0, s2); a1 = p_add(s1, s3); a2 = p_sub(s0, s2); a3 = p_sub(s1, s3);
}

// In-place forward 32-point DFT, natural order in and out: a[k] = sum_n a[n] W_32^{nk}.
// n = na + 8 nb, k = kb + 4 ka: eight 4-point DFTs over nb, twiddle W_32^{na kb}, four 8-point DFTs over na.
AFS_HD void dft32_p(float2 (&a)[32]) {
  const float wre[22] = {1.00000000000000000000f, 0.98078528040323043058f, 0.92387953251128673848f, 0.83146961230254523567f,
                         0.70710678118654757274f, 0.55557023301960228867f, 0.38268343236508983729f, 0.19509032201612833135f,
                         0.0f, -0.19509032201612819257f, -0.38268343236508972627f, -0.55557023301960195560f,
                         -0.70710678118654746172f, -0.83146961230254534669f, -0.92387953251128673848f, -0.98078528040323043058f,
                         -1.0f, -0.98078528040323043058f, -0.92387953251128684951f, -0.83146961230254545772f,
                         -0.70710678118654768376f, -0.55557023301960217765f};
  const float wim[22] = {0.0f, -0.19509032201612824808f, -0.38268343236508978178f, -0.55557023301960217765f,
                         -0.70710678118654746172f, -0.83146961230254523567f, -0.92387953251128673848f, -0.98078528040323043058f,
                         -1.0f, -0.98078528040323043058f, -0.92387953251128673848f, -0.83146961230254545772f,
                         -0.70710678118654757274f, -0.55557023301960217765f, -0.38268343236508989280f, -0.19509032201612860891f,
                         0.0f, 0.19509032201612835911f, 0.38268343236508967076f, 0.55557023301960195560f,
                         0.70710678118654746172f, 0.83146961230254523567f};
#pragma unroll
  for (int na = 0; na < 8; ++na) dft4_p(a[na], a[na + 8], a[na + 16], a[na + 24]);  // a[na + 8 kb]
  float2 o[32];
#pragma unroll
  for (int kb = 0; kb < 4; ++kb) {
    float2 v[8];
#pragma unroll
    for (int na = 0; na < 8; ++na) {
      const int j = na * kb;
      const float2 s = a[na + 8 * kb];
      v[na] = (j == 0) ? s : (j == 8) ? c_negi(s) : c_mul_k(s, wre[j], wim[j]);
    }
    dft8_p(v);
#pragma unroll
    for (int ka = 0; ka < 8; ++ka) o[kb + 4 * ka] = v[ka];
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) a[i] = o[i];
}

// Pass 1.  in: z[n2] = windowed (x_f[n], x_{f+1}[n]) as (re, im), n = lane + 32 n2; w = W_1024^lane.
// The 31 twiddles W_1024^{lane k1} are rebuilt per pair from w (7 powers held, advanced by w^8 per group of eight):
// 30 complex multiplications, at most 6 roundings deep.
AFS_HD void pair_pass1(int lane, float2 (&z)[32], float2 w, float2* exch) {
  dft32_p(z);
  float2* row = exch + lane * kPairStride;
  float2 pw[8];
  powers7_p(w, pw);
  const float2 w8 = c_mul(pw[4], pw[4]);
  float2 wa = w8;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    row[8 * a] = a == 0 ? z[0] : c_mul(z[8 * a], wa);
#pragma unroll
    for (int b = 1; b < 8; ++b) {
      if (a > 0) pw[b] = c_mul(pw[b], w8);
      row[8 * a + b] = c_mul(z[8 * a + b], pw[b]);
    }
    if (a > 0 && a < 3) wa = c_mul(wa, w8);
  }
}

// Pass 2.  out: x[k2] = X[lane + 32 k2].
AFS_HD void pair_pass2(int lane, const float2* exch, float2 (&x)[32]) {
#pragma unroll
  for (int n1 = 0; n1 < 32; ++n1) x[n1] = exch[n1 * kPairStride + lane];
  dft32_p(x);
}

// Power of bin k of both frames, (|F_f[k]|^2, |F_{f+1}[k]|^2), from a = X[k] and b = X[1024 - k] of the half-scaled
// transform.
AFS_HD float2 pair_power(float2 a, float2 b) {
  const float2 cb = c_conj(b);
  const float2 e = p_add(a, cb);  // F_f[k]
  const float2 o = p_sub(a, cb);  // i F_{f+1}[k]
  return make_float2(fmaf(e.x, e.x, e.y * e.y), fmaf(o.x, o.x, o.y * o.y));
}

// The value a lane SENDS for bin row k2 (k2 = 0..15) of its partner: x[31 - k2], lane 0 (its own partner) x[(32 - k2) & 31].
AFS_HD float2 pair_split_src(int lane, const float2 (&x)[32], int k2) {
  return lane == 0 ? x[(32 - k2) & 31] : x[31 - k2];
}
AFS_HD int pair_partner(int lane) { return (32 - lane) & 31; }

// Banded mel projection of four frames.  A power plane holds, per bin, the powers of one frame PAIR as one 64-bit word:
// pA = frames 0, 1, pB = frames 2, 3 (the second pair's plane overwrites the warp's exchange buffer).  Same products and
// the same summation order per frame as mel_dot_batch_p.
AFS_HD void mel_dot_pairs(const float2* pA, const float2* pB, const float* weights, int lo, int len, float (&acc)[4]) {
  float2 a01 = make_float2(0.f, 0.f), a23 = make_float2(0.f, 0.f);
  const float2* a = pA + lo;
  const float2* b = pB + lo;
  for (int i = 0; i < len; ++i) {
    const float w = weights[i * kEllStride];
    const float2 ww = make_float2(w, w);
    a01 = p_fma(ww, a[i], a01);
    a23 = p_fma(ww, b[i], a23);
  }
  acc[0] = a01.x; acc[1] = a01.y; acc[2] = a23.x; acc[3] = a23.y;
}
// The same for one frame pair.
AFS_HD float2 mel_dot_pair(const float2* pA, const float* weights, int lo, int len) {
  float2 a01 = make_float2(0.f, 0.f);
  const float2* a = pA + lo;
  for (int i = 0; i < len; ++i) {
    const float w = weights[i * kEllStride];
    a01 = p_fma(make_float2(w, w), a[i], a01);
  }
  return a01;
}

// Filter owned by (lane, virtual warp vw, pass) in the ELL table of pack_mel_ell (a warp of this engine walks the four
// warp-passes of the 64-thread layout one after the other), or -1.
AFS_HD int pair_mel_id(int lane, int vw, int pass, int n_mels) {
  const int t = 32 * vw + lane;
  if (pass == 0) return t < n_mels ? t : -1;
  return n_mels - 1 - t >= kGroup ? n_mels - 1 - t : -1;
}

}  // namespace logmel
}  // namespace afs
