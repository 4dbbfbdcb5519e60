// CPU emulation of the fused log-mel kernel's per-thread phases (host logic test).
//
// Runs the SAME __host__ __device__ phase functions the kernel runs
// (audio_fewshot_b200/csrc/logmel_core.cuh, logmel_fft.cuh) with the 64 threads of a frame group
// looped sequentially per phase (the loop boundary plays the role of bar.sync).
// Built by tests/test_logmel_emul.py with `nvcc -x cu` (host code only is used).
#include <math.h>
#include <stdint.h>

#include <vector>

#include "logmel_core.cuh"
#include "logmel_fft.cuh"

using namespace afs::logmel;

static int emul_logmel_impl(const float* wav, int64_t L, int hop, int center, const float* fb,
                            const float* window, int n_mels, const float* mean, const float* stdv,
                            float log_mult, float log_eps, float* out /*[n_mels, T]*/,
                            float* power_out /*[T, 513] nullable*/) {
  std::vector<int> band;
  std::vector<float> weights;
  pack_mel_ell(fb, n_mels, band, weights);  // the table the kernel reads, start shifts included
  std::vector<float2> tw(kNfft);
  for (int k = 0; k < kNfft; ++k) {
    const double a = 6.283185307179586476925286766559 * k / kNfft;
    tw[k] = make_float2(static_cast<float>(cos(a)), static_cast<float>(-sin(a)));
  }
  const int pad = center ? kNfft / 2 : 0;
  const int T = center ? static_cast<int>(1 + L / hop) : static_cast<int>(1 + (L - kNfft) / hop);
  std::vector<float> bufA(kBufA), bufB(kBufB), bufP(kMelBatch * kPStride, 0.f);
  std::vector<ThreadTw> tws(kGroup);
  for (int t = 0; t < kGroup; ++t) load_thread_tw(tws[t], t, tw.data());
  for (int f = 0; f < T; ++f) {
    const int64_t s0 = static_cast<int64_t>(f) * hop - pad;
    const bool interior = s0 >= 0 && s0 + kNfft <= L;
    for (int t = 0; t < kGroup; ++t) {
      float2 zp[8];
      for (int r = 0; r < 8; ++r) {
        const int n = t + 64 * r;
        int64_t i0 = s0 + 2 * n, i1 = i0 + 1;
        if (!interior) { i0 = reflect_index(i0, L); i1 = reflect_index(i1, L); }
        zp[r] = make_float2(wav[i0] * window[2 * n], wav[i1] * window[2 * n + 1]);
      }
      phase_a_p(t, zp, tws[t], reinterpret_cast<float2*>(bufA.data()));
    }
    const int slot = f % kMelBatch;
    float* power = bufP.data() + slot * kPStride;
    float2* a2 = reinterpret_cast<float2*>(bufA.data());
    float2* b2 = reinterpret_cast<float2*>(bufB.data());
    for (int t = 0; t < kGroup; ++t) phase_b_p(t, tws[t], a2, b2);
    for (int t = 0; t < kGroup; ++t) phase_c_p(t, b2, a2);
    for (int t = 0; t < kGroup; ++t) phase_d_p(t, tws[t], a2, power);
    if (power_out) for (int k = 0; k < kBins; ++k) power_out[static_cast<size_t>(f) * kBins + k] = power[k];
    if (slot != kMelBatch - 1 && f + 1 != T) continue;
    for (int t = 0; t < kGroup; ++t) {  // batched mel projection, as the kernel's epilogue
      int mel_id[2];
      mel_id[0] = (t < n_mels) ? t : -1;
      mel_id[1] = (n_mels - 1 - t >= kGroup) ? n_mels - 1 - t : -1;
      for (int i = 0; i < 2; ++i) {
        const int m = mel_id[i];
        if (m < 0) continue;
        float acc[kMelBatch];
        mel_dot_batch_p(bufP.data(), weights.data() + band[2 * kMaxMels + m], kEllStride, band[m], band[kMaxMels + m], acc);
        const float scale = log_mult * 0.30102999566398120f / stdv[m];
        const float shift = -mean[m] / stdv[m];
        for (int b = 0; b <= slot; ++b)
          out[static_cast<size_t>(m) * T + (f - slot) + b] = norm_db(acc[b], log_eps, scale, shift);
      }
    }
  }
  return T;
}

extern "C" int emul_logmel(const float* wav, int64_t L, int hop, int center, const float* fb,
                           const float* window, int n_mels, const float* mean, const float* stdv,
                           float log_mult, float log_eps, float* out, float* power_out) {
  return emul_logmel_impl(wav, L, hop, center, fb, window, n_mels, mean, stdv, log_mult, log_eps, out, power_out);
}

// Exchange layouts: worst number of distinct 64-bit words that one half-warp (16 consecutive threads) puts on one of
// the 16 eight-byte banks, over every shared-memory access pattern of the packed phases (1 = conflict-free), and
// whether the exchange-2 slot function is a bijection onto [0, 512).  Returns the worst count, or -1.
extern "C" int emul_packed_bank_check() {
  bool seen[kHalf] = {false};
  for (int q = 0; q < 8; ++q)
    for (int j = 0; j < 8; ++j)
      for (int p = 0; p < 8; ++p) {
        const int s = e2p_slot(q, j, p);
        if (s < 0 || s >= kHalf || seen[s]) return -1;
        seen[s] = true;
      }
  int worst = 0;
  auto half_warp = [&](auto addr_of_thread) {
    for (int h = 0; h < 4; ++h) {
      int words[16][16], cnt[16] = {0};
      for (int l = 0; l < 16; ++l) {
        const int a = addr_of_thread(16 * h + l), b = a & 15;
        bool dup = false;
        for (int i = 0; i < cnt[b]; ++i) dup = dup || words[b][i] == a;
        if (!dup) words[b][cnt[b]++] = a;
        if (cnt[b] > worst) worst = cnt[b];
      }
    }
  };
  for (int r = 0; r < 8; ++r) {
    half_warp([&](int t) { return e1p_slot(r, t); });                        // phase A writes (q = r)
    half_warp([&](int t) { return e1p_slot(t & 7, (t >> 3) + 8 * r); });     // phase B reads (j1 = r)
    half_warp([&](int t) { return e2p_slot(t & 7, t >> 3, r); });            // phase B writes (p0 = r)
    half_warp([&](int t) { return e2p_slot(t & 7, r, t >> 3); });            // phase C reads (j0 = r)
    half_warp([&](int t) { return t + 64 * r; });                            // phase C writes (p1 = r)
  }
  for (int m = 0; m < 4; ++m) {
    half_warp([&](int t) { return t + 64 * m; });                            // phase D reads Z[k]
    half_warp([&](int t) { return (kHalf - (t + 64 * m)) & (kHalf - 1); });  // phase D reads Z[512 - k]
  }
  return worst;
}

// The kernel's ELL weight table (pack_mel_ell) expanded back to a dense [513, n_mels] matrix, plus, per warp-pass,
// the number of table rows: lets the test check that the table reproduces the filterbank exactly, that every
// filter is owned by exactly one (thread, pass), and how much padding the layout costs.  Returns the table size.
extern "C" int emul_mel_ell_dense(const float* fb, int n_mels, float* dense /*[513, n_mels]*/, int* owner_count /*[n_mels]*/) {
  std::vector<int> band;
  std::vector<float> weights;
  const int n = pack_mel_ell(fb, n_mels, band, weights);
  for (int i = 0; i < kBins * n_mels; ++i) dense[i] = 0.f;
  for (int m = 0; m < n_mels; ++m) owner_count[m] = 0;
  for (int t = 0; t < kGroup; ++t) {
    const int ids[2] = {t < n_mels ? t : -1, n_mels - 1 - t >= kGroup ? n_mels - 1 - t : -1};
    for (int i = 0; i < 2; ++i) {
      const int m = ids[i];
      if (m < 0) continue;
      owner_count[m] += 1;
      const int lo = band[m], len = band[kMaxMels + m], off = band[2 * kMaxMels + m];
      if (off % kEllStride != (t & 31)) return -1;  // weight i of lane l must sit at column l of its table row
      for (int j = 0; j < len; ++j) dense[(lo + j) * n_mels + m] = weights[off + j * kEllStride];
    }
  }
  return n;
}

// Shared-memory wavefronts per frame of the mel projection's power reads, replayed from the packed table exactly as
// the kernel's warps issue them (bank = word index mod 32; the four frames of a batch are separate instructions
// with the same pattern).  *iterations = number of read instructions per frame = the conflict-free count.
// shifted = 0 replays the un-shifted layout (every filter starts at its first non-zero bin) for comparison.
extern "C" int emul_mel_read_wavefronts(const float* fb, int n_mels, int shifted, int* iterations) {
  std::vector<int> band, b0;
  std::vector<float> weights, w0;
  pack_mel_ell(fb, n_mels, band, weights);
  pack_mel_bands(fb, n_mels, b0, w0);
  int total = 0;
  *iterations = 0;
  for (int pass = 0; pass < 2; ++pass) {
    for (int warp = 0; warp < kGroup / 32; ++warp) {
      int start[32], n[32], longest = 0;
      for (int lane = 0; lane < 32; ++lane) {
        const int t = warp * 32 + lane;
        const int m = pass == 0 ? (t < n_mels ? t : -1) : (n_mels - 1 - t >= kGroup ? n_mels - 1 - t : -1);
        start[lane] = m < 0 ? 0 : (shifted ? band[m] : b0[m]);
        n[lane] = m < 0 ? 0 : (shifted ? band[kMaxMels + m] : b0[kMaxMels + m]);
        if (start[lane] < 0 || start[lane] + n[lane] > kBins) return -1;  // every read must hit a written bin
        if (n[lane] > longest) longest = n[lane];
      }
      *iterations += longest;
      total += mel_read_wavefronts(start, n);
    }
  }
  return total;
}
