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
#include "logmel_pair.cuh"

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

// ---- warp-per-frame-pair engine (logmel_pair.cuh): 32 'lanes' looped per phase; the shuffle of the split is a read of
// the partner lane's register array; power planes and mel batches exactly as logmel_pair.cu lays them out. ----
extern "C" int emul_logmel_pair(const float* wav, int64_t L, int hop, int center, const float* fb,
                                const float* window, int n_mels, const float* mean, const float* stdv,
                                float log_mult, float log_eps, float* out, float* power_out) {
  std::vector<int> band;
  std::vector<float> weights;
  pack_mel_ell(fb, n_mels, band, weights, kBanks64);
  const int pad = center ? kNfft / 2 : 0;
  const int T = center ? static_cast<int>(1 + L / hop) : static_cast<int>(1 + (L - kNfft) / hop);
  std::vector<float2> exch(kPairExch);
  std::vector<float2> planeA(kPStride, make_float2(0.f, 0.f)), planeB(kPStride, make_float2(0.f, 0.f));
  std::vector<float2> regs(32 * 32);
  auto sample = [&](int f, int n) -> float {
    if (f >= T) return 0.f;
    int64_t i = static_cast<int64_t>(f) * hop - pad + n;
    i = reflect_index(i, L);
    return wav[i];
  };
  const int chunk_frames = 8;
  for (int t0 = 0; t0 < T; t0 += chunk_frames) {
    const int nfr = T - t0 < chunk_frames ? T - t0 : chunk_frames;
    const int npairs = (nfr + 1) / 2;
    for (int j = 0; j < npairs; ++j) {
      const int fa = t0 + 2 * j;
      for (int lane = 0; lane < 32; ++lane) {
        float2 z[32];
        for (int n2 = 0; n2 < 32; ++n2) {
          const int n = lane + 32 * n2;
          const float wv = 0.5f * window[n];  // the kernel stages the window halved (logmel_pair.cuh)
          z[n2] = make_float2(sample(fa, n) * wv, sample(fa + 1, n) * wv);
        }
        const double a = 6.283185307179586476925286766559 * lane / kNfft;
        pair_pass1(lane, z, make_float2(static_cast<float>(cos(a)), static_cast<float>(-sin(a))), exch.data());
      }
      for (int lane = 0; lane < 32; ++lane) {
        float2 x[32];
        pair_pass2(lane, exch.data(), x);
        for (int i = 0; i < 32; ++i) regs[lane * 32 + i] = x[i];
      }
      float2* plane = (j & 1) ? planeB.data() : planeA.data();
      for (int lane = 0; lane < 32; ++lane) {
        const int partner = pair_partner(lane);
        float2 xs[32], xp[32];
        for (int i = 0; i < 32; ++i) { xs[i] = regs[lane * 32 + i]; xp[i] = regs[partner * 32 + i]; }
        for (int k2 = 0; k2 < 16; ++k2) {
          const float2 b = pair_split_src(partner, xp, k2);
          plane[lane + 32 * k2] = pair_power(xs[k2], b);
        }
        if (lane == 0) plane[512] = pair_power(xs[16], xs[16]);
      }
      if (power_out) {
        for (int h = 0; h < 2; ++h)
          if (fa + h < T)
            for (int k = 0; k < kBins; ++k) power_out[static_cast<size_t>(fa + h) * kBins + k] = h ? plane[k].y : plane[k].x;
      }
      const bool last = j + 1 >= npairs;
      if (!((j & 1) || last)) continue;
      const int c0 = (j & ~1) * 2;
      for (int lane = 0; lane < 32; ++lane)
        for (int vp = 0; vp < 4; ++vp) {
          const int m = pair_mel_id(lane, vp & 1, vp >> 1, n_mels);
          if (m < 0) continue;
          if (band[2 * kMaxMels + m] % kEllStride != lane) return -1;
          float acc[4];
          mel_dot_pairs(planeA.data(), planeB.data(), weights.data() + band[2 * kMaxMels + m], band[m], band[kMaxMels + m], acc);
          const float scale = log_mult * 0.30102999566398120f / stdv[m];
          const float shift = -mean[m] / stdv[m];
          for (int f = 0; f < 4; ++f)
            if (c0 + f < nfr) out[static_cast<size_t>(m) * T + t0 + c0 + f] = norm_db(acc[f], log_eps, scale, shift);
        }
    }
  }
  return T;
}

// Worst number of distinct 64-bit words one half-warp puts on one of the 16 eight-byte banks over the exchange's
// access patterns (1 = conflict-free), and -1 unless every (n1, k1) has its own slot inside the buffer.
extern "C" int emul_pair_bank_check() {
  std::vector<char> seen(kPairExch, 0);
  for (int n1 = 0; n1 < 32; ++n1)
    for (int k1 = 0; k1 < 32; ++k1) {
      const int s = n1 * kPairStride + k1;
      if (s >= kPairExch || seen[s]) return -1;
      seen[s] = 1;
    }
  int worst = 0;
  // group = 16 lanes on 16 eight-byte banks (64-bit access) or 8 lanes on 8 sixteen-byte banks (128-bit access)
  auto access = [&](int group, auto word_of_lane) {
    for (int g0 = 0; g0 < 32; g0 += group) {
      int words[16][16], cnt[16] = {0};
      for (int l = g0; l < g0 + group; ++l) {
        const int a = word_of_lane(l), b = a & (group - 1);
        bool dup = false;
        for (int i = 0; i < cnt[b]; ++i) dup = dup || words[b][i] == a;
        if (!dup) words[b][cnt[b]++] = a;
        if (cnt[b] > worst) worst = cnt[b];
      }
    }
  };
  for (int r = 0; r < 32; ++r) {
    access(16, [&](int lane) { return lane * kPairStride + r; });  // pass 1 writes k1 = r
    access(16, [&](int lane) { return r * kPairStride + lane; });  // pass 2 reads n1 = r
  }
  for (int q = 0; q < 8; ++q) access(8, [&](int lane) { return (lane * 36 + 4 * q) / 4; });  // window reads (kPWinStride = 36)
  return worst;
}

// Shared-memory wavefronts per frame PAIR of the pair engine's mel power reads (64-bit word per bin, served per
// half-warp), replayed from the table packed for that bank model; *iterations = read instructions per pair.
extern "C" int emul_mel_read_wavefronts64(const float* fb, int n_mels, int shifted, int* iterations) {
  std::vector<int> band, b0;
  std::vector<float> weights, w0;
  pack_mel_ell(fb, n_mels, band, weights, kBanks64);
  pack_mel_bands(fb, n_mels, b0, w0);
  int total = 0;
  *iterations = 0;
  for (int vp = 0; vp < 4; ++vp) {
    int start[32], n[32], longest = 0;
    for (int lane = 0; lane < 32; ++lane) {
      const int m = pair_mel_id(lane, vp & 1, vp >> 1, n_mels);
      start[lane] = m < 0 ? 0 : (shifted ? band[m] : b0[m]);
      n[lane] = m < 0 ? 0 : (shifted ? band[kMaxMels + m] : b0[kMaxMels + m]);
      if (start[lane] < 0 || start[lane] + n[lane] > kBins) return -1;
      if (n[lane] > longest) longest = n[lane];
    }
    *iterations += longest;
    total += mel_read_wavefronts(start, n, kBanks64);
  }
  return total;
}
