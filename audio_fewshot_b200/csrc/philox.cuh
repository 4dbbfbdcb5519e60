// Philox4x32-10 counter-based RNG (Salmon et al., SC'11) and the waveform
// augmentation draws built on it.  The integer stream is bit-exact with
// oracle/philox.py; both are pinned to the Random123 known-answer vectors in
// tests/test_philox.py.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define AFS_PHILOX_HD __host__ __device__ __forceinline__
#else
#define AFS_PHILOX_HD inline
#endif

namespace afs {

struct u32x4 {
  uint32_t x, y, z, w;
};

AFS_PHILOX_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32);
#endif
}

AFS_PHILOX_HD u32x4 philox4x32_10(u32x4 c, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mulhi32(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = mulhi32(M1, c.z), lo1 = M1 * c.z;
    u32x4 n;
    n.x = hi1 ^ c.y ^ k0;
    n.y = lo1;
    n.z = hi0 ^ c.w ^ k1;
    n.w = lo0;
    c = n;
    k0 += W0;
    k1 += W1;
  }
  return c;
}

// uint32 -> float in (0, 1): ((r >> 9) + 0.5) * 2^-23; every step is exact in fp32
// (2^23 - 0.5 needs 24 mantissa bits), so the value is bit-identical on host and device.
AFS_PHILOX_HD float u01(uint32_t r) {
  return (static_cast<float>(r >> 9) + 0.5f) * 1.1920928955078125e-07f;
}

// Streams (counter word y): 0 = per-clip parameters, 1 = additive noise.
constexpr uint32_t kStreamParams = 0u;
constexpr uint32_t kStreamNoise = 1u;

}  // namespace afs
