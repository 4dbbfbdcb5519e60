// Branch-free streaming top-k over accumulator columns, shared by the tensor-core DN4 kernels.
#pragma once
#include <stdint.h>

namespace afs {
namespace topk {

// ---- branch-free streaming top-NK -------------------------------------------------------------------
// A relation value and its column inside the 128-column tile travel as ONE sortable 32-bit key: the fp32
// bit pattern mapped to an order-preserving unsigned integer, low 7 bits replaced by (127 - column).  A
// sorted insert is then 2*NK-1 integer min/max instructions with no branch (the per-lane `if (x > worst)`
// of a scalar insertion diverges on almost every column: 32 lanes each own a different row).  The 7 bits
// cost 2^-16 relative resolution on the value, below the TF32 rounding of the operands; ties resolve
// to the lower column, as torch.topk / the fp32 path.
__device__ __forceinline__ uint32_t topk_key(uint32_t bits, int col_in_tile) {
  const uint32_t mono = bits ^ (static_cast<uint32_t>(static_cast<int32_t>(bits) >> 31) | 0x80000000u);
  return (mono & ~127u) | static_cast<uint32_t>(127 - col_in_tile);
}
__device__ __forceinline__ float topk_key_value(uint32_t key) {
  const uint32_t mono = key & ~127u;
  const uint32_t bits = (mono & 0x80000000u) ? (mono ^ 0x80000000u) : ~mono;
  return __uint_as_float(bits);
}
template <int NK>
__device__ __forceinline__ void topk_push(uint32_t (&t)[NK], uint32_t key) {
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const uint32_t hi = max(t[k], key);
    key = min(t[k], key);
    t[k] = hi;
  }
}
// scalar sorted insert of (value, global column) into the running result (once per selected key per tile)
template <int NK>
__device__ __forceinline__ void topk_merge(float (&tv)[NK], int (&ti)[NK], float x, int col) {
  if (x > tv[NK - 1] || (x == tv[NK - 1] && col < ti[NK - 1])) {
    tv[NK - 1] = x;
    ti[NK - 1] = col;
#pragma unroll
    for (int k = NK - 1; k > 0; --k) {
      if (tv[k] > tv[k - 1] || (tv[k] == tv[k - 1] && ti[k] < ti[k - 1])) {
        const float fv = tv[k]; tv[k] = tv[k - 1]; tv[k - 1] = fv;
        const int iv = ti[k]; ti[k] = ti[k - 1]; ti[k - 1] = iv;
      }
    }
  }
}

}  // namespace topk
}  // namespace afs
