// First Conv64F block, inference: Conv2d(1->C, 3x3, pad 1) + BatchNorm2d(eval) + ReLU/LeakyReLU +
// MaxPool2d(3, 3), fused, for sm_100a.
//
// Reference: libfewshot_core/model/backbone/conv_four.py:61-66,101-103 (layer1 of Conv64F).  In the
// reference's eager path this block is five kernels (cuDNN conv with a layout round trip, bias add,
// batch-norm, ReLU, max-pool) that each stream a [N,64,128,157] fp32 activation (5.1 MB per clip)
// through HBM: 88 % of a whole 5w5s15q evaluation step on B200 (profiles/r01_bench_launches.csv).
// Here the activation never exists: a thread owns one POOLED output pixel, keeps its 5x5 input patch
// in registers, evaluates the 3x3 conv at the 9 positions of the pooling window for every channel
// with the (BatchNorm-folded) weights as constant-bank operands, takes the max, adds the folded
// shift, applies the activation (max and a monotone activation commute; the BatchNorm scale is
// folded into the weights BEFORE the max because it may be negative), and the CTA writes its
// [128 pixels x C] tile as one contiguous, coalesced NHWC span -- the channels-last layout the
// tensor-core convolution of the next block consumes without a transpose.
//
// HBM per clip: 4*H*W read + 4*C*(H/3)*(W/3) write (80 KB + 559 KB at 128x157, C=64).
#include "common.cuh"

namespace afs {
namespace {

constexpr int kC = 64;
constexpr int kPix = 128;           // pooled pixels per CTA = threads per CTA
constexpr int kTileStride = kC + 1;  // conflict-free transposition through shared memory

struct Conv1Params {
  float w[kC * 9];  // folded weights, [c][ky][kx]
  float shift[kC];  // folded (bias - mean) * scale + beta
};

__global__ void __launch_bounds__(kPix)
conv1_bn_act_pool3_kernel(const float* __restrict__ x, int64_t total_pix, int H, int Wd, int PH, int PW,
                          float slope, float* __restrict__ out, const __grid_constant__ Conv1Params prm) {
  __shared__ float s_tile[kPix * kTileStride];
  const int tid = threadIdx.x;
  const int64_t pix0 = static_cast<int64_t>(blockIdx.x) * kPix;
  const int64_t pix = pix0 + tid;
  const bool live = pix < total_pix;

  float p[5][5];
  {
    const int64_t pp = live ? pix : total_pix - 1;
    const int per = PH * PW;
    const int64_t n = pp / per;
    const int r = static_cast<int>(pp - n * per);
    const int ph = r / PW;
    const int pw = r - ph * PW;
    const float* img = x + n * static_cast<int64_t>(H) * Wd;
    const int y0 = 3 * ph - 1, x0 = 3 * pw - 1;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int yy = y0 + i;
      const bool yin = (yy >= 0) && (yy < H);
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int xx = x0 + j;
        p[i][j] = (yin && xx >= 0 && xx < Wd) ? __ldg(img + static_cast<int64_t>(yy) * Wd + xx) : 0.f;
      }
    }
  }

#pragma unroll 2
  for (int c = 0; c < kC; ++c) {
    const float* w = prm.w + c * 9;
    float best = -INFINITY;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        // cross-correlation in the reference's summation order (ky outer, kx inner)
        float acc = p[dy][dx] * w[0];
        acc = fmaf(p[dy][dx + 1], w[1], acc);
        acc = fmaf(p[dy][dx + 2], w[2], acc);
        acc = fmaf(p[dy + 1][dx], w[3], acc);
        acc = fmaf(p[dy + 1][dx + 1], w[4], acc);
        acc = fmaf(p[dy + 1][dx + 2], w[5], acc);
        acc = fmaf(p[dy + 2][dx], w[6], acc);
        acc = fmaf(p[dy + 2][dx + 1], w[7], acc);
        acc = fmaf(p[dy + 2][dx + 2], w[8], acc);
        best = fmaxf(best, acc);
      }
    }
    float v = best + prm.shift[c];
    v = v > 0.f ? v : v * slope;
    s_tile[tid * kTileStride + c] = v;
  }
  __syncthreads();

  // [kPix][C] tile -> contiguous NHWC span, 128 B per warp store
  const int64_t remain = total_pix - pix0;
  const int npix = remain < kPix ? static_cast<int>(remain) : kPix;
  float* dst = out + pix0 * kC;
  for (int i = tid; i < npix * kC; i += kPix) {
    const int pxl = i >> 6;
    const int c = i & (kC - 1);
    dst[i] = s_tile[pxl * kTileStride + c];
  }
}

}  // namespace
}  // namespace afs

extern "C" int afs_conv1_bn_act_pool3_fwd(const float* x, int32_t N, int32_t H, int32_t Wd,
                                          const float* w_folded_host, const float* shift_host,
                                          int32_t C, float negative_slope, float* out,
                                          afs_stream_t stream_) {
  using namespace afs;
  if (x == nullptr || w_folded_host == nullptr || shift_host == nullptr || out == nullptr || N < 0 ||
      H < 3 || Wd < 3 || negative_slope < 0.f)
    return AFS_ERR_INVALID_ARG;
  if (C != kC) return AFS_ERR_UNSUPPORTED;
  if (N == 0) return AFS_OK;
  const int PH = H / 3, PW = Wd / 3;
  const int64_t total = static_cast<int64_t>(N) * PH * PW;
  const int64_t blocks = (total + kPix - 1) / kPix;
  if (blocks > 0x7fffffffLL) return AFS_ERR_UNSUPPORTED;
  Conv1Params prm;
  for (int i = 0; i < kC * 9; ++i) prm.w[i] = w_folded_host[i];
  for (int i = 0; i < kC; ++i) prm.shift[i] = shift_host[i];
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  conv1_bn_act_pool3_kernel<<<static_cast<unsigned>(blocks), kPix, 0, stream>>>(x, total, H, Wd, PH, PW,
                                                                              negative_slope, out, prm);
  AFS_LAUNCH_CHECK();
  return AFS_OK;
}
