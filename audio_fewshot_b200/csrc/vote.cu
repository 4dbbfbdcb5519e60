// Window argmax -> per-query majority vote -> accuracy, and the DeepBDC energy
// score, on device.
//
// Replaces majority_vote + vote_catagorical_acc (reference
// libfewshot_core/utils/utils.py:432-446), which loop in Python over every
// query, call torch.mode on a slice and write the result into a CPU tensor
// (one device->host sync per query), and average_logits + -logsumexp
// (utils.py:449-471, libfewshot_core/model/metric/deepbdc.py:318-319).
// Tie rule of the vote (several labels equally frequent in a query's windows):
//   AFS_VOTE_TIE_SMALLEST    torch.mode on a CPU tensor: the smallest tied label;
//   AFS_VOTE_TIE_TORCH_CUDA  torch.mode on a CUDA tensor, which is what the reference's
//     set_forward executes (utils.py:443 on a 'cuda' slice, proto_net.py:116).  Its
//     fused small-slice kernel sorts the slice, gives sorted positions (2t, 2t+1) to
//     thread t, and max-reduces (count, position) with shuffle-down trees in which the
//     lower lane wins ties: among the tied labels the winner is the one whose run END
//     sits in the lane with the smallest bit-reversed (warp, lane) id.  Measured on
//     B200 with torch 2.11 (tools/probe_torch_mode.py; 0 mismatches in 4 000 slices).
// The argmax is taken on the raw logits (softmax is monotone; the reference
// argmaxes softmax(logits), utils.py:437), lowest index on ties.
#include "common.cuh"

namespace afs {
namespace {

constexpr int kMaxWay = 64;

__global__ void __launch_bounds__(256)
vote_kernel(const float* __restrict__ logits, int W, const int32_t* __restrict__ q_start, int nq,
            const int32_t* __restrict__ q_target, int tie_rule, int32_t* __restrict__ q_pred,
            int32_t* __restrict__ stats, float* __restrict__ acc_pct) {
  __shared__ int s_correct;
  __shared__ bool s_last;
  if (threadIdx.x == 0) s_correct = 0;
  __syncthreads();

  int local_correct = 0;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
    const int r0 = q_start[q];
    const int r1 = q_start[q + 1];
    int best_label = -1;
    if (r1 - r0 == 1) {  // the common case (repeats == 1): mode == the argmax
      const float* row = logits + static_cast<int64_t>(r0) * W;
      float best = row[0];
      best_label = 0;
      for (int w = 1; w < W; ++w) {
        const float v = row[w];
        if (v > best) { best = v; best_label = w; }
      }
    } else if (r1 > r0) {
      int cnt32[kMaxWay];
      for (int w = 0; w < W; ++w) cnt32[w] = 0;
      for (int r = r0; r < r1; ++r) {
        const float* row = logits + static_cast<int64_t>(r) * W;
        float best = row[0];
        int arg = 0;
        for (int w = 1; w < W; ++w) {
          const float v = row[w];
          if (v > best) { best = v; arg = w; }
        }
        cnt32[arg] += 1;
      }
      int best_cnt = 0;
      for (int w = 0; w < W; ++w) {
        if (cnt32[w] > best_cnt) { best_cnt = cnt32[w]; best_label = w; }
      }
      if (tie_rule == AFS_VOTE_TIE_TORCH_CUDA && best_cnt > 1) {
        unsigned best_key = 0xffffffffu;
        int cum = 0;
        for (int w = 0; w < W; ++w) {
          cum += cnt32[w];
          if (cnt32[w] == best_cnt) {
            const unsigned lane = static_cast<unsigned>(cum - 1) >> 1;  // thread holding the run end
            const unsigned key = ((__brev((lane >> 5) & 31u) >> 27) << 5) | (__brev(lane & 31u) >> 27);
            if (key < best_key) { best_key = key; best_label = w; }
          }
        }
      }
    }
    q_pred[q] = best_label;
    local_correct += (best_label == q_target[q]) ? 1 : 0;
  }
  if (local_correct) atomicAdd(&s_correct, local_correct);
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_correct) atomicAdd(&stats[0], s_correct);
    __threadfence();
    const int ticket = atomicAdd(&stats[2], 1);
    s_last = (ticket == static_cast<int>(gridDim.x) - 1);
    if (s_last) {
      __threadfence();
      const int correct = atomicAdd(&stats[0], 0);
      stats[1] = nq;
      // (predictions == targets).sum().float() / n * 100.0  (utils.py:433)
      *acc_pct = static_cast<float>(correct) / static_cast<float>(nq) * 100.0f;
    }
  }
}

__global__ void __launch_bounds__(256)
energy_kernel(const float* __restrict__ logits, int W, const int32_t* __restrict__ q_start, int nq,
              float* __restrict__ energy) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const int r0 = q_start[q];
  const int r1 = q_start[q + 1];
  float avg[kMaxWay];
  for (int w = 0; w < W; ++w) avg[w] = 0.f;
  for (int r = r0; r < r1; ++r) {
    const float* row = logits + static_cast<int64_t>(r) * W;
    for (int w = 0; w < W; ++w) avg[w] += row[w];
  }
  const float n = static_cast<float>(r1 - r0);
  float m = -INFINITY;
  for (int w = 0; w < W; ++w) {
    avg[w] = (r1 > r0) ? avg[w] / n : 0.f;  // average_logits: zeros for an empty group
    m = fmaxf(m, avg[w]);
  }
  float s = 0.f;
  for (int w = 0; w < W; ++w) s += expf(avg[w] - m);
  energy[q] = -(m + logf(s));
}

}  // namespace
}  // namespace afs

extern "C" int afs_vote_acc(const float* logits, int32_t W, const int32_t* q_start, int32_t nq,
                            const int32_t* q_target, int32_t tie_rule, int32_t* q_pred,
                            int32_t* stats, float* acc_pct, afs_stream_t stream_) {
  using namespace afs;
  if (logits == nullptr || q_start == nullptr || q_target == nullptr || q_pred == nullptr ||
      stats == nullptr || acc_pct == nullptr || W < 1 || W > kMaxWay || nq < 1 ||
      (tie_rule != AFS_VOTE_TIE_SMALLEST && tie_rule != AFS_VOTE_TIE_TORCH_CUDA))
    return AFS_ERR_INVALID_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  AFS_CUDA_TRY(cudaMemsetAsync(stats, 0, 4 * sizeof(int32_t), stream));
  int blocks = (nq + 255) / 256;
  if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
  vote_kernel<<<blocks, 256, 0, stream>>>(logits, W, q_start, nq, q_target, tie_rule, q_pred, stats,
                                          acc_pct);
  AFS_LAUNCH_CHECK();
  return AFS_OK;
}

extern "C" int afs_energy_score(const float* logits, int32_t W, const int32_t* q_start, int32_t nq,
                                float* energy, afs_stream_t stream_) {
  using namespace afs;
  if (logits == nullptr || q_start == nullptr || energy == nullptr || W < 1 || W > kMaxWay || nq < 1)
    return AFS_ERR_INVALID_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  energy_kernel<<<(nq + 255) / 256, 256, 0, stream>>>(logits, W, q_start, nq, energy);
  AFS_LAUNCH_CHECK();
  return AFS_OK;
}
