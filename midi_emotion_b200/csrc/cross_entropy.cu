// Fused cross-entropy over the output head's logits (train.py:124,288-290: CrossEntropyLoss(ignore_index=pad),
// mean over the non-pad targets) together with its gradient and the top-1 / top-5 hit counts of
// utils.accuracy (utils.py:15-80, train.py:256), in one pass over the [M, V] logits:
//   loss_i = logsumexp(x_i) - x_i[t_i]           grad_i = (softmax(x_i) - onehot(t_i)) / count
// One warp per row, the row lives in registers (V <= 4096), fp32 math on the stored (bf16 or fp32) logits.
// The reference materialises log_softmax [M, V] fp32, its gradient and the one-hot scatter as separate
// kernels; here the logits are read once and the gradient is written once (in place if the caller wants).
#include "common.cuh"
#include "../../include/midi_emotion_b200.h"

namespace me {

constexpr int CE_WARPS = 8;
constexpr int CE_MAXV = 4096;

// rows that count: the same predicate as ce_kernel's (an out-of-range target -- torch raises there -- is skipped by
// both, so it can never bias the mean)
__global__ void ce_count_kernel(const int64_t* __restrict__ targets, int M, int V, int64_t ignore_index,
                                float* __restrict__ stats) {
  __shared__ int part[32];
  int c = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
    const int64_t t = targets[i];
    c += (t != ignore_index && t >= 0 && t < V);
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) s += part[w];
    atomicAdd(&stats[1], static_cast<float>(s));
  }
}

template <typename T, int NV>  // NV = values per lane (V <= 32 * NV)
__global__ void __launch_bounds__(CE_WARPS * 32)
ce_kernel(const T* __restrict__ logits, int M, int V, int ld, const int64_t* __restrict__ targets, int64_t ignore_index,
          T* __restrict__ grad, int ld_grad, float* __restrict__ stats) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float count = stats[1];
  const float inv_count = count > 0.f ? 1.f / count : 0.f;
  float loss_acc = 0.f, top1 = 0.f, top5 = 0.f;
  for (int row = blockIdx.x * CE_WARPS + warp; row < M; row += gridDim.x * CE_WARPS) {
    const int64_t t = targets[row];
    T* grow = grad ? grad + static_cast<int64_t>(row) * ld_grad : nullptr;
    if (t == ignore_index || t < 0 || t >= V) {  // ignored rows contribute nothing and get a zero gradient
      if (grow)
        for (int j = lane; j < ld_grad; j += 32) grow[j] = from_f32<T>(0.f);
      continue;
    }
    const T* xrow = logits + static_cast<int64_t>(row) * ld;
    float x[NV];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int j = lane + 32 * k;
      x[k] = j < V ? to_f32<T>(xrow[j]) : -INFINITY;
      mx = fmaxf(mx, x[k]);
    }
    mx = warp_max(mx);
    float se = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) se += (lane + 32 * k < V) ? __expf(x[k] - mx) : 0.f;
    se = warp_sum(se);
    const float lse = mx + __logf(se);
    const int tt = static_cast<int>(t);
    float xsel = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) xsel = (k == (tt >> 5)) ? x[k] : xsel;   // (no dynamic register indexing)
    const float xt = __shfl_sync(0xffffffffu, xsel, tt & 31);
    // rank of the target among the logits (number of strictly larger ones): top-k hit iff rank < k
    int bigger = 0;
#pragma unroll
    for (int k = 0; k < NV; ++k) bigger += (lane + 32 * k < V) && (x[k] > xt);
    for (int o = 16; o > 0; o >>= 1) bigger += __shfl_xor_sync(0xffffffffu, bigger, o);
    if (lane == 0) {
      loss_acc += lse - xt;
      top1 += bigger < 1 ? 1.f : 0.f;
      top5 += bigger < 5 ? 1.f : 0.f;
    }
    if (grow) {
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int j = lane + 32 * k;
        if (j < V) grow[j] = from_f32<T>((__expf(x[k] - lse) - (j == tt ? 1.f : 0.f)) * inv_count);
        else if (j < ld_grad) grow[j] = from_f32<T>(0.f);
      }
      for (int j = 32 * NV + lane; j < ld_grad; j += 32) grow[j] = from_f32<T>(0.f);
    }
  }
  __shared__ float red[3][CE_WARPS];
  if (lane == 0) { red[0][warp] = loss_acc; red[1][warp] = top1; red[2][warp] = top5; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int w = 0; w < CE_WARPS; ++w) s += red[threadIdx.x][w];
    if (s != 0.f) atomicAdd(&stats[threadIdx.x == 0 ? 0 : threadIdx.x + 1], s);
  }
}

template <typename T>
static int launch_ce(const void* logits, int M, int V, int ld, const int64_t* targets, int64_t ignore_index, void* grad,
                     int ld_grad, float* stats, cudaStream_t st) {
  const int blocks = min((M + CE_WARPS - 1) / CE_WARPS, sm_count() * 8);
  const T* x = static_cast<const T*>(logits);
  T* g = static_cast<T*>(grad);
#define ME_CE(NV) ce_kernel<T, NV><<<blocks, CE_WARPS * 32, 0, st>>>(x, M, V, ld, targets, ignore_index, g, ld_grad, stats)
  const int nv = (V + 31) / 32;
  if (nv <= 8) ME_CE(8);
  else if (nv <= 16) ME_CE(16);
  else if (nv <= 32) ME_CE(32);
  else if (nv <= 64) ME_CE(64);
  else ME_CE(128);
#undef ME_CE
  ME_LAUNCH_CHECK();
  return 0;
}

}  // namespace me

using namespace me;

extern "C" int me_cross_entropy(const void* logits, int dtype, int M, int V, int ld, const int64_t* targets,
                                int64_t ignore_index, void* grad_logits, int ld_grad, float* stats, void* stream) {
  ME_CHECK(logits && targets && stats, "me_cross_entropy: NULL pointer");
  ME_CHECK(M > 0 && V > 0 && V <= CE_MAXV && ld >= V, "me_cross_entropy: bad sizes (M=%d, V=%d <= %d, ld=%d)", M, V, CE_MAXV, ld);
  ME_CHECK(dtype == ME_F32 || dtype == ME_BF16, "me_cross_entropy: bad dtype %d", dtype);
  ME_CHECK(grad_logits == nullptr || ld_grad >= V, "me_cross_entropy: ld_grad %d < V %d", ld_grad, V);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ME_CUDA(cudaMemsetAsync(stats, 0, 4 * sizeof(float), st));
  ce_count_kernel<<<min((M + 255) / 256, 64), 256, 0, st>>>(targets, M, V, ignore_index, stats);
  ME_LAUNCH_CHECK();
  if (dtype == ME_BF16) return launch_ce<bf16>(logits, M, V, ld, targets, ignore_index, grad_logits, ld_grad, stats, st);
  return launch_ce<float>(logits, M, V, ld, targets, ignore_index, grad_logits, ld_grad, stats, st);
}
