// Output head of the regression side model (models/music_regression.py:65-68,89): the first position of
// every sequence through Linear(d, n_out) + tanh, and its backward.  B x d x n_out multiply-adds: one
// thread block per sequence, no tensor cores needed.
#include "common.cuh"
#include "../../include/midi_emotion_b200.h"

namespace me {

constexpr int RH_THREADS = 128;
constexpr int RH_MAX_OUT = 8;

__device__ __forceinline__ float rh_round(float v, int dtype) {
  return dtype == ME_BF16 ? __bfloat162float(__float2bfloat16_rn(v)) : v;
}

// x: [B, Ls, d] (fp32, or bf16 when dtype == ME_BF16); only row (b, 0) is read
template <typename T>
__global__ void __launch_bounds__(RH_THREADS)
pooled_head_fwd_kernel(const T* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias, int Ls,
                       int d, int n_out, int dtype, float* __restrict__ out) {
  __shared__ float red[RH_MAX_OUT][RH_THREADS / 32];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const T* row = x + static_cast<int64_t>(b) * Ls * d;
  float acc[RH_MAX_OUT];
#pragma unroll
  for (int o = 0; o < RH_MAX_OUT; ++o) acc[o] = 0.f;
  for (int c = threadIdx.x; c < d; c += RH_THREADS) {
    const float xv = to_f32<T>(row[c]);
#pragma unroll
    for (int o = 0; o < RH_MAX_OUT; ++o)
      if (o < n_out) acc[o] = fmaf(xv, rh_round(W[o * d + c], dtype), acc[o]);  // autocast casts the weight to bf16
  }
#pragma unroll
  for (int o = 0; o < RH_MAX_OUT; ++o) {
    const float s = warp_sum(acc[o]);
    if (lane == 0) red[o][warp] = s;
  }
  __syncthreads();
  if (threadIdx.x < n_out) {
    const int o = threadIdx.x;
    float s = 0.f;
    for (int w = 0; w < RH_THREADS / 32; ++w) s += red[o][w];
    const float pre = rh_round(s + bias[o], dtype);   // Linear output (bf16 under autocast)
    out[b * n_out + o] = rh_round(tanhf(pre), dtype);
  }
}

// dpre = g * (1 - out^2);  db[o] = sum_b dpre;  dW[o, c] = sum_b dpre[b, o] x[b, 0, c];
// d_x[b, 0, c] = sum_o dpre[b, o] W[o, c]  (the other rows of d_x are the caller's zeros)
template <typename T>
__global__ void __launch_bounds__(RH_THREADS)
pooled_head_bwd_kernel(const float* __restrict__ g_out, const float* __restrict__ out, const T* __restrict__ x,
                       const float* __restrict__ W, int B, int Ls, int d, int n_out, int dtype,
                       float* __restrict__ dW, float* __restrict__ db, float* __restrict__ d_x) {
  const int c = blockIdx.x * RH_THREADS + threadIdx.x;
  float w[RH_MAX_OUT], gw[RH_MAX_OUT], gb[RH_MAX_OUT];
#pragma unroll
  for (int o = 0; o < RH_MAX_OUT; ++o) {
    w[o] = (o < n_out && c < d) ? rh_round(W[o * d + c], dtype) : 0.f;
    gw[o] = gb[o] = 0.f;
  }
  for (int b = 0; b < B; ++b) {
    const float xv = c < d ? to_f32<T>(x[static_cast<int64_t>(b) * Ls * d + c]) : 0.f;
    float dx = 0.f;
#pragma unroll
    for (int o = 0; o < RH_MAX_OUT; ++o) {
      if (o < n_out) {
        const float y = out[b * n_out + o];
        const float dpre = g_out[b * n_out + o] * (1.f - y * y);
        gw[o] = fmaf(dpre, xv, gw[o]);
        gb[o] += dpre;
        dx = fmaf(dpre, w[o], dx);
      }
    }
    if (c < d) d_x[static_cast<int64_t>(b) * Ls * d + c] = dx;
  }
  if (c < d) {
#pragma unroll
    for (int o = 0; o < RH_MAX_OUT; ++o)
      if (o < n_out) dW[o * d + c] = gw[o];
  }
  if (c == 0) {
#pragma unroll
    for (int o = 0; o < RH_MAX_OUT; ++o)
      if (o < n_out) db[o] = gb[o];
  }
}

}  // namespace me

using namespace me;

extern "C" int me_pooled_head_forward(const void* x, int dtype, const float* W, const float* bias, int B, int Ls,
                                      int d, int n_out, float* out, void* stream) {
  ME_CHECK(x && W && bias && out, "me_pooled_head_forward: NULL pointer");
  ME_CHECK(B > 0 && Ls > 0 && d > 0 && n_out > 0 && n_out <= RH_MAX_OUT, "me_pooled_head_forward: bad dims (n_out <= %d)",
           RH_MAX_OUT);
  ME_CHECK(dtype == ME_F32 || dtype == ME_BF16, "me_pooled_head_forward: bad dtype");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == ME_BF16)
    pooled_head_fwd_kernel<bf16><<<B, RH_THREADS, 0, st>>>(static_cast<const bf16*>(x), W, bias, Ls, d, n_out, dtype, out);
  else
    pooled_head_fwd_kernel<float><<<B, RH_THREADS, 0, st>>>(static_cast<const float*>(x), W, bias, Ls, d, n_out, dtype, out);
  ME_LAUNCH_CHECK();
  return 0;
}

extern "C" int me_pooled_head_backward(const float* g_out, const float* out, const void* x, int dtype, const float* W,
                                       int B, int Ls, int d, int n_out, float* dW, float* db, float* d_x,
                                       void* stream) {
  ME_CHECK(g_out && out && x && W && dW && db && d_x, "me_pooled_head_backward: NULL pointer");
  ME_CHECK(B > 0 && Ls > 0 && d > 0 && n_out > 0 && n_out <= RH_MAX_OUT, "me_pooled_head_backward: bad dims");
  ME_CHECK(dtype == ME_F32 || dtype == ME_BF16, "me_pooled_head_backward: bad dtype");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blocks = (d + RH_THREADS - 1) / RH_THREADS;
  if (dtype == ME_BF16)
    pooled_head_bwd_kernel<bf16><<<blocks, RH_THREADS, 0, st>>>(g_out, out, static_cast<const bf16*>(x), W, B, Ls, d,
                                                               n_out, dtype, dW, db, d_x);
  else
    pooled_head_bwd_kernel<float><<<blocks, RH_THREADS, 0, st>>>(g_out, out, static_cast<const float*>(x), W, B, Ls, d,
                                                                n_out, dtype, dW, db, d_x);
  ME_LAUNCH_CHECK();
  return 0;
}
