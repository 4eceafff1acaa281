// Next-token sampling of the generation loop (generate.py:122-189) as one launch, no host round trips:
//   NaN -> 0, special symbols -> -inf, log-softmax, per-sequence temperature (note / rest rule on the token
//   just fed, plus the repeat penalty), top-k (full descending sort), top-p on the sorted cumulative
//   softmax (the first entry always stays), softmax over what is left, draw, repeat-count update.
// One CTA per sequence; the vocabulary row (V <= 4096) is sorted with a bitonic network in shared memory,
// ties broken by the lower token id.  The draw is inverse-CDF with a caller-supplied uniform per sequence
// (torch.multinomial's generator stream cannot be reproduced outside PyTorch; tests pin the kept set, the
// probabilities and the draw for given uniforms).
#include "common.cuh"
#include "../../include/midi_emotion_b200.h"

namespace me {

constexpr int SP_THREADS = 256;

__device__ __forceinline__ float block_reduce(float v, float* scratch, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();  // scratch may still be read from the previous reduction
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float r = is_max ? -INFINITY : 0.f;
#pragma unroll
  for (int w = 0; w < SP_THREADS / 32; ++w) r = is_max ? fmaxf(r, scratch[w]) : r + scratch[w];
  return r;
}

template <typename T>
__global__ void __launch_bounds__(SP_THREADS)
sample_step_kernel(me_sample_args a, int NP) {
  extern __shared__ float sp_smem[];
  float* key = sp_smem;                                   // [NP] scaled log-probabilities, then probabilities
  int* idx = reinterpret_cast<int*>(sp_smem + NP);        // [NP] token ids
  float* scratch = sp_smem + 2 * NP;                      // [SP_THREADS / 32 + 8]
  const int b = blockIdx.x, tid = threadIdx.x, V = a.V;
  const T* row = static_cast<const T*>(a.logits) + static_cast<int64_t>(b) * a.ld_logits;

  // 1. load, NaN -> 0 (generate.py:124), excluded symbols -> -inf (:131-136)
  float mx = -INFINITY;
  for (int i = tid; i < NP; i += SP_THREADS) {
    float v = -INFINITY;
    if (i < V) {
      v = to_f32<T>(row[i]);
      if (v != v) v = 0.f;
      if (a.exclude && a.exclude[i]) v = -INFINITY;
    }
    key[i] = v;
    idx[i] = i;
    mx = fmaxf(mx, v);
  }
  mx = block_reduce(mx, scratch, true);
  // 2. log-softmax (:152)
  float se = 0.f;
  for (int i = tid; i < NP; i += SP_THREADS) se += (key[i] == -INFINITY) ? 0.f : expf(key[i] - mx);
  se = block_reduce(se, scratch, false);
  const float lse = mx + logf(se);
  // 3. temperature of this sequence (:139-162): rest temperature unless the token just fed is a TIMESHIFT
  //    tuple, then the repeat penalty  temp += max(0, log((count + 1) / 4) * coeff) * temp
  const int64_t prev = a.prev_tokens ? a.prev_tokens[b] : -1;
  float temp = a.temp_rest;
  if (a.is_timeshift && prev >= 0 && prev < V && a.is_timeshift[prev]) temp = a.temp_note;
  const int rc = a.repeat_counts ? a.repeat_counts[b] : 0;
  if (a.penalty_coeff > 0.f) temp += fmaxf(0.f, logf((static_cast<float>(rc) + 1.f) * 0.25f) * a.penalty_coeff) * temp;
  for (int i = tid; i < NP; i += SP_THREADS) key[i] = (key[i] == -INFINITY) ? -INFINITY : (key[i] - lse) / temp;
  __syncthreads();

  // 4. descending sort by (value, then lower token id): torch.topk with k = V (:165-170)
  for (int k = 2; k <= NP; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < NP; i += SP_THREADS) {
        const int p = i ^ j;
        if (p > i) {
          const float ka = key[i], kb = key[p];
          const int ia = idx[i], ib = idx[p];
          const bool a_first = (ka > kb) || (ka == kb && ia < ib);   // a belongs before b in descending order
          const bool desc = (i & k) == 0;
          if (desc ? !a_first : a_first) {
            key[i] = kb; key[p] = ka;
            idx[i] = ib; idx[p] = ia;
          }
        }
      }
      __syncthreads();
    }
  }
  const int k_eff = (a.top_k <= 0 || a.top_k > V) ? V : a.top_k;

  // 5. top-p (:173-177): softmax over the k_eff sorted entries, inclusive cumulative sum, drop where it
  //    exceeds top_p except the first entry.  key[0] is the maximum.
  const float top = key[0];
  float s = 0.f;
  for (int i = tid; i < k_eff; i += SP_THREADS) s += (key[i] == -INFINITY) ? 0.f : expf(key[i] - top);
  s = block_reduce(s, scratch, false);
  // block-wide inclusive scan over the sorted order: each thread owns a contiguous segment
  const int seg = (k_eff + SP_THREADS - 1) / SP_THREADS;
  const int i0 = tid * seg, i1 = min(k_eff, i0 + seg);
  float local = 0.f;
  for (int i = i0; i < i1; ++i) local += (key[i] == -INFINITY) ? 0.f : expf(key[i] - top) / s;
  __shared__ float seg_sum[SP_THREADS];
  seg_sum[tid] = local;
  __syncthreads();
  float before = 0.f;
  for (int t = 0; t < tid; ++t) before += seg_sum[t];
  const bool use_p = a.top_p > 0.f && a.top_p < 1.f;
  float run = before, kept_mass = 0.f;
  for (int i = i0; i < i1; ++i) {
    const float pr = (key[i] == -INFINITY) ? 0.f : expf(key[i] - top) / s;
    run += pr;
    const bool keep = !use_p || i == 0 || !(run > a.top_p);
    const float kp = keep ? pr : 0.f;
    key[i] = kp;           // from here on: unnormalised kept probability
    kept_mass += kp;
  }
  for (int i = k_eff + tid; i < NP; i += SP_THREADS) key[i] = 0.f;
  kept_mass = block_reduce(kept_mass, scratch, false);

  // 6. softmax over what is left (:179), number of choices (:186), inverse-CDF draw (:182-183)
  float cnt = 0.f, local2 = 0.f;
  for (int i = i0; i < i1; ++i) {
    const float pr = key[i] / kept_mass;
    key[i] = pr;
    local2 += pr;
    cnt += pr > 0.f ? 1.f : 0.f;
  }
  cnt = block_reduce(cnt, scratch, false);
  __syncthreads();
  seg_sum[tid] = local2;
  __syncthreads();
  if (a.out_probs) {
    float* op = a.out_probs + static_cast<int64_t>(b) * V;
    for (int i = tid; i < NP; i += SP_THREADS)
      if (idx[i] < V) op[idx[i]] = i < k_eff ? key[i] : 0.f;
  }
  if (tid == 0) {
    const float u = a.uniforms[b];
    // first sorted position whose inclusive cumulative probability exceeds u (the last kept one otherwise)
    float c = 0.f;
    int t = 0;
    for (; t < SP_THREADS; ++t) {
      if (c + seg_sum[t] > u) break;
      c += seg_sum[t];
    }
    int pick = -1;
    if (t < SP_THREADS) {
      for (int i = t * seg; i < min(k_eff, (t + 1) * seg); ++i) {
        c += key[i];
        if (c > u && key[i] > 0.f) { pick = i; break; }
      }
    }
    if (pick < 0) {  // u at the very top of the range (rounding): the last entry with non-zero probability
      for (int i = k_eff - 1; i >= 0; --i)
        if (key[i] > 0.f) { pick = i; break; }
      if (pick < 0) pick = 0;
    }
    a.out_tokens[b] = idx[pick];
    const int n = static_cast<int>(cnt + 0.5f);
    if (a.out_num_choices) a.out_num_choices[b] = n;
    if (a.repeat_counts) a.repeat_counts[b] = n <= 2 ? rc + 1 : rc / 2;   // :187-189
  }
}

}  // namespace me

using namespace me;

extern "C" int me_sample_step(const me_sample_args* a) {
  ME_CHECK(a != nullptr, "me_sample_step: NULL args");
  ME_CHECK(a->B > 0 && a->V > 0 && a->V <= 4096, "me_sample_step: bad sizes (B=%d, V=%d; V <= 4096)", a->B, a->V);
  ME_CHECK(a->ld_logits >= a->V, "me_sample_step: ld_logits %d < V %d", a->ld_logits, a->V);
  ME_CHECK(a->logits && a->uniforms && a->out_tokens, "me_sample_step: logits, uniforms and out_tokens are required");
  ME_CHECK(a->logits_dtype == ME_F32 || a->logits_dtype == ME_BF16, "me_sample_step: bad logits dtype %d", a->logits_dtype);
  ME_CHECK(a->temp_note > 0.f && a->temp_rest > 0.f, "me_sample_step: temperatures must be positive");
  int NP = 1;
  while (NP < a->V) NP <<= 1;
  if (NP < SP_THREADS) NP = SP_THREADS;
  const size_t smem = (2 * static_cast<size_t>(NP) + 64) * sizeof(float);
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  if (a->logits_dtype == ME_BF16) sample_step_kernel<bf16><<<a->B, SP_THREADS, smem, st>>>(*a, NP);
  else sample_step_kernel<float><<<a->B, SP_THREADS, smem, st>>>(*a, NP);
  ME_LAUNCH_CHECK();
  return 0;
}
