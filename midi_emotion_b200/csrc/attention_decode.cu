// Decode-step relative attention over the KV cache (one query row per sequence), bf16 storage, fp32 math.
//
// HBM-bound: per (batch, head) the step streams t+1 rows of K and V once (2 * (t+1) * dh * 2 bytes); the
// matching rows of E (shared by every batch/head, 256 KB per layer) come out of L2.
// One CTA per (batch, head), 8 warps, lane = key: every lane reads whole 128-byte rows with 16-byte loads
// (24 independent loads in flight per key), keeps its own running (max, sum, o[dh]) and the partial
// states are merged once at the end (warp butterfly, then across warps through shared memory).
// The position comes from device memory so the launch is CUDA-graph friendly.
#include "common.cuh"
#include "../../include/midi_emotion_b200.h"

namespace me {

constexpr int AD_WARPS = 8;

struct AdParams {
  int H, max_seq, q_pos0;
  int64_t q_sb, q_sh, k_sb, k_sh, k_sj, v_sb, v_sh, v_sj, o_sb, keypad_ld;
  const int32_t* pos_dev;
  const uint8_t* keypad;
  float scale_log2;
};

// a row of DH bf16 as DH/8 raw 16-byte words; conversion to fp32 happens at the point of use so that the
// loads of K, E and V rows of a key are all in flight together
template <int DH>
__device__ __forceinline__ void load_row_raw(const bf16* __restrict__ row, uint4 (&r)[DH / 8]) {
#pragma unroll
  for (int c = 0; c < DH / 8; ++c) r[c] = __ldg(reinterpret_cast<const uint4*>(row) + c);
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }
template <int DH>
__device__ __forceinline__ void dot_row(const float (&qf)[DH], const uint4 (&r)[DH / 8], float& s0, float& s1) {
#pragma unroll
  for (int c = 0; c < DH / 8; ++c) {
    const uint32_t w[4] = {r[c].x, r[c].y, r[c].z, r[c].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s0 = fmaf(qf[c * 8 + 2 * e], bf_lo(w[e]), s0);
      s1 = fmaf(qf[c * 8 + 2 * e + 1], bf_hi(w[e]), s1);
    }
  }
}

template <int DH>
__global__ void __launch_bounds__(AD_WARPS * 32)
attn_decode_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                   const bf16* __restrict__ E, bf16* __restrict__ out, AdParams p) {
  __shared__ float red[AD_WARPS][DH + 2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int t = p.pos_dev ? *p.pos_dev : p.q_pos0;  // position of the query == index of the newest key
  float qf[DH];
  {
    uint4 qr[DH / 8];
    load_row_raw<DH>(q + b * p.q_sb + h * p.q_sh, qr);
#pragma unroll
    for (int c = 0; c < DH / 8; ++c) {
      const uint32_t w[4] = {qr[c].x, qr[c].y, qr[c].z, qr[c].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        qf[c * 8 + 2 * e] = bf_lo(w[e]);
        qf[c * 8 + 2 * e + 1] = bf_hi(w[e]);
      }
    }
  }
  const bf16* kb = k + b * p.k_sb + h * p.k_sh;
  const bf16* vb = v + b * p.v_sb + h * p.v_sh;
  const bf16* Eb = E + static_cast<int64_t>(p.max_seq - 1 - t) * DH;  // row for key j is Eb + j*DH
  const uint8_t* kp = p.keypad ? p.keypad + b * p.keypad_ld : nullptr;

  float m = -INFINITY, l = 0.f;
  float o[DH];
#pragma unroll
  for (int c = 0; c < DH; ++c) o[c] = 0.f;

  for (int j = warp * 32 + lane; j <= t; j += AD_WARPS * 32) {
    if (kp && kp[j]) continue;
    uint4 kr[DH / 8], er[DH / 8], vr[DH / 8];
    load_row_raw<DH>(kb + j * p.k_sj, kr);
    load_row_raw<DH>(Eb + static_cast<int64_t>(j) * DH, er);
    load_row_raw<DH>(vb + j * p.v_sj, vr);
    float s0 = 0.f, s1 = 0.f;
    dot_row<DH>(qf, kr, s0, s1);
    dot_row<DH>(qf, er, s0, s1);
    const float x = (s0 + s1) * p.scale_log2;
    if (x > m) {  // rescale the running state (rare once the maximum has settled)
      const float a = fast_exp2(m - x);
      l *= a;
#pragma unroll
      for (int c = 0; c < DH; ++c) o[c] *= a;
      m = x;
    }
    const float pj = fast_exp2(x - m);
    l += pj;
#pragma unroll
    for (int c = 0; c < DH / 8; ++c) {
      const uint32_t w[4] = {vr[c].x, vr[c].y, vr[c].z, vr[c].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        o[c * 8 + 2 * e] = fmaf(pj, bf_lo(w[e]), o[c * 8 + 2 * e]);
        o[c * 8 + 2 * e + 1] = fmaf(pj, bf_hi(w[e]), o[c * 8 + 2 * e + 1]);
      }
    }
  }

  // merge the 32 lane states of the warp, then the warps of the block
  const float wm = warp_max(m);
  const float sc = (m == -INFINITY) ? 0.f : fast_exp2(m - wm);
  l = warp_sum(l * sc);
#pragma unroll
  for (int c = 0; c < DH; ++c) o[c] = warp_sum(o[c] * sc);
  if (lane == 0) {
    red[warp][DH] = wm;
    red[warp][DH + 1] = l;
#pragma unroll
    for (int c = 0; c < DH; ++c) red[warp][c] = o[c];
  }
  __syncthreads();
  if (warp == 0) {
    float gm = -INFINITY;
#pragma unroll
    for (int w = 0; w < AD_WARPS; ++w) gm = fmaxf(gm, red[w][DH]);
    float gl = 0.f;
    float acc0 = 0.f, acc1 = 0.f;  // lane owns columns lane and lane + 32
#pragma unroll
    for (int w = 0; w < AD_WARPS; ++w) {
      const float wmx = red[w][DH];
      const float s = (wmx == -INFINITY) ? 0.f : fast_exp2(wmx - gm);
      gl += red[w][DH + 1] * s;
      if (lane < DH) acc0 += red[w][lane] * s;
      if (lane + 32 < DH) acc1 += red[w][lane + 32] * s;
    }
    const float inv = gl > 0.f ? 1.f / gl : 0.f;  // fully masked row -> 0
    bf16* orow = out + b * p.o_sb + h * DH;
    if (lane < DH) orow[lane] = __float2bfloat16_rn(acc0 * inv);
    if (lane + 32 < DH) orow[lane + 32] = __float2bfloat16_rn(acc1 * inv);
  }
}

int launch_attn_decode(const me_attn_args* a) {
  AdParams p;
  p.H = a->H; p.max_seq = a->max_seq; p.q_pos0 = a->q_pos0;
  p.q_sb = a->q_sb; p.q_sh = a->q_sh;
  p.k_sb = a->k_sb; p.k_sh = a->k_sh; p.k_sj = a->k_sj;
  p.v_sb = a->v_sb; p.v_sh = a->v_sh; p.v_sj = a->v_sj;
  p.o_sb = a->o_sb; p.keypad_ld = a->keypad_ld;
  p.pos_dev = a->pos_dev; p.keypad = a->keypad;
  p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(a->dh));
  dim3 grid(a->H, a->B);
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  const bf16* q = static_cast<const bf16*>(a->q);
  const bf16* k = static_cast<const bf16*>(a->k);
  const bf16* v = static_cast<const bf16*>(a->v);
  const bf16* E = static_cast<const bf16*>(a->E);
  bf16* out = static_cast<bf16*>(a->out);
  if (a->dh == 64) attn_decode_kernel<64><<<grid, AD_WARPS * 32, 0, st>>>(q, k, v, E, out, p);
  else if (a->dh == 48) attn_decode_kernel<48><<<grid, AD_WARPS * 32, 0, st>>>(q, k, v, E, out, p);
  else attn_decode_kernel<32><<<grid, AD_WARPS * 32, 0, st>>>(q, k, v, E, out, p);
  ME_LAUNCH_CHECK();
  return 0;
}

}  // namespace me
