// HBM-bound kernels of the hot path: input stage (embedding + conditioning + positional table),
// residual-add + LayerNorm forward/backward, column sums, dtype conversion, KV-cache append.
// All of them stream each byte once with 16-byte vector accesses where alignment allows.
#include "common.cuh"
#include "../../include/midi_emotion_b200.h"

namespace me {

// =====================================================================================
// Input stage (music_multi.py:89-102, music_continuous_token.py:81-100)
// =====================================================================================
template <typename T>
__global__ void embed_fwd_kernel(const int64_t* __restrict__ tokens, const float* __restrict__ cond,
                                 const float* __restrict__ emb_w, const float* __restrict__ cw0,
                                 const float* __restrict__ cb0, const float* __restrict__ cw1,
                                 const float* __restrict__ cb1, const float* __restrict__ pe, int B, int L,
                                 int d, int dc, int V, int mode, int pad_token, float p, uint64_t seed,
                                 const int32_t* __restrict__ t_dev, int T_max, float* __restrict__ x_f32,
                                 T* __restrict__ x_T, uint8_t* __restrict__ keypad) {
  // one block per output row (b, s); decode mode (t_dev != NULL): L == 1, s = *t_dev
  const bool ctoken = (mode == ME_COND_CONTINUOUS_TOKEN);
  const int row = blockIdx.x;
  int b, s, pos;
  int64_t out_row;
  if (t_dev != nullptr) {  // decode: single position per sequence, never a prefix slot
    b = row;
    pos = *t_dev;
    s = ctoken ? 2 : 0;
    out_row = b;
  } else {
    const int Ls = ctoken ? L + 2 : L;
    b = row / Ls;
    s = row - b * Ls;
    pos = s;
    out_row = row;
  }
  const int de = d - dc;
  const float scale = sqrtf(static_cast<float>(de));
  const float inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  int64_t tok = -1;
  const bool prefix = ctoken && t_dev == nullptr && s < 2;
  if (!prefix) {
    tok = (t_dev != nullptr) ? tokens[b] : tokens[static_cast<int64_t>(b) * L + (ctoken ? s - 2 : s)];
  }
  if (threadIdx.x == 0 && keypad != nullptr) {
    const uint8_t kp = (!prefix && tok == pad_token) ? 1 : 0;
    if (t_dev != nullptr) keypad[static_cast<int64_t>(b) * T_max + pos] = kp;
    else keypad[row] = kp;
  }
  const bool tok_ok = tok >= 0 && tok < V;
  float c0 = 0.f, c1 = 0.f;
  if (mode == ME_COND_CONTINUOUS_CONCAT || prefix) {
    c0 = cond[b * 2 + 0];
    c1 = cond[b * 2 + 1];
  }
  const float* perow = pe + static_cast<int64_t>(pos) * d;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float v;
    // bf16 path: the condition Linears run under autocast in the reference (inputs, weights and bias cast to
    // bf16, bf16 result), while the embedding lookup, its scale and the positional add stay fp32
    constexpr bool lowp = sizeof(T) == 2;
    auto rb = [](float x) { return lowp ? __bfloat162float(__float2bfloat16_rn(x)) : x; };
    if (prefix) {
      // Linear(1, d): cond[b, s] * w[c, 0] + bias[c]
      v = (s == 0) ? rb(fmaf(rb(c0), rb(cw0[c]), rb(cb0[c]))) : rb(fmaf(rb(c1), rb(cw1[c]), rb(cb1[c])));
    } else if (c < de) {
      v = tok_ok ? emb_w[tok * de + c] * scale : 0.f;
    } else {
      const int j = c - de;  // Linear(2, dc)
      v = rb(rb(cb0[j]) + (rb(c0) * rb(cw0[j * 2 + 0]) + rb(c1) * rb(cw0[j * 2 + 1])));
    }
    v += perow[c];
    v *= dropout_scale(p, inv_keep, seed, static_cast<uint64_t>(out_row) * d + c);
    x_f32[out_row * d + c] = v;
    if (x_T != nullptr && static_cast<void*>(x_T) != static_cast<void*>(x_f32))
      x_T[out_row * d + c] = from_f32<T>(v);
  }
}

// dEmb[tok] += dx * sqrt(de) (pad rows skipped: Embedding(padding_idx=0) gets zero gradient);
// fc_condition grads reduced over the sequence.  A block owns EB_ROWS consecutive rows: the token part
// goes out as 16-byte vector atomics (one per 4 columns), the condition part is summed over the block's
// rows in registers first (one atomic per column per block instead of one per element).
constexpr int EB_ROWS = 16;
constexpr int EB_THREADS = 256;

__global__ void __launch_bounds__(EB_THREADS)
embed_bwd_kernel(const float* __restrict__ dx, const bf16* __restrict__ dx_T, const int64_t* __restrict__ tokens,
                 const float* __restrict__ cond, int B, int L, int d, int dc, int V, int mode, int pad_token, float p,
                 uint64_t seed, float* __restrict__ d_emb, float* __restrict__ d_cw0, float* __restrict__ d_cb0,
                 float* __restrict__ d_cw1, float* __restrict__ d_cb1) {
  // the gradient w.r.t. the stage's output is dx (+ dx_T: the compute-type part the first layer's QKV projection
  // leaves, me_layer_bwd_args.d_x_T) -- added here instead of by a separate pass over [M, d]
  auto gx = [&](int64_t idx) { return dx_T ? dx[idx] + __bfloat162float(dx_T[idx]) : dx[idx]; };
  const bool ctoken = (mode == ME_COND_CONTINUOUS_TOKEN);
  const int Ls = ctoken ? L + 2 : L;
  const int64_t rows = static_cast<int64_t>(B) * Ls;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * EB_ROWS;
  const int64_t r1 = min(rows, r0 + EB_ROWS);
  const int de = d - dc;
  const float scale = sqrtf(static_cast<float>(de));
  const float inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  const uint32_t seed32 = dropout_seed32(seed);
  const bool vec = (de % 4 == 0) && (d % 4 == 0);

  // ---- token embedding columns (and the two prefix rows of continuous_token)
  for (int64_t row = r0; row < r1; ++row) {
    const int b = static_cast<int>(row / Ls);
    const int s = static_cast<int>(row - static_cast<int64_t>(b) * Ls);
    if (ctoken && s < 2) {  // Linear(1, d) prefix vectors: rare rows, element-wise atomics
      const float cv = cond[b * 2 + s];
      float* dw = s == 0 ? d_cw0 : d_cw1;
      float* db = s == 0 ? d_cb0 : d_cb1;
      for (int c = threadIdx.x; c < d; c += EB_THREADS) {
        const float g = gx(row * d + c) * dropout_scale(p, inv_keep, seed, static_cast<uint64_t>(row) * d + c);
        atomicAdd(&dw[c], g * cv);
        atomicAdd(&db[c], g);
      }
      continue;
    }
    const int64_t tok = tokens[static_cast<int64_t>(b) * L + (ctoken ? s - 2 : s)];
    if (tok == pad_token || tok < 0 || tok >= V) continue;
    float* erow = d_emb + tok * de;
    if (vec) {
      for (int c = 4 * threadIdx.x; c < de; c += 4 * EB_THREADS) {
        float4 g = *reinterpret_cast<const float4*>(dx + row * d + c);
        if (dx_T) {
          const uint2 u = *reinterpret_cast<const uint2*>(dx_T + row * d + c);
          const float2 t0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
          const float2 t1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
          g.x += t0.x; g.y += t0.y; g.z += t1.x; g.w += t1.y;
        }
        float dm[4];
        dropout_scale4(p, inv_keep, seed32, static_cast<uint64_t>(row) * d + c, dm);
        atomicAdd(reinterpret_cast<float4*>(erow + c),
                  make_float4(g.x * dm[0] * scale, g.y * dm[1] * scale, g.z * dm[2] * scale, g.w * dm[3] * scale));
      }
    } else {
      for (int c = threadIdx.x; c < de; c += EB_THREADS)
        atomicAdd(&erow[c], gx(row * d + c) * scale * dropout_scale(p, inv_keep, seed, static_cast<uint64_t>(row) * d + c));
    }
  }
  // ---- concatenated condition columns: Linear(2, dc), summed over this block's rows
  if (dc > 0) {
    for (int j = threadIdx.x; j < dc; j += EB_THREADS) {
      float a0 = 0.f, a1 = 0.f, ab = 0.f;
      for (int64_t row = r0; row < r1; ++row) {
        const int b = static_cast<int>(row / Ls);
        const float g = gx(row * d + de + j) *
                        dropout_scale(p, inv_keep, seed, static_cast<uint64_t>(row) * d + de + j);
        a0 = fmaf(g, cond[b * 2 + 0], a0);
        a1 = fmaf(g, cond[b * 2 + 1], a1);
        ab += g;
      }
      atomicAdd(&d_cw0[j * 2 + 0], a0);
      atomicAdd(&d_cw0[j * 2 + 1], a1);
      atomicAdd(&d_cb0[j], ab);
    }
  }
}

// =====================================================================================
// out = LayerNorm(x_res + dropout(y))   (music_multi.py:128-129,133-134; eps 1e-6)
// One warp per row, the row lives in registers as float4 chunks (lane l owns columns 4*(l + 32k)..+3),
// two-pass statistics in fp32 (mean, then biased variance), 16-byte global accesses throughout.
// =====================================================================================
constexpr int LN_THREADS = 256;

template <typename T> struct Vec4 {};
template <> struct Vec4<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec4<bf16> {
  static __device__ __forceinline__ void load(const bf16* p, float (&v)[4]) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static __device__ __forceinline__ void store(bf16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
  }
};

// The normalisation itself; one definition so that a deferred output re-derived by the next LayerNorm has the bits
// the producing LayerNorm would have written (explicit fma: no contraction choice left to the compiler).
__device__ __forceinline__ float ln_apply(float z, float mean, float rstd, float g, float b) {
  return __fmaf_rn(__fmul_rn(__fsub_rn(z, mean), rstd), g, b);
}

template <typename T, int NCH>
__global__ void __launch_bounds__(LN_THREADS)
add_ln_fwd_kernel(const float* __restrict__ x_res, const T* __restrict__ y, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float eps, int M, int d, float p, uint64_t seed,
                  float* __restrict__ out_f32, T* __restrict__ out_T, float* __restrict__ zsave,
                  float* __restrict__ mean_out, float* __restrict__ rstd_out, const float* __restrict__ src_mean,
                  const float* __restrict__ src_rstd, const float* __restrict__ src_gamma,
                  const float* __restrict__ src_beta) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int wpb = LN_THREADS / 32;
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * wpb + (threadIdx.x >> 5);
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * wpb;
  const float inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  const uint32_t seed32 = dropout_seed32(seed);
  const float inv_d = 1.f / static_cast<float>(d);
  const bool alias = static_cast<const void*>(out_T) == static_cast<const void*>(out_f32);
  extern __shared__ float ln_smem[];  // gamma | beta (| source gamma | beta): keeps them out of every thread's registers
  // src_mean != NULL: x_res holds the pre-normalisation sums of the LayerNorm that produced the residual input, and
  // the input is re-derived here (me_layer_args.xin_*) with that LayerNorm's own expression, ln_apply()
  const bool lazy = src_mean != nullptr;
  for (int c = threadIdx.x; c < d; c += LN_THREADS) {
    ln_smem[c] = gamma[c];
    ln_smem[d + c] = beta[c];
    if (lazy) {
      ln_smem[2 * d + c] = src_gamma[c];
      ln_smem[3 * d + c] = src_beta[c];
    }
  }
  __syncthreads();
  for (int64_t row = warp0; row < M; row += nwarps) {
    // All loads of the row are issued before the first use (no branches in between: the column guard is
    // applied by clamping the address, a branch per chunk made the compiler serialise load -> use per chunk,
    // i.e. one memory latency per chunk instead of one per row).
    float z[NCH][4], yv[NCH][4];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int c = 4 * (lane + 32 * k);
      const int64_t idx = row * d + (c < d ? c : 0);
      Vec4<T>::load(y + idx, yv[k]);
      Vec4<float>::load(x_res + idx, z[k]);
    }
    if (lazy) {
      const float smu = src_mean[row], srs = src_rstd[row];
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        const int c = 4 * (lane + 32 * k);
        float gm[4], bt[4];
        Vec4<float>::load(ln_smem + 2 * d + (c < d ? c : 0), gm);
        Vec4<float>::load(ln_smem + 3 * d + (c < d ? c : 0), bt);
#pragma unroll
        for (int e = 0; e < 4; ++e) z[k][e] = ln_apply(z[k][e], smu, srs, gm[e], bt[e]);
      }
    }
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int c = 4 * (lane + 32 * k);
      const bool ok = c < d;
      float dm[4];
      dropout_scale4(p, inv_keep, seed32, static_cast<uint64_t>(row * d + c), dm);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        z[k][e] = ok ? yv[k][e] * dm[e] + z[k][e] : 0.f;
        sum += z[k][e];
      }
    }
    const float mean = warp_sum(sum) * inv_d;
    float vs = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      if (4 * (lane + 32 * k) < d) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float dz = z[k][e] - mean;
          vs += dz * dz;
        }
      }
    }
    const float rstd = 1.f / sqrtf(warp_sum(vs) * inv_d + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int c = 4 * (lane + 32 * k);
      if (c < d) {
        const int64_t idx = row * d + c;
        float o[4], gm[4], bt[4];
        Vec4<float>::load(ln_smem + c, gm);
        Vec4<float>::load(ln_smem + d + c, bt);
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = ln_apply(z[k][e], mean, rstd, gm[e], bt[e]);
        if (zsave) Vec4<float>::store(zsave + idx, z[k]);
        if (out_f32) Vec4<float>::store(out_f32 + idx, o);
        if (out_T != nullptr && !alias) Vec4<T>::store(out_T + idx, o);
      }
    }
  }
}

// LayerNorm backward.  Persistent warps walk the rows.  The column reductions (d_gamma, d_beta and,
// optionally, the column sums of the masked dy = bias gradient of the preceding Linear) accumulate in a
// private shared-memory strip per warp (plain read-modify-write, no atomics, no register arrays, which
// keeps occupancy high enough to cover HBM latency); the strips are summed per block at the end and
// flushed with one atomicAdd per column per block.
template <typename T, int NCH>
__global__ void __launch_bounds__(LN_THREADS, 2)   // two blocks per SM (<= 128 registers): the kernel lives on occupancy
add_ln_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ dout_add,
                  const bf16* __restrict__ dout_add_T, const float* __restrict__ z,
                  const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                  int M, int d, float p, uint64_t seed, float* __restrict__ dz_f32, T* __restrict__ dy_T,
                  float* __restrict__ d_gamma, float* __restrict__ d_beta, float* __restrict__ d_ybias) {
  extern __shared__ float ln_smem[];  // gamma[d] | per warp: dg[d] db[d] dyb[d]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int wpb = LN_THREADS / 32;
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * wpb + warp;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * wpb;
  float* gam_s = ln_smem;
  float* acc = ln_smem + d + warp * 3 * d;
  for (int c = threadIdx.x; c < d; c += LN_THREADS) gam_s[c] = gamma[c];
  for (int c = lane; c < 3 * d; c += 32) acc[c] = 0.f;
  __syncthreads();
  const float inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  const uint32_t seed32 = dropout_seed32(seed);
  const float inv_d = 1.f / static_cast<float>(d);
  const bool alias = static_cast<const void*>(dy_T) == static_cast<const void*>(dz_f32);
  for (int64_t row = warp0; row < M; row += nwarps) {
    // all loads of the row first, branch-free (see the forward kernel)
    float dy[NCH][4], xh[NCH][4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int c = 4 * (lane + 32 * k);
      const int64_t idx = row * d + (c < d ? c : 0);
      Vec4<float>::load(dout + idx, dy[k]);
      Vec4<float>::load(z + idx, xh[k]);
    }
    if (dout_add) {
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        const int c = 4 * (lane + 32 * k);
        float t[4];
        Vec4<float>::load(dout_add + row * d + (c < d ? c : 0), t);
#pragma unroll
        for (int e = 0; e < 4; ++e) dy[k][e] += t[e];
      }
    }
    if (dout_add_T) {   // the sub-layer's input gradient in the compute type (a bf16 GEMM output, as autocast has it)
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        const int c = 4 * (lane + 32 * k);
        float t[4];
        Vec4<bf16>::load(dout_add_T + row * d + (c < d ? c : 0), t);
#pragma unroll
        for (int e = 0; e < 4; ++e) dy[k][e] += t[e];
      }
    }
    const float mu = mean[row], rs = rstd[row];
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const bool ok = 4 * (lane + 32 * k) < d;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        xh[k][e] = ok ? (xh[k][e] - mu) * rs : 0.f;
        dy[k][e] = ok ? dy[k][e] : 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int c = 4 * (lane + 32 * k);
      if (c < d) {
        float gm[4], ag[4], ab[4];
        Vec4<float>::load(gam_s + c, gm);
        Vec4<float>::load(acc + c, ag);
        Vec4<float>::load(acc + d + c, ab);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float g = dy[k][e] * gm[e];
          s1 += g;
          s2 += g * xh[k][e];
          ag[e] += dy[k][e] * xh[k][e];
          ab[e] += dy[k][e];
          dy[k][e] = g;  // from here on: dy * gamma
        }
        Vec4<float>::store(acc + c, ag);
        Vec4<float>::store(acc + d + c, ab);
      }
    }
    s1 = warp_sum(s1) * inv_d;
    s2 = warp_sum(s2) * inv_d;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int c = 4 * (lane + 32 * k);
      if (c < d) {
        const int64_t idx = row * d + c;
        float dzv[4], dm[4], dyv[4], ay[4];
        dropout_scale4(p, inv_keep, seed32, static_cast<uint64_t>(idx), dm);
        Vec4<float>::load(acc + 2 * d + c, ay);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          dzv[e] = rs * (dy[k][e] - s1 - xh[k][e] * s2);
          dyv[e] = dzv[e] * dm[e];
          ay[e] += dyv[e];
        }
        Vec4<float>::store(acc + 2 * d + c, ay);
        if (dz_f32) Vec4<float>::store(dz_f32 + idx, dzv);
        if (dy_T != nullptr && !alias) Vec4<T>::store(dy_T + idx, dyv);
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 3 * d; c += LN_THREADS) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < wpb; ++w) sum += ln_smem[d + w * 3 * d + c];
    if (c < d) atomicAdd(&d_gamma[c], sum);
    else if (c < 2 * d) atomicAdd(&d_beta[c - d], sum);
    else if (d_ybias) atomicAdd(&d_ybias[c - 2 * d], sum);
  }
}

// =====================================================================================
// column sums (bias gradients): 8 columns per thread (16-byte loads), rows split over blockIdx.y
// =====================================================================================
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ X, int M, int N, int ldx, int rows_per_block,
                              float* __restrict__ out) {
  constexpr int VEC = 16 / sizeof(T);
  const int n0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (n0 >= N) return;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_block;
  const int64_t r1 = min(static_cast<int64_t>(M), r0 + rows_per_block);
  float acc[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
  const bool vec_ok = (n0 + VEC <= N) && (ldx % VEC == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  if (vec_ok) {
#pragma unroll 8   // eight 16-byte loads in flight per thread: at four the kernel sat at half the HBM rate
    for (int64_t r = r0; r < r1; ++r) {
      const uint4 u = *reinterpret_cast<const uint4*>(X + r * ldx + n0);
      const T* v = reinterpret_cast<const T*>(&u);
#pragma unroll
      for (int e = 0; e < VEC; ++e) acc[e] += to_f32<T>(v[e]);
    }
  } else {
    for (int64_t r = r0; r < r1; ++r)
      for (int e = 0; e < VEC; ++e)
        if (n0 + e < N) acc[e] += to_f32<T>(X[r * ldx + n0 + e]);
  }
#pragma unroll
  for (int e = 0; e < VEC; ++e)
    if (n0 + e < N) atomicAdd(&out[n0 + e], acc[e]);
}

template <typename TS, typename TD>
__global__ void convert2d_kernel(const TS* __restrict__ src, int ld_src, TD* __restrict__ dst, int ld_dst,
                                 int rows, int cols) {
  const int64_t total = static_cast<int64_t>(rows) * ld_dst;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / ld_dst;
    const int c = static_cast<int>(i - r * ld_dst);
    dst[i] = from_f32<TD>(c < cols ? to_f32<TS>(src[r * ld_src + c]) : 0.f);
  }
}

// =====================================================================================
// KV cache append: rows of a packed QKV buffer -> [B, H, T_max, dh] caches
// =====================================================================================
template <typename T>
__global__ void kv_write_kernel(const T* __restrict__ qkv, int B, int Ls, int H, int dh, T* __restrict__ kc,
                                T* __restrict__ vc, int T_max, int pos0, const int32_t* __restrict__ t_dev) {
  pdl_launch_dependents();
  pdl_wait();
  const int d = H * dh;
  const int64_t total = static_cast<int64_t>(B) * Ls * d;
  const int p0 = t_dev ? *t_dev : pos0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % d);
    const int64_t row = i / d;
    const int s = static_cast<int>(row % Ls);
    const int b = static_cast<int>(row / Ls);
    const int h = c / dh, e = c - h * dh;
    const int64_t dst = ((static_cast<int64_t>(b) * H + h) * T_max + (p0 + s)) * dh + e;
    kc[dst] = qkv[row * 3 * d + d + c];
    vc[dst] = qkv[row * 3 * d + 2 * d + c];
  }
}

int launch_kv_write(const void* qkv, int dtype, int B, int Ls, int H, int dh, void* kc, void* vc, int T_max,
                    int pos0, const int32_t* t_dev, cudaStream_t st) {
  const int64_t total = static_cast<int64_t>(B) * Ls * H * dh;
  const int64_t want = (total + 255) / 256;
  const int blocks = static_cast<int>(want < 148 * 8 ? want : 148 * 8);
  if (dtype == ME_BF16)
    launch_pdl(kv_write_kernel<bf16>, dim3(blocks), dim3(256), 0, st, static_cast<const bf16*>(qkv), B, Ls, H, dh,
               static_cast<bf16*>(kc), static_cast<bf16*>(vc), T_max, pos0, t_dev);
  else
    kv_write_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(qkv), B, Ls, H, dh,
                                                   static_cast<float*>(kc), static_cast<float*>(vc), T_max, pos0,
                                                   t_dev);
  ME_LAUNCH_CHECK();
  return 0;
}

int launch_embed(const int64_t* tokens, const float* cond, const float* emb_w, const float* cw0,
                 const float* cb0, const float* cw1, const float* cb1, const float* pe, int B, int L, int d,
                 int d_cond, int V, int mode, int pad_token, float p, uint64_t seed, const int32_t* t_dev,
                 int T_max, int dtype, float* x_f32, void* x_T, uint8_t* keypad, cudaStream_t st) {
  const int Ls = (mode == ME_COND_CONTINUOUS_TOKEN && t_dev == nullptr) ? L + 2 : L;
  const int rows = B * Ls;
  const int threads = d >= 512 ? 256 : 128;
  if (dtype == ME_BF16)
    embed_fwd_kernel<bf16><<<rows, threads, 0, st>>>(tokens, cond, emb_w, cw0, cb0, cw1, cb1, pe, B, L, d, d_cond,
                                                     V, mode, pad_token, p, seed, t_dev, T_max, x_f32,
                                                     static_cast<bf16*>(x_T), keypad);
  else
    embed_fwd_kernel<float><<<rows, threads, 0, st>>>(tokens, cond, emb_w, cw0, cb0, cw1, cb1, pe, B, L, d,
                                                      d_cond, V, mode, pad_token, p, seed, t_dev, T_max, x_f32,
                                                      static_cast<float*>(x_T), keypad);
  ME_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int ln_fwd_dispatch(int nch, int blocks, cudaStream_t st, const float* x_res, const T* y, const float* gamma,
                           const float* beta, float eps, int M, int d, float p, uint64_t seed, float* out_f32,
                           T* out_T, float* z, float* mean, float* rstd, const float* src_mean,
                           const float* src_rstd, const float* src_gamma, const float* src_beta) {
#define ME_LN_FWD(N)                                                                                        \
  launch_pdl(add_ln_fwd_kernel<T, N>, dim3(blocks), dim3(LN_THREADS), (src_mean ? 4 : 2) * d * sizeof(float), st, \
             x_res, y, gamma, beta, eps, M, d, p, seed, out_f32, out_T, z, mean, rstd, src_mean, src_rstd, src_gamma, src_beta)
  switch (nch) {
    case 1: ME_LN_FWD(1); break;
    case 2: ME_LN_FWD(2); break;
    case 3: ME_LN_FWD(3); break;
    case 4: ME_LN_FWD(4); break;
    case 6: ME_LN_FWD(6); break;
    case 8: ME_LN_FWD(8); break;
    default: ME_LN_FWD(8); break;
  }
#undef ME_LN_FWD
  return 0;
}

static int ln_nch(int d) {
  const int n = (d + 127) / 128;
  if (n <= 4) return n;
  if (n <= 6) return 6;
  return 8;
}

int launch_add_ln_fwd(const float* x_res, const void* y, int dtype, const float* gamma, const float* beta,
                      float eps, int M, int d, float p, uint64_t seed, float* out_f32, void* out_T, float* z,
                      float* mean, float* rstd, const float* src_mean, const float* src_rstd, const float* src_gamma,
                      const float* src_beta, cudaStream_t st) {
  ME_CHECK(d % 4 == 0 && d <= 1024, "layernorm: d=%d must be a multiple of 4 and <= 1024", d);
  ME_CHECK(out_f32 != nullptr || (out_T != nullptr && dtype == ME_BF16), "layernorm: no output");
  const int wpb = LN_THREADS / 32;
  int blocks = (M + wpb - 1) / wpb;
  const int cap = sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (dtype == ME_BF16)
    ln_fwd_dispatch<bf16>(ln_nch(d), blocks, st, x_res, static_cast<const bf16*>(y), gamma, beta, eps, M, d, p, seed,
                          out_f32, static_cast<bf16*>(out_T), z, mean, rstd, src_mean, src_rstd, src_gamma, src_beta);
  else
    ln_fwd_dispatch<float>(ln_nch(d), blocks, st, x_res, static_cast<const float*>(y), gamma, beta, eps, M, d, p,
                           seed, out_f32, static_cast<float*>(out_T), z, mean, rstd, src_mean, src_rstd, src_gamma,
                           src_beta);
  ME_LAUNCH_CHECK();
  return 0;
}

template <typename T>
static int ln_bwd_dispatch(int nch, int blocks, size_t smem, cudaStream_t st, const float* dout,
                           const float* dout_add, const bf16* dout_add_T, const float* z, const float* mean,
                           const float* rstd,
                           const float* gamma, int M, int d, float p, uint64_t seed, float* dz_f32, T* dy_T,
                           float* d_gamma, float* d_beta, float* d_ybias) {
#define ME_LN_BWD(N)                                                                                             \
  do {                                                                                                           \
    static bool configured = false;                                                                              \
    if (!configured) {                                                                                           \
      ME_CUDA(cudaFuncSetAttribute(add_ln_bwd_kernel<T, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
      configured = true;                                                                                         \
    }                                                                                                            \
    add_ln_bwd_kernel<T, N><<<blocks, LN_THREADS, smem, st>>>(dout, dout_add, dout_add_T, z, mean, rstd, gamma, M, d, \
                                                              p, seed, dz_f32, dy_T, d_gamma, d_beta, d_ybias);   \
  } while (0)
  switch (nch) {
    case 1: ME_LN_BWD(1); break;
    case 2: ME_LN_BWD(2); break;
    case 3: ME_LN_BWD(3); break;
    case 4: ME_LN_BWD(4); break;
    case 6: ME_LN_BWD(6); break;
    case 8: ME_LN_BWD(8); break;
    default: ME_LN_BWD(8); break;
  }
#undef ME_LN_BWD
  return 0;
}

int launch_add_ln_bwd(const float* dout, const float* dout_add, const void* dout_add_T, const float* z,
                      const float* mean, const float* rstd, const float* gamma, int M, int d, float p, uint64_t seed,
                      int dtype, float* dz_f32, void* dy_T, float* d_gamma, float* d_beta, float* d_ybias,
                      cudaStream_t st) {
  ME_CHECK(d % 4 == 0 && d <= 1024, "layernorm backward: d=%d must be a multiple of 4 and <= 1024", d);
  const int wpb = LN_THREADS / 32;
  int blocks = (M + wpb - 1) / wpb;
  const int cap = sm_count() * 2;
  if (blocks > cap) blocks = cap;
  const size_t smem = (1 + 3 * (LN_THREADS / 32)) * static_cast<size_t>(d) * sizeof(float);
  if (dtype == ME_BF16)
    ln_bwd_dispatch<bf16>(ln_nch(d), blocks, smem, st, dout, dout_add, static_cast<const bf16*>(dout_add_T), z, mean, rstd,
                          gamma, M, d, p, seed, dz_f32, static_cast<bf16*>(dy_T), d_gamma, d_beta, d_ybias);
  else
    ln_bwd_dispatch<float>(ln_nch(d), blocks, smem, st, dout, dout_add, nullptr, z, mean, rstd, gamma, M, d, p, seed,
                           dz_f32, static_cast<float*>(dy_T), d_gamma, d_beta, d_ybias);
  ME_LAUNCH_CHECK();
  return 0;
}

// Row-streaming variant for bf16 matrices whose width is a multiple of 256 (the bias-gradient sums of the
// layer: [M, d_inner], [M, 3d]): like the LayerNorm kernels a warp reads whole rows -- NCH 16-byte loads per
// lane, issued together, two rows in flight -- and keeps its 8*NCH column sums in registers.  Same-address
// atomics from different SMs serialise at ~0.1 us each (measured: kernel time grew linearly with the number
// of adds per column), so the block folds its warps in shared memory and writes ONE partial row to a
// workspace; a second tiny kernel adds the partial rows into `out`.
template <int NCH>
__global__ void __launch_bounds__(256, 1)
colsum_rows_kernel(const bf16* __restrict__ X, int M, int N, int ldx, float* __restrict__ partial) {
  __shared__ float fold[NCH * 256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warp0 = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * 8;
  float acc[NCH][8];
#pragma unroll
  for (int k = 0; k < NCH; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[k][e] = 0.f;
  for (int64_t row = warp0; row < M; row += 2 * nwarps) {
    const bool two = row + nwarps < M;
    const bf16* p0 = X + row * ldx + lane * 8;
    const bf16* p1 = X + (two ? row + nwarps : row) * ldx + lane * 8;
    uint4 u0[NCH], u1[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      // chunks past the row pitch are redirected to chunk 0 (branch-free loads); their sums are never written
      const int off = (256 * k + lane * 8 + 8 <= ldx) ? 256 * k : -lane * 8;
      u0[k] = *reinterpret_cast<const uint4*>(p0 + off);
      u1[k] = *reinterpret_cast<const uint4*>(p1 + off);
    }
    const float w1 = two ? 1.f : 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const uint32_t a[4] = {u0[k].x, u0[k].y, u0[k].z, u0[k].w};
      const uint32_t b[4] = {u1[k].x, u1[k].y, u1[k].z, u1[k].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc[k][2 * e] += __uint_as_float(a[e] << 16) + w1 * __uint_as_float(b[e] << 16);
        acc[k][2 * e + 1] += __uint_as_float(a[e] & 0xFFFF0000u) + w1 * __uint_as_float(b[e] & 0xFFFF0000u);
      }
    }
  }
  for (int w = 0; w < 8; ++w) {   // fold the eight warps, one after the other
    if (warp == w) {
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        float4* f = reinterpret_cast<float4*>(fold + 256 * k + lane * 8);
        float4 lo = w ? f[0] : make_float4(0.f, 0.f, 0.f, 0.f), hi = w ? f[1] : make_float4(0.f, 0.f, 0.f, 0.f);
        lo.x += acc[k][0]; lo.y += acc[k][1]; lo.z += acc[k][2]; lo.w += acc[k][3];
        hi.x += acc[k][4]; hi.y += acc[k][5]; hi.z += acc[k][6]; hi.w += acc[k][7];
        f[0] = lo; f[1] = hi;
      }
    }
    __syncthreads();
  }
  float* prow = partial + static_cast<int64_t>(blockIdx.x) * N;
  for (int c = threadIdx.x; c < N; c += 256) prow[c] = fold[c];
}
// out[n] += sum of the partial rows; blockIdx.y splits the rows (a handful of adds per column, not hundreds)
__global__ void colsum_finish_kernel(const float* __restrict__ partial, int rows, int N, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int per = (rows + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float s = 0.f;
#pragma unroll 8
  for (int r = r0; r < r1; ++r) s += partial[static_cast<int64_t>(r) * N + n];
  if (r1 > r0) atomicAdd(&out[n], s);
}

int launch_colsum(const void* X, int dtype, int M, int N, int ldx, float* out, cudaStream_t st) {
  const int threads = 64;
  const int vec = dtype == ME_BF16 ? 8 : 4;
  const int gx = ((N + vec - 1) / vec + threads - 1) / threads;
  int gy = max(1, (sm_count() * 16) / gx);
  int rpb = (M + gy - 1) / gy;
  if (rpb < 16) rpb = 16;
  gy = (M + rpb - 1) / rpb;
  dim3 grid(gx, gy);
  if (dtype == ME_BF16)
    colsum_kernel<bf16><<<grid, threads, 0, st>>>(static_cast<const bf16*>(X), M, N, ldx, rpb, out);
  else
    colsum_kernel<float><<<grid, threads, 0, st>>>(static_cast<const float*>(X), M, N, ldx, rpb, out);
  ME_LAUNCH_CHECK();
  return 0;
}

// out[n] += column sums; `ws` (>= sm_count * N floats, 16-byte aligned) is optional scratch for the partial rows.
int launch_colsum_ws(const void* X, int dtype, int M, int N, int ldx, float* out, float* ws, int64_t ws_floats,
                     cudaStream_t st) {
  const int blocks = sm_count();
  if (ws && ws_floats >= static_cast<int64_t>(blocks) * N && dtype == ME_BF16 && N <= 3072 && ldx % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(ws) & 15) == 0 && M >= 2048) {
    const bf16* x = static_cast<const bf16*>(X);
    switch ((N + 255) / 256) {
#define ME_CS(NC) case NC: colsum_rows_kernel<NC><<<blocks, 256, 0, st>>>(x, M, N, ldx, ws); break;
      ME_CS(1) ME_CS(2) ME_CS(3) ME_CS(4) ME_CS(5) ME_CS(6) ME_CS(7) ME_CS(8) ME_CS(9) ME_CS(10) ME_CS(11) ME_CS(12)
#undef ME_CS
    }
    ME_LAUNCH_CHECK();
    colsum_finish_kernel<<<dim3((N + 255) / 256, 8), 256, 0, st>>>(ws, blocks, N, out);
    ME_LAUNCH_CHECK();
    return 0;
  }
  return launch_colsum(X, dtype, M, N, ldx, out, st);
}


}  // namespace me

using namespace me;

extern "C" int me_embed_forward(const int64_t* tokens, const float* cond, const float* emb_w, const float* cw0,
                                const float* cb0, const float* cw1, const float* cb1, const float* pe, int B,
                                int L, int d, int d_cond, int V, int mode, int pad_token, float dropout_p,
                                uint64_t seed, int dtype, float* x_f32, void* x_T, uint8_t* keypad, void* stream) {
  ME_CHECK(B > 0 && L > 0 && d > 0 && d_cond >= 0 && d_cond < d, "me_embed_forward: bad dims");
  ME_CHECK(mode >= 0 && mode <= 3, "me_embed_forward: bad mode %d", mode);
  ME_CHECK(mode == ME_COND_CONTINUOUS_CONCAT || d_cond == 0, "me_embed_forward: d_cond>0 needs concat mode");
  const int Ls = mode == ME_COND_CONTINUOUS_TOKEN ? L + 2 : L;
  ME_CHECK(Ls <= 2048, "me_embed_forward: sequence length %d exceeds max_seq", Ls);
  return launch_embed(tokens, cond, emb_w, cw0, cb0, cw1, cb1, pe, B, L, d, d_cond, V, mode, pad_token,
                      dropout_p, seed, nullptr, 0, dtype, x_f32, x_T, keypad, static_cast<cudaStream_t>(stream));
}

extern "C" int me_embed_decode(const int64_t* tokens, const float* cond, const float* emb_w, const float* cw0,
                               const float* cb0, const float* pe, int B, int d, int d_cond, int V, int mode,
                               int pad_token, const int32_t* t_dev, int dtype, float* x_f32, void* x_T,
                               uint8_t* keypad, int T_max, void* stream) {
  ME_CHECK(t_dev != nullptr, "me_embed_decode: t_dev is NULL");
  return launch_embed(tokens, cond, emb_w, cw0, cb0, nullptr, nullptr, pe, B, 1, d, d_cond, V, mode, pad_token,
                      0.f, 0, t_dev, T_max, dtype, x_f32, x_T, keypad, static_cast<cudaStream_t>(stream));
}

extern "C" int me_embed_backward_split(const float* dx, const void* dx_T, const int64_t* tokens, const float* cond,
                                       int B, int L, int d, int d_cond, int V, int mode, int pad_token,
                                       float dropout_p, uint64_t seed, float* d_emb, float* d_cw0, float* d_cb0,
                                       float* d_cw1, float* d_cb1, void* stream) {
  ME_CHECK(dx_T == nullptr || ((reinterpret_cast<uintptr_t>(dx_T) & 7) == 0 && d % 4 == 0),
           "me_embed_backward_split: dx_T needs 8-byte alignment and d a multiple of 4");
  const int Ls = mode == ME_COND_CONTINUOUS_TOKEN ? L + 2 : L;
  const int64_t rows = static_cast<int64_t>(B) * Ls;
  embed_bwd_kernel<<<static_cast<int>((rows + EB_ROWS - 1) / EB_ROWS), EB_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      dx, static_cast<const bf16*>(dx_T), tokens, cond, B, L, d, d_cond, V, mode, pad_token, dropout_p, seed, d_emb,
      d_cw0, d_cb0, d_cw1, d_cb1);
  ME_LAUNCH_CHECK();
  return 0;
}

extern "C" int me_embed_backward(const float* dx, const int64_t* tokens, const float* cond, int B, int L, int d,
                                 int d_cond, int V, int mode, int pad_token, float dropout_p, uint64_t seed,
                                 float* d_emb, float* d_cw0, float* d_cb0, float* d_cw1, float* d_cb1,
                                 void* stream) {
  return me_embed_backward_split(dx, nullptr, tokens, cond, B, L, d, d_cond, V, mode, pad_token, dropout_p, seed, d_emb,
                                 d_cw0, d_cb0, d_cw1, d_cb1, stream);
}

extern "C" int me_add_layernorm_forward(const float* x_res, const void* y, int dtype, const float* gamma,
                                        const float* beta, float eps, int M, int d, float dropout_p,
                                        uint64_t seed, float* out_f32, void* out_T, float* z, float* mean,
                                        float* rstd, void* stream) {
  ME_CHECK(M > 0 && d > 0 && d <= 1024, "me_add_layernorm_forward: bad dims M=%d d=%d", M, d);
  return launch_add_ln_fwd(x_res, y, dtype, gamma, beta, eps, M, d, dropout_p, seed, out_f32, out_T, z, mean,
                           rstd, nullptr, nullptr, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

extern "C" int me_add_layernorm_backward(const float* dout, const float* dout_add, const float* z,
                                         const float* mean, const float* rstd, const float* gamma, int M, int d,
                                         float dropout_p, uint64_t seed, int dtype, float* dz_f32, void* dy_T,
                                         float* d_gamma, float* d_beta, void* stream) {
  ME_CHECK(M > 0 && d > 0 && d <= 1024, "me_add_layernorm_backward: bad dims");
  return launch_add_ln_bwd(dout, dout_add, nullptr, z, mean, rstd, gamma, M, d, dropout_p, seed, dtype, dz_f32, dy_T,
                           d_gamma, d_beta, nullptr, static_cast<cudaStream_t>(stream));
}

extern "C" int me_colsum_ws(const void* X, int dtype, int M, int N, int ldx, float* out, float* ws, int64_t ws_floats,
                            void* stream) {
  return launch_colsum_ws(X, dtype, M, N, ldx, out, ws, ws_floats, static_cast<cudaStream_t>(stream));
}

extern "C" int me_colsum(const void* X, int dtype, int M, int N, int ldx, float* out, void* stream) {
  return launch_colsum(X, dtype, M, N, ldx, out, static_cast<cudaStream_t>(stream));
}

extern "C" int me_convert_2d(const void* src, int src_dtype, int ld_src, void* dst, int dst_dtype, int ld_dst,
                             int rows, int cols, void* stream) {
  const int64_t total = static_cast<int64_t>(rows) * ld_dst;
  const int64_t want = (total + 255) / 256;
  const int blocks = static_cast<int>(want < 148 * 16 ? want : 148 * 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (src_dtype == ME_F32 && dst_dtype == ME_BF16)
    convert2d_kernel<float, bf16><<<blocks, 256, 0, st>>>(static_cast<const float*>(src), ld_src,
                                                          static_cast<bf16*>(dst), ld_dst, rows, cols);
  else if (src_dtype == ME_BF16 && dst_dtype == ME_F32)
    convert2d_kernel<bf16, float><<<blocks, 256, 0, st>>>(static_cast<const bf16*>(src), ld_src,
                                                          static_cast<float*>(dst), ld_dst, rows, cols);
  else if (src_dtype == ME_F32 && dst_dtype == ME_F32)
    convert2d_kernel<float, float><<<blocks, 256, 0, st>>>(static_cast<const float*>(src), ld_src,
                                                           static_cast<float*>(dst), ld_dst, rows, cols);
  else
    convert2d_kernel<bf16, bf16><<<blocks, 256, 0, st>>>(static_cast<const bf16*>(src), ld_src,
                                                         static_cast<bf16*>(dst), ld_dst, rows, cols);
  ME_LAUNCH_CHECK();
  return 0;
}

// One launch for a whole table of 2-D copies/casts (the per-step refresh of the compute-type weight copies).
// blockIdx.y = table entry, blockIdx.x strides over its 4-element groups.
template <typename TS, typename TD>
__device__ __forceinline__ void convert_entry(const me_convert_desc& e) {
  const TS* src = static_cast<const TS*>(e.src);
  TD* dst = static_cast<TD*>(e.dst);
  const bool vec = (e.cols % 4 == 0) && (e.ld_src % 4 == 0) && (e.ld_dst % 4 == 0) && e.cols == e.ld_dst &&
                   ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  if (vec) {
    const int q = e.cols / 4;
    const int64_t total = static_cast<int64_t>(e.rows) * q;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
      const int64_t r = i / q;
      const int c = static_cast<int>(i - r * q) * 4;
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = to_f32<TS>(src[r * e.ld_src + c + k]);
#pragma unroll
      for (int k = 0; k < 4; ++k) dst[r * e.ld_dst + c + k] = from_f32<TD>(v[k]);
    }
  } else {
    const int64_t total = static_cast<int64_t>(e.rows) * e.ld_dst;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
      const int64_t r = i / e.ld_dst;
      const int c = static_cast<int>(i - r * e.ld_dst);
      dst[i] = from_f32<TD>(c < e.cols ? to_f32<TS>(src[r * e.ld_src + c]) : 0.f);
    }
  }
}
__global__ void convert_batched_kernel(const me_convert_desc* __restrict__ table) {
  const me_convert_desc e = table[blockIdx.y];
  if (e.src_dtype == ME_F32 && e.dst_dtype == ME_BF16) convert_entry<float, bf16>(e);
  else if (e.src_dtype == ME_F32) convert_entry<float, float>(e);
  else if (e.dst_dtype == ME_F32) convert_entry<bf16, float>(e);
  else convert_entry<bf16, bf16>(e);
}

extern "C" int me_convert_batched(const me_convert_desc* table_dev, int n, void* stream) {
  ME_CHECK(table_dev != nullptr && n > 0 && n <= 65535, "me_convert_batched: bad table (n = %d)", n);
  convert_batched_kernel<<<dim3(sm_count() * 2, n), 256, 0, static_cast<cudaStream_t>(stream)>>>(table_dev);
  ME_LAUNCH_CHECK();
  return 0;
}

extern "C" int me_kv_cache_write(const void* qkv, int dtype, int B, int Ls, int H, int dh, void* k_cache,
                                 void* v_cache, int T_max, int pos0, void* stream) {
  ME_CHECK(pos0 >= 0 && pos0 + Ls <= T_max, "me_kv_cache_write: positions %d..%d exceed T_max %d", pos0,
           pos0 + Ls, T_max);
  return launch_kv_write(qkv, dtype, B, Ls, H, dh, k_cache, v_cache, T_max, pos0, nullptr,
                         static_cast<cudaStream_t>(stream));
}
