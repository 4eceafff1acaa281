// Relative global attention on CUDA cores, fp32 arithmetic, any storage dtype.
//   S[i,j] = (q_i.k_j + q_i.E[max_seq-1-(i-j)]) / sqrt(dh)   for j <= i and key j not pad
// (ME_ATTN_NONCAUSAL: every key j < Lk is visible and the E term exists for j <= i only -- the
//  regression side model, models/music_regression.py:78,256-262)
// This is (a) the fp32 parity path, (b) the decode-step attention over the KV cache (one query
// row per sequence, HBM-bound: it streams the cache once), (c) the correctness reference that the
// tensor-core attention kernels are tested against on the device.
// One warp per query row (forward, dQ/dE) or per key row (dK/dV); online softmax; warp shuffles.
#include "common.cuh"
#include "../../include/midi_emotion_b200.h"

namespace me {

constexpr int AT_WARPS = 4;
constexpr int AT_MAX_DH = 128;

template <typename T>
__device__ __forceinline__ float dot_row(const float* __restrict__ qs, const T* __restrict__ row, int dh) {
  float acc = 0.f;
  for (int c = 0; c < dh; ++c) acc = fmaf(qs[c], to_f32<T>(row[c]), acc);
  return acc;
}
template <>
__device__ __forceinline__ float dot_row<bf16>(const float* __restrict__ qs, const bf16* __restrict__ row, int dh) {
  float acc = 0.f;
  const uint4* r4 = reinterpret_cast<const uint4*>(row);  // dh % 8 == 0 and 16-byte aligned rows
  for (int c = 0; c < dh / 8; ++c) {
    const uint4 u = r4[c];
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __bfloat1622float2(h[e]);
      acc = fmaf(qs[c * 8 + 2 * e], f.x, acc);
      acc = fmaf(qs[c * 8 + 2 * e + 1], f.y, acc);
    }
  }
  return acc;
}

struct AttnP {
  int B, H, Lq, Lk, dh, max_seq, q_pos0, noncausal;
  int64_t q_sb, q_sh, q_si, k_sb, k_sh, k_sj, v_sb, v_sh, v_sj, o_sb, o_si, keypad_ld;
  const int32_t* pos_dev;
};

// ------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(AT_WARPS * 32)
attn_fwd_simt(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, const T* __restrict__ E,
              const uint8_t* __restrict__ keypad, T* __restrict__ out, float* __restrict__ lse, AttnP p) {
  __shared__ float qs_all[AT_WARPS][AT_MAX_DH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int iq = blockIdx.x * AT_WARPS + warp;
  const int h = blockIdx.y, b = blockIdx.z;
  if (iq >= p.Lq) return;
  const int pos0 = p.pos_dev ? *p.pos_dev : p.q_pos0;
  const int i = pos0 + iq;  // absolute position of this query
  const int dh = p.dh;
  float* qs = qs_all[warp];
  const T* qrow = q + b * p.q_sb + h * p.q_sh + iq * p.q_si;
  for (int c = lane; c < dh; c += 32) qs[c] = to_f32<T>(qrow[c]);
  __syncwarp();
  const float sqrt_dh = sqrtf(static_cast<float>(dh));
  const T* kb = k + b * p.k_sb + h * p.k_sh;
  const T* vb = v + b * p.v_sb + h * p.v_sh;
  const uint8_t* kp = keypad ? keypad + b * p.keypad_ld : nullptr;

  float m = -INFINITY, l = 0.f;
  float acc[AT_MAX_DH / 32];
#pragma unroll
  for (int c = 0; c < AT_MAX_DH / 32; ++c) acc[c] = 0.f;

  const int jend = p.noncausal ? (p.pos_dev ? i + 1 : p.Lk) : i + 1;  // keys 0 .. jend-1 can be visible
  for (int j0 = 0; j0 < jend; j0 += 32) {
    const int j = j0 + lane;
    const bool valid = (j < jend) && !(kp && kp[j]);
    float s = -INFINITY;
    if (valid) {
      const float qk = dot_row<T>(qs, kb + j * p.k_sj, dh);
      const float qe = j <= i ? dot_row<T>(qs, E + static_cast<int64_t>(p.max_seq - 1 - (i - j)) * dh, dh) : 0.f;
      s = (qk + qe) / sqrt_dh;
    }
    const float cm = warp_max(s);
    if (cm == -INFINITY) continue;  // whole chunk masked
    const float m_new = fmaxf(m, cm);
    const float alpha = (m == -INFINITY) ? 0.f : expf(m - m_new);
    const float pj = valid ? expf(s - m_new) : 0.f;
    l = l * alpha + warp_sum(pj);
#pragma unroll
    for (int c = 0; c < AT_MAX_DH / 32; ++c) acc[c] *= alpha;
    const int jn = min(32, jend - j0);
    for (int jj = 0; jj < jn; ++jj) {
      const float pb = __shfl_sync(0xffffffffu, pj, jj);
      if (pb != 0.f) {
        const T* vrow = vb + (j0 + jj) * p.v_sj;
#pragma unroll
        for (int c = 0; c < AT_MAX_DH / 32; ++c) {
          const int e = lane + 32 * c;
          if (e < dh) acc[c] = fmaf(pb, to_f32<T>(vrow[e]), acc[c]);
        }
      }
    }
    m = m_new;
  }
  const float inv_l = l > 0.f ? 1.f / l : 0.f;  // fully-masked rows give 0 (reference: NaN)
  T* orow = out + b * p.o_sb + iq * p.o_si + h * dh;
#pragma unroll
  for (int c = 0; c < AT_MAX_DH / 32; ++c) {
    const int e = lane + 32 * c;
    if (e < dh) orow[e] = from_f32<T>(acc[c] * inv_l);
  }
  if (lse && lane == 0) lse[(static_cast<int64_t>(b) * p.H + h) * p.Lq + iq] = l > 0.f ? m + logf(l) : -INFINITY;
}

// ------------------------------------------------------------------------------------
// backward, pass A: per query row -> dq, dE (atomics), dsum = rowsum(dO * O)
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(AT_WARPS * 32)
attn_bwd_dq_simt(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                 const T* __restrict__ E, const uint8_t* __restrict__ keypad, const T* __restrict__ out,
                 const T* __restrict__ dout, const float* __restrict__ lse, T* __restrict__ dq,
                 float* __restrict__ dE, float* __restrict__ dsum, AttnP p) {
  __shared__ float qs_all[AT_WARPS][AT_MAX_DH];
  __shared__ float dos_all[AT_WARPS][AT_MAX_DH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * AT_WARPS + warp;
  const int h = blockIdx.y, b = blockIdx.z;
  if (i >= p.Lq) return;
  const int dh = p.dh;
  float* qs = qs_all[warp];
  float* dos = dos_all[warp];
  const T* qrow = q + b * p.q_sb + h * p.q_sh + i * p.q_si;
  const T* orow = out + b * p.o_sb + i * p.o_si + h * dh;
  const T* dorow = dout + b * p.o_sb + i * p.o_si + h * dh;
  float dpart = 0.f;
  for (int c = lane; c < dh; c += 32) {
    qs[c] = to_f32<T>(qrow[c]);
    const float g = to_f32<T>(dorow[c]);
    dos[c] = g;
    dpart += g * to_f32<T>(orow[c]);
  }
  const float Di = warp_sum(dpart);
  __syncwarp();
  const int64_t row_id = (static_cast<int64_t>(b) * p.H + h) * p.Lq + i;
  if (lane == 0) dsum[row_id] = Di;
  const float lse_i = lse[row_id];
  const float sqrt_dh = sqrtf(static_cast<float>(dh));
  const T* kb = k + b * p.k_sb + h * p.k_sh;
  const T* vb = v + b * p.v_sb + h * p.v_sh;
  const uint8_t* kp = keypad ? keypad + b * p.keypad_ld : nullptr;
  float acc[AT_MAX_DH / 32];
#pragma unroll
  for (int c = 0; c < AT_MAX_DH / 32; ++c) acc[c] = 0.f;

  const int jend = p.noncausal ? p.Lk : i + 1;
  for (int j0 = 0; j0 < jend; j0 += 32) {
    const int j = j0 + lane;
    const bool valid = (j < jend) && !(kp && kp[j]) && lse_i != -INFINITY;
    float ds = 0.f;
    if (valid) {
      const float qk = dot_row<T>(qs, kb + j * p.k_sj, dh);
      const float qe = j <= i ? dot_row<T>(qs, E + static_cast<int64_t>(p.max_seq - 1 - (i - j)) * dh, dh) : 0.f;
      const float s = (qk + qe) / sqrt_dh;
      const float pj = expf(s - lse_i);
      const float dp = dot_row<T>(dos, vb + j * p.v_sj, dh);
      ds = pj * (dp - Di) / sqrt_dh;
    }
    const int jn = min(32, jend - j0);
    for (int jj = 0; jj < jn; ++jj) {
      const float dsb = __shfl_sync(0xffffffffu, ds, jj);
      if (dsb != 0.f) {
        const int jx = j0 + jj;
        const T* krow = kb + jx * p.k_sj;
        if (jx <= i) {
          const int64_t eidx = static_cast<int64_t>(p.max_seq - 1 - (i - jx)) * dh;
#pragma unroll
          for (int c = 0; c < AT_MAX_DH / 32; ++c) {
            const int e = lane + 32 * c;
            if (e < dh) {
              acc[c] = fmaf(dsb, to_f32<T>(krow[e]) + to_f32<T>(E[eidx + e]), acc[c]);
              atomicAdd(&dE[eidx + e], dsb * qs[e]);
            }
          }
        } else {  // above the diagonal (non-causal only): no relative term
#pragma unroll
          for (int c = 0; c < AT_MAX_DH / 32; ++c) {
            const int e = lane + 32 * c;
            if (e < dh) acc[c] = fmaf(dsb, to_f32<T>(krow[e]), acc[c]);
          }
        }
      }
    }
  }
  T* dqrow = dq + b * p.q_sb + h * p.q_sh + i * p.q_si;
#pragma unroll
  for (int c = 0; c < AT_MAX_DH / 32; ++c) {
    const int e = lane + 32 * c;
    if (e < dh) dqrow[e] = from_f32<T>(acc[c]);
  }
}

// ------------------------------------------------------------------------------------
// backward, pass B: per key row -> dk, dv
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(AT_WARPS * 32)
attn_bwd_dkv_simt(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                  const T* __restrict__ E, const uint8_t* __restrict__ keypad, const T* __restrict__ dout,
                  const float* __restrict__ lse, const float* __restrict__ dsum, T* __restrict__ dk,
                  T* __restrict__ dv, AttnP p) {
  __shared__ float ks_all[AT_WARPS][AT_MAX_DH];
  __shared__ float vs_all[AT_WARPS][AT_MAX_DH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * AT_WARPS + warp;
  const int h = blockIdx.y, b = blockIdx.z;
  if (j >= p.Lk) return;
  const int dh = p.dh;
  float* ks = ks_all[warp];
  float* vs = vs_all[warp];
  const T* krow = k + b * p.k_sb + h * p.k_sh + j * p.k_sj;
  const T* vrow = v + b * p.v_sb + h * p.v_sh + j * p.v_sj;
  for (int c = lane; c < dh; c += 32) {
    ks[c] = to_f32<T>(krow[c]);
    vs[c] = to_f32<T>(vrow[c]);
  }
  __syncwarp();
  const bool key_masked = keypad && keypad[b * p.keypad_ld + j];
  const float sqrt_dh = sqrtf(static_cast<float>(dh));
  const T* qb = q + b * p.q_sb + h * p.q_sh;
  const T* dob = dout + b * p.o_sb + h * dh;
  float acck[AT_MAX_DH / 32], accv[AT_MAX_DH / 32];
#pragma unroll
  for (int c = 0; c < AT_MAX_DH / 32; ++c) acck[c] = accv[c] = 0.f;

  if (!key_masked) {
    for (int i0 = p.noncausal ? 0 : j; i0 < p.Lq; i0 += 32) {
      const int i = i0 + lane;
      float pj = 0.f, ds = 0.f;
      if (i < p.Lq) {
        const int64_t row_id = (static_cast<int64_t>(b) * p.H + h) * p.Lq + i;
        const float lse_i = lse[row_id];
        if (lse_i != -INFINITY) {
          const T* qrow = qb + i * p.q_si;
          // q_i . k_j and q_i . E[idx]: dot_row wants the fp32 vector first
          float qk = 0.f, qe = 0.f, dp = 0.f;
          const bool rel = i >= j;  // the relative term is lower-triangular
          const T* erow = E + static_cast<int64_t>(p.max_seq - 1 - (rel ? i - j : 0)) * dh;
          const T* dorow = dob + b * 0 + i * p.o_si;
          for (int c = 0; c < dh; ++c) {
            const float qv = to_f32<T>(qrow[c]);
            qk = fmaf(qv, ks[c], qk);
            qe = fmaf(qv, to_f32<T>(erow[c]), qe);
            dp = fmaf(to_f32<T>(dorow[c]), vs[c], dp);
          }
          if (!rel) qe = 0.f;
          const float s = (qk + qe) / sqrt_dh;
          pj = expf(s - lse_i);
          ds = pj * (dp - dsum[row_id]) / sqrt_dh;
        }
      }
      const int in = min(32, p.Lq - i0);
      for (int ii = 0; ii < in; ++ii) {
        const float pb = __shfl_sync(0xffffffffu, pj, ii);
        const float dsb = __shfl_sync(0xffffffffu, ds, ii);
        if (pb != 0.f || dsb != 0.f) {
          const T* qrow = qb + (i0 + ii) * p.q_si;
          const T* dorow = dob + (i0 + ii) * p.o_si;
#pragma unroll
          for (int c = 0; c < AT_MAX_DH / 32; ++c) {
            const int e = lane + 32 * c;
            if (e < dh) {
              accv[c] = fmaf(pb, to_f32<T>(dorow[e]), accv[c]);
              acck[c] = fmaf(dsb, to_f32<T>(qrow[e]), acck[c]);
            }
          }
        }
      }
    }
  }
  T* dkrow = dk + b * p.k_sb + h * p.k_sh + j * p.k_sj;
  T* dvrow = dv + b * p.v_sb + h * p.v_sh + j * p.v_sj;
#pragma unroll
  for (int c = 0; c < AT_MAX_DH / 32; ++c) {
    const int e = lane + 32 * c;
    if (e < dh) {
      dkrow[e] = from_f32<T>(acck[c]);
      dvrow[e] = from_f32<T>(accv[c]);
    }
  }
}

static AttnP to_p(const me_attn_args* a) {
  AttnP p;
  p.B = a->B; p.H = a->H; p.Lq = a->Lq; p.Lk = a->Lk; p.dh = a->dh; p.max_seq = a->max_seq; p.q_pos0 = a->q_pos0;
  p.q_sb = a->q_sb; p.q_sh = a->q_sh; p.q_si = a->q_si;
  p.k_sb = a->k_sb; p.k_sh = a->k_sh; p.k_sj = a->k_sj;
  p.v_sb = a->v_sb; p.v_sh = a->v_sh; p.v_sj = a->v_sj;
  p.o_sb = a->o_sb; p.o_si = a->o_si; p.keypad_ld = a->keypad_ld; p.pos_dev = a->pos_dev;
  p.noncausal = (a->flags & ME_ATTN_NONCAUSAL) ? 1 : 0;
  return p;
}

int attn_check(const me_attn_args* a, const char* who) {
  ME_CHECK(a->B > 0 && a->H > 0 && a->Lq > 0 && a->dh > 0, "%s: bad dims", who);
  ME_CHECK(a->dh <= AT_MAX_DH && a->dh % 8 == 0, "%s: head dim %d unsupported (<=128, multiple of 8)", who, a->dh);
  ME_CHECK(a->pos_dev || (a->q_pos0 + a->Lq <= a->max_seq), "%s: positions exceed max_seq", who);
  ME_CHECK(a->dtype == ME_F32 || a->dtype == ME_BF16, "%s: bad dtype", who);
  if (a->dtype == ME_BF16) {
    ME_CHECK(a->k_sj % 8 == 0 && a->k_sh % 8 == 0 && a->k_sb % 8 == 0 && a->v_sj % 8 == 0,
             "%s: bf16 rows must be 16-byte aligned", who);
  }
  return 0;
}

int launch_attn_decode(const me_attn_args* a);

int launch_attn_fwd_simt(const me_attn_args* a) {
  if (attn_check(a, "me_attention_forward")) return 1;
  // single query row per sequence over a bf16 KV cache: the HBM-streaming decode kernel
  if (a->dtype == ME_BF16 && a->Lq == 1 && a->lse == nullptr && (a->dh == 32 || a->dh == 48 || a->dh == 64) &&
      a->q_sb % 8 == 0 && a->q_sh % 8 == 0 && a->o_sb % 2 == 0)
    return launch_attn_decode(a);
  AttnP p = to_p(a);
  dim3 grid((a->Lq + AT_WARPS - 1) / AT_WARPS, a->H, a->B);
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  if (a->dtype == ME_BF16)
    attn_fwd_simt<bf16><<<grid, AT_WARPS * 32, 0, st>>>(
        static_cast<const bf16*>(a->q), static_cast<const bf16*>(a->k), static_cast<const bf16*>(a->v),
        static_cast<const bf16*>(a->E), a->keypad, static_cast<bf16*>(a->out), a->lse, p);
  else
    attn_fwd_simt<float><<<grid, AT_WARPS * 32, 0, st>>>(
        static_cast<const float*>(a->q), static_cast<const float*>(a->k), static_cast<const float*>(a->v),
        static_cast<const float*>(a->E), a->keypad, static_cast<float*>(a->out), a->lse, p);
  ME_LAUNCH_CHECK();
  return 0;
}

int launch_attn_bwd_simt(const me_attn_bwd_args* ba) {
  const me_attn_args* a = &ba->f;
  if (attn_check(a, "me_attention_backward")) return 1;
  ME_CHECK(a->q_pos0 == 0 && a->Lq == a->Lk && a->pos_dev == nullptr, "me_attention_backward: self-attention only");
  ME_CHECK(a->lse && ba->dsum && ba->dE, "me_attention_backward: lse/dsum/dE required");
  AttnP p = to_p(a);
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  dim3 gq((a->Lq + AT_WARPS - 1) / AT_WARPS, a->H, a->B);
  dim3 gk((a->Lk + AT_WARPS - 1) / AT_WARPS, a->H, a->B);
  if (a->dtype == ME_BF16) {
    attn_bwd_dq_simt<bf16><<<gq, AT_WARPS * 32, 0, st>>>(
        static_cast<const bf16*>(a->q), static_cast<const bf16*>(a->k), static_cast<const bf16*>(a->v),
        static_cast<const bf16*>(a->E), a->keypad, static_cast<const bf16*>(a->out),
        static_cast<const bf16*>(ba->dout), a->lse, static_cast<bf16*>(ba->dq), ba->dE, ba->dsum, p);
    attn_bwd_dkv_simt<bf16><<<gk, AT_WARPS * 32, 0, st>>>(
        static_cast<const bf16*>(a->q), static_cast<const bf16*>(a->k), static_cast<const bf16*>(a->v),
        static_cast<const bf16*>(a->E), a->keypad, static_cast<const bf16*>(ba->dout), a->lse, ba->dsum,
        static_cast<bf16*>(ba->dk), static_cast<bf16*>(ba->dv), p);
  } else {
    attn_bwd_dq_simt<float><<<gq, AT_WARPS * 32, 0, st>>>(
        static_cast<const float*>(a->q), static_cast<const float*>(a->k), static_cast<const float*>(a->v),
        static_cast<const float*>(a->E), a->keypad, static_cast<const float*>(a->out),
        static_cast<const float*>(ba->dout), a->lse, static_cast<float*>(ba->dq), ba->dE, ba->dsum, p);
    attn_bwd_dkv_simt<float><<<gk, AT_WARPS * 32, 0, st>>>(
        static_cast<const float*>(a->q), static_cast<const float*>(a->k), static_cast<const float*>(a->v),
        static_cast<const float*>(a->E), a->keypad, static_cast<const float*>(ba->dout), a->lse, ba->dsum,
        static_cast<float*>(ba->dk), static_cast<float*>(ba->dv), p);
  }
  ++g_launch_count;
  ME_LAUNCH_CHECK();
  return 0;
}

}  // namespace me
