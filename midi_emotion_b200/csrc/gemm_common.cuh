// Pieces shared by the 1-CTA and the 2-CTA tcgen05 GEMM kernels.
#pragma once
#include "common.cuh"
#include "../../include/midi_emotion_b200.h"

namespace me {

cudaEvent_t prof_begin(double flops, cudaStream_t st, int cls = 0);   // (declared again in attention_tc.cuh)
void prof_end(cudaEvent_t e, cudaStream_t st);

struct GemmParams {
  int M, N, K;
  int ldd, ldmask;
  int flags, out_dtype;
  int num_m_tiles, num_n_tiles, splits, kb_per_split, num_kb;
  const float* bias;
  const float* addend;
  const void* relu_mask;
  void* D;
  float* colsum;   // staged bf16 epilogue of the pair kernel: colsum[n] += sum over rows of the stored output, or NULL
};


// Epilogue arithmetic on one chunk of CW accumulator columns (already loaded from TMEM into r):
// bias / residual / ReLU / ReLU-mask.  addv / maskw were fetched while the TMEM load was in flight.
template <int CW>
__device__ __forceinline__ void gemm_epilogue_math(const GemmParams& p, const uint32_t (&r)[CW], const float* bs,
                                                   const float (&addv)[CW], const uint32_t (&maskw)[CW / 2],
                                                   bool use_bias, bool do_add, bool do_mask, float (&v)[CW]) {
#pragma unroll
  for (int j = 0; j < CW; ++j) v[j] = __uint_as_float(r[j]);
  if (use_bias) {
#pragma unroll
    for (int j = 0; j < CW; ++j) v[j] += bs[j];
  }
  if (do_add) {
#pragma unroll
    for (int j = 0; j < CW; ++j) v[j] += addv[j];
  }
  if (p.flags & ME_EPI_RELU) {
#pragma unroll
    for (int j = 0; j < CW; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (do_mask) {
#pragma unroll
    for (int j = 0; j < CW / 2; ++j) {
      const uint32_t w = maskw[j];  // bf16 > 0  <=>  sign bit clear and magnitude bits non-zero
      if (!((w & 0x8000u) == 0u && (w & 0x7FFFu) != 0u)) v[2 * j] = 0.f;
      if (!((w & 0x80000000u) == 0u && (w & 0x7FFF0000u) != 0u)) v[2 * j + 1] = 0.f;
    }
  }
}

// Direct stores of one chunk of row m: bf16 or fp32 (16-byte vectors when the row pitch allows) or split-K atomics.
template <int CW>
__device__ __forceinline__ void gemm_epilogue_chunk(const GemmParams& p, const uint32_t (&r)[CW], const float* bs,
                                                    const float (&addv)[CW], const uint32_t (&maskw)[CW / 2],
                                                    bool use_bias, bool do_add, bool do_mask, int m, int nb,
                                                    bool full, bool vec_ok) {
  float v[CW];
  gemm_epilogue_math<CW>(p, r, bs, addv, maskw, use_bias, do_add, do_mask, v);
  if (p.splits > 1) {
    float* dp = static_cast<float*>(p.D) + static_cast<int64_t>(m) * p.ldd + nb;
    if (full && (p.ldd % 4 == 0)) {
      // 16-byte vector reductions: a quarter of the instructions of scalar atomics (the split-K epilogue of the
      // weight-gradient GEMMs spent ~20 us per launch issuing 32768 scalar atomics per CTA)
#pragma unroll
      for (int j = 0; j < CW; j += 4)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dp + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]),
                     "f"(v[j + 3])
                     : "memory");
    } else {
#pragma unroll
      for (int j = 0; j < CW; ++j)
        if (nb + j < p.N) atomicAdd(dp + j, v[j]);
    }
  } else if (p.out_dtype == ME_BF16) {
    bf16* dp = static_cast<bf16*>(p.D) + static_cast<int64_t>(m) * p.ldd + nb;
    if (vec_ok && full) {
#pragma unroll
      for (int j = 0; j < CW / 8; ++j) {
        uint4 u;
        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
        reinterpret_cast<uint4*>(dp)[j] = u;
      }
    } else {
#pragma unroll
      for (int j = 0; j < CW; ++j)
        if (nb + j < p.N) dp[j] = __float2bfloat16_rn(v[j]);
    }
  } else {
    float* dp = static_cast<float*>(p.D) + static_cast<int64_t>(m) * p.ldd + nb;
    if (vec_ok && full) {
#pragma unroll
      for (int j = 0; j < CW / 4; ++j)
        reinterpret_cast<float4*>(dp)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < CW; ++j)
        if (nb + j < p.N) dp[j] = v[j];
    }
  }
}

// fp32 output of one 32-column chunk with full-sector global traffic.  The accumulator layout is one row
// per lane; stored that way a warp-wide 16-byte store touches 32 rows (half a sector each), and so does the
// residual load.  Here the chunk is transposed through a 2 KB per-warp scratch tile, 16 columns at a time:
// afterwards lane l of store k holds row 8k + l/4, columns 4*(l%4).. of the half, i.e. every instruction
// moves eight 64-byte row segments; the fp32 residual is fetched in the same layout and added there.
// v: bias / ReLU / mask already applied.  m_warp: global row of lane 0.  Converged warp required.
__device__ __forceinline__ void gemm_addend_coalesced(const GemmParams& p, bool do_add, int m_warp, int nb, int lane,
                                                      float4 (&a)[8]) {  // issued before the TMEM wait
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 8 * (i & 3) + (lane >> 2), m = m_warp + r;
    a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (do_add && m < p.M)
      a[i] = __ldg(reinterpret_cast<const float4*>(p.addend + static_cast<int64_t>(m) * p.ldd + nb + 16 * (i >> 2) +
                                                   4 * (lane & 3)));
  }
}
__device__ __forceinline__ void gemm_store_f32_coalesced(const GemmParams& p, const float (&v)[32], const float4 (&a)[8],
                                                         int m_warp, int nb, uint8_t* scratch, int lane) {
  const int my_key = (lane >> 1) & 3;
#pragma unroll
  for (int hlf = 0; hlf < 2; ++hlf) {
    __syncwarp();  // the previous round's reads are done
#pragma unroll
    for (int c = 0; c < 4; ++c)
      *reinterpret_cast<float4*>(scratch + lane * 64 + ((c ^ my_key) << 4)) =
          make_float4(v[16 * hlf + 4 * c], v[16 * hlf + 4 * c + 1], v[16 * hlf + 4 * c + 2], v[16 * hlf + 4 * c + 3]);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = 8 * k + (lane >> 2), c = lane & 3;
      float4 t = *reinterpret_cast<const float4*>(scratch + r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
      const float4 ad = a[4 * hlf + k];
      t.x += ad.x; t.y += ad.y; t.z += ad.z; t.w += ad.w;
      const int m = m_warp + r;
      if (m < p.M)
        *reinterpret_cast<float4*>(static_cast<float*>(p.D) + static_cast<int64_t>(m) * p.ldd + nb + 16 * hlf + 4 * c) = t;
    }
  }
}

// Fetch the residual addend / ReLU-mask words of one chunk (issued before the TMEM wait).
template <int CW>
__device__ __forceinline__ void gemm_epilogue_prefetch(const GemmParams& p, float (&addv)[CW], uint32_t (&maskw)[CW / 2],
                                                       bool do_add, bool do_mask, int m, int nb, bool full) {
  if (do_add) {
    const float* ap = p.addend + static_cast<int64_t>(m) * p.ldd + nb;
    if ((p.ldd % 4 == 0) && full) {
#pragma unroll
      for (int j = 0; j < CW / 4; ++j) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(ap) + j);
        addv[4 * j] = t.x; addv[4 * j + 1] = t.y; addv[4 * j + 2] = t.z; addv[4 * j + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < CW; ++j) addv[j] = (nb + j < p.N) ? ap[j] : 0.f;
    }
  }
  if (do_mask) {
    const bf16* mp = static_cast<const bf16*>(p.relu_mask) + static_cast<int64_t>(m) * p.ldmask + nb;
    if ((p.ldmask % 8 == 0) && full) {
#pragma unroll
      for (int j = 0; j < CW / 8; ++j) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(mp) + j);
        maskw[4 * j] = u.x; maskw[4 * j + 1] = u.y; maskw[4 * j + 2] = u.z; maskw[4 * j + 3] = u.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < CW / 2; ++j) {
        const uint32_t lo = (nb + 2 * j < p.N) ? __bfloat16_as_ushort(mp[2 * j]) : 0u;
        const uint32_t hi = (nb + 2 * j + 1 < p.N) ? __bfloat16_as_ushort(mp[2 * j + 1]) : 0u;
        maskw[j] = lo | (hi << 16);
      }
    }
  }
}

}  // namespace me
