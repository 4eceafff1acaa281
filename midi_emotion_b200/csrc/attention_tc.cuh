// Shared pieces of the tensor-core relative-attention kernels (forward and backward).
#pragma once
#include "common.cuh"
#include "../../include/midi_emotion_b200.h"

namespace me {

// out[b] = r[b + s] for b < 32, s in [0, 31]: five conditional shifts by 16, 8, 4, 2, 1.
// (selp through inline PTX: left to itself the compiler turns the first stage into a dynamically
// indexed local-memory array.)
__device__ __forceinline__ uint32_t sel_b32(uint32_t a, uint32_t b, uint32_t on) {
  uint32_t d;
  asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tselp.b32 %0, %1, %2, p;\n\t}" : "=r"(d) : "r"(a), "r"(b), "r"(on));
  return d;
}
template <int SH>
__device__ __forceinline__ void skew_stage(uint32_t (&r)[64], uint32_t s) {
  const uint32_t on = s & SH;
#pragma unroll
  for (int i = 0; i < 32 + SH - 1; ++i) r[i] = sel_b32(r[i + SH], r[i], on);
}
__device__ __forceinline__ void skew_select(uint32_t (&r)[64], int s) {
  skew_stage<16>(r, s);
  skew_stage<8>(r, s);
  skew_stage<4>(r, s);
  skew_stage<2>(r, s);
  skew_stage<1>(r, s);
}

// 4-D TMA view of a strided [B, L, H, dh] tensor (q, k, v, dO): dims (dh, H, L, B); the 64-element box
// along dh zero-fills head dims below 64, rows past L are zero-filled as well.
inline int qkv_map(CUtensorMap* m, const void* base, int dh, int H, int L, int B, int64_t sh, int64_t si,
                   int64_t sb, int rows) {
  const uint64_t dims[4] = {static_cast<uint64_t>(dh), static_cast<uint64_t>(H), static_cast<uint64_t>(L),
                            static_cast<uint64_t>(B)};
  const uint64_t strides[3] = {static_cast<uint64_t>(sh), static_cast<uint64_t>(si), static_cast<uint64_t>(sb)};
  const uint32_t box[4] = {64, 1, static_cast<uint32_t>(rows), 1};
  return make_tmap_nd_bf16(m, base, 4, dims, strides, box);
}

// live timing hooks of bench.py (api.cu): class 1 = forward, 2 = backward key side, 3 = backward query side
cudaEvent_t prof_begin(double flops, cudaStream_t st, int cls);
void prof_end(cudaEvent_t e, cudaStream_t st);
// one "unit" of attention work: a causal-minimum L x L x dh product over all heads
inline double attn_unit_flops(int B, int H, int L, int dh) { return 2.0 * B * H * (0.5 * L * L) * dh; }

// ---- pieces shared by the two backward kernels (attention_tc_bwd.cu: dK/dV, attention_tc_bwd_q.cu: dQ/dE) ----
constexpr int FB_DE_COPIES = 32;  // private dE accumulators: concurrently running CTAs walk the same bands of E,
                                  // and same-address reduce-adds serialise in the L2 slices

template <int NCOLS>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[NCOLS]) {  // no wait
  static_assert(NCOLS % 8 == 0, "column groups of 8");
#pragma unroll
  for (int c0 = 0; c0 < NCOLS; c0 += 8) {
    uint32_t t[8];
    tmem_ld8(taddr + c0, t);
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c0 + c] = t[c];
  }
}
// chunk-swizzle key of the fp32 staging / workspace rows (dh floats, no padding; the 16-byte chunk c of row r sits
// at position c ^ (r & key)): 8 rows when the row is a multiple of 128 bytes, else 4
__host__ __device__ constexpr int fb_swizzle_mask(int dh) { return dh % 32 == 0 ? 7 : 3; }
// NCOLS floats of row r, starting at 16-byte chunk `chunk0` of the (unswizzled) row
template <int NCOLS, int SWZ>
__device__ __forceinline__ void sts_row_swz(uint8_t* row, int r, int chunk0, const uint32_t (&v)[NCOLS]) {
#pragma unroll
  for (int c = 0; c < NCOLS / 4; ++c)
    *reinterpret_cast<uint4*>(row + (((chunk0 + c) ^ (r & SWZ)) << 4)) =
        make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}

}  // namespace me
