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


}  // namespace me
