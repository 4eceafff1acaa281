// Decode-step relative attention over the KV cache (one query row per sequence), bf16 storage, fp32 math.
//
// HBM-bound: per (batch, head) the step streams t+1 rows of K and V once (2 * (t+1) * dh * 2 bytes); the
// matching rows of E (shared by every batch/head, 256 KB per layer) come out of L2.
// One CTA per (batch, head), 8 warps, four lanes per key (each owns a quarter of the row's 16-byte chunks),
// every 4-lane group keeps its own running (max, sum, o) and the partial states are merged once at the
// end (warp butterfly over the key slots, then across warps through shared memory).
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

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// Four lanes share one key: lane g = lane & 3 owns the 16-byte chunks {g, g + 4, ...} of the 128-byte rows
// (so the 4 lanes of a key read 64 contiguous bytes per load instruction), i.e. NC = DH/32 chunks of 8 dims.
template <int NC>
__device__ __forceinline__ void load_chunks(const bf16* __restrict__ row, int g, uint4 (&r)[NC]) {
#pragma unroll
  for (int c = 0; c < NC; ++c) r[c] = __ldg(reinterpret_cast<const uint4*>(row) + g + 4 * c);
}
template <int NC>
__device__ __forceinline__ float dot_chunks(const float (&qf)[NC * 8], const uint4 (&r)[NC]) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const uint32_t w[4] = {r[c].x, r[c].y, r[c].z, r[c].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s0 = fmaf(qf[c * 8 + 2 * e], bf_lo(w[e]), s0);
      s1 = fmaf(qf[c * 8 + 2 * e + 1], bf_hi(w[e]), s1);
    }
  }
  return s0 + s1;
}

// DH in {32, 64, 128}: rows are split into 16-byte chunks dealt round-robin to the 4 lanes of a key.
// DH = 48 (6 chunks) is handled by the DH = 64 instantiation with the tail chunks predicated off.
template <int DHP, int DH>
__global__ void __launch_bounds__(AD_WARPS * 32, 2)
attn_decode_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                   const bf16* __restrict__ E, bf16* __restrict__ out, AdParams p) {
  constexpr int NC = DHP / 32;       // chunks per lane
  constexpr int ND = NC * 8;         // dims per lane
  __shared__ float red[AD_WARPS][DHP + 2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane & 3, kslot = lane >> 2;            // chunk group, key slot inside the warp (0..7)
  const int h = blockIdx.x, b = blockIdx.y;
  const int t = p.pos_dev ? *p.pos_dev : p.q_pos0;      // position of the query == index of the newest key
  // chunk c of this lane covers dims 8*(g + 4c) .. +7 ; for DH = 48 the chunks with g + 4c >= 6 do not exist
  bool live[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) live[c] = 8 * (g + 4 * c) < DH;
  float qf[ND];
  {
    const bf16* qrow = q + b * p.q_sb + h * p.q_sh;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      uint4 u = make_uint4(0, 0, 0, 0);
      if (live[c]) u = __ldg(reinterpret_cast<const uint4*>(qrow) + g + 4 * c);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        qf[c * 8 + 2 * e] = bf_lo(w[e]) * p.scale_log2;   // fold 1/sqrt(dh) * log2(e) into q
        qf[c * 8 + 2 * e + 1] = bf_hi(w[e]) * p.scale_log2;
      }
    }
  }
  const bf16* kb = k + b * p.k_sb + h * p.k_sh;
  const bf16* vb = v + b * p.v_sb + h * p.v_sh;
  const bf16* Eb = E + static_cast<int64_t>(p.max_seq - 1 - t) * DH;  // row for key j is Eb + j*DH
  const uint8_t* kp = p.keypad ? p.keypad + b * p.keypad_ld : nullptr;

  float m = -INFINITY, l = 0.f;
  float o[ND];
#pragma unroll
  for (int c = 0; c < ND; ++c) o[c] = 0.f;

  // Software pipeline: the loads of the next 8 keys are in flight while the current ones are folded in.
  // Rows of pad keys are loaded like any other (the key-pad byte arrives with them) and skipped afterwards,
  // so no load waits on another one.
  struct Tile {
    uint4 kr[NC], er[NC], vr[NC];
    uint32_t pad;
  };
  auto fetch = [&](int j0, Tile& tl) {
    const int j = j0 + kslot;
#pragma unroll
    for (int c = 0; c < NC; ++c) tl.kr[c] = tl.er[c] = tl.vr[c] = make_uint4(0, 0, 0, 0);
    tl.pad = 1u;
    if (j <= t) {
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        if (live[c]) {
          tl.kr[c] = __ldg(reinterpret_cast<const uint4*>(kb + j * p.k_sj) + g + 4 * c);
          tl.er[c] = __ldg(reinterpret_cast<const uint4*>(Eb + static_cast<int64_t>(j) * DH) + g + 4 * c);
          tl.vr[c] = __ldg(reinterpret_cast<const uint4*>(vb + j * p.v_sj) + g + 4 * c);
        }
      }
      tl.pad = kp ? static_cast<uint32_t>(__ldg(kp + j)) : 0u;
    }
  };
  auto fold = [&](const Tile& tl) {
    float x = dot_chunks<NC>(qf, tl.kr) + dot_chunks<NC>(qf, tl.er);
    x += __shfl_xor_sync(0xffffffffu, x, 1);
    x += __shfl_xor_sync(0xffffffffu, x, 2);
    if (tl.pad == 0u) {
      if (x > m) {  // rescale the running state (rare once the maximum has settled)
        const float a = fast_exp2(m - x);
        l *= a;
#pragma unroll
        for (int c = 0; c < ND; ++c) o[c] *= a;
        m = x;
      }
      const float pj = fast_exp2(x - m);
      l += pj;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const uint32_t w[4] = {tl.vr[c].x, tl.vr[c].y, tl.vr[c].z, tl.vr[c].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          o[c * 8 + 2 * e] = fmaf(pj, bf_lo(w[e]), o[c * 8 + 2 * e]);
          o[c * 8 + 2 * e + 1] = fmaf(pj, bf_hi(w[e]), o[c * 8 + 2 * e + 1]);
        }
      }
    }
  };
  {
    constexpr int STEP = AD_WARPS * 8;
    Tile ta, tb;
    int j0 = warp * 8;
    if (j0 <= t) fetch(j0, ta);
    for (; j0 <= t; j0 += 2 * STEP) {   // warp-uniform trip count
      const bool more = j0 + STEP <= t;
      if (more) fetch(j0 + STEP, tb);
      fold(ta);
      if (j0 + 2 * STEP <= t) fetch(j0 + 2 * STEP, ta);
      if (more) fold(tb);
    }
  }

  // merge the 8 key slots of the warp (lanes with equal g), then the warps of the block
  float wm = m;
#pragma unroll
  for (int sft = 4; sft <= 16; sft <<= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, sft));
  const float sc = (m == -INFINITY) ? 0.f : fast_exp2(m - wm);
  l *= sc;
#pragma unroll
  for (int c = 0; c < ND; ++c) o[c] *= sc;
#pragma unroll
  for (int sft = 4; sft <= 16; sft <<= 1) {
    l += __shfl_xor_sync(0xffffffffu, l, sft);
#pragma unroll
    for (int c = 0; c < ND; ++c) o[c] += __shfl_xor_sync(0xffffffffu, o[c], sft);
  }
  if (kslot == 0) {
    if (g == 0) {
      red[warp][DHP] = wm;
      red[warp][DHP + 1] = l;
    }
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int e = 0; e < 8; ++e) red[warp][8 * (g + 4 * c) + e] = o[c * 8 + e];
  }
  __syncthreads();
  if (warp == 0) {
    float gm = -INFINITY;
#pragma unroll
    for (int w = 0; w < AD_WARPS; ++w) gm = fmaxf(gm, red[w][DHP]);
    float gl = 0.f;
    float acc[DHP / 32];
#pragma unroll
    for (int c = 0; c < DHP / 32; ++c) acc[c] = 0.f;
#pragma unroll
    for (int w = 0; w < AD_WARPS; ++w) {
      const float wmx = red[w][DHP];
      const float s2 = (wmx == -INFINITY) ? 0.f : fast_exp2(wmx - gm);
      gl += red[w][DHP + 1] * s2;
#pragma unroll
      for (int c = 0; c < DHP / 32; ++c) acc[c] += red[w][lane + 32 * c] * s2;
    }
    const float inv = gl > 0.f ? 1.f / gl : 0.f;  // fully masked row -> 0
    bf16* orow = out + b * p.o_sb + h * DH;
#pragma unroll
    for (int c = 0; c < DHP / 32; ++c)
      if (lane + 32 * c < DH) orow[lane + 32 * c] = __float2bfloat16_rn(acc[c] * inv);
  }
}

// ------------------------------------------------------------------------------------------------------
// Streaming variant for contiguous cache rows (k_sj == v_sj == dh: the [B, H, T_max, dh] KV cache): the K, V and E
// rows of 64 keys are brought in by bulk copies (cp.async.bulk, one elected thread, mbarrier completion) into a
// four-stage shared-memory ring, so that 128 KB per SM are in flight without costing a register -- the register
// double buffer of the kernel above holds 65 KB per SM, 1.5x the bandwidth-latency product, and stops at
// 5.05 TB/s.  Same thread layout and arithmetic: four lanes per key, every 4-lane group its own running state.
// ------------------------------------------------------------------------------------------------------
constexpr int ADS_KEYS = 64;      // keys per stage
constexpr int ADS_STAGES = 4;

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int DHP, int DH>
__global__ void __launch_bounds__(AD_WARPS * 32 + 32, 2)
attn_decode_stream_kernel(const bf16* __restrict__ q, const bf16* __restrict__ k, const bf16* __restrict__ v,
                          const bf16* __restrict__ E, bf16* __restrict__ out, AdParams p, int n_items) {
  constexpr int NC = DHP / 32;       // chunks per lane
  constexpr int ND = NC * 8;         // dims per lane
  constexpr int ROW = DH * 2;        // bytes per cache row
  constexpr int TILE = ADS_KEYS * ROW;
  extern __shared__ __align__(128) uint8_t ads_smem[];
  // per stage: K | V | E tiles of 64 rows
  uint64_t* full = reinterpret_cast<uint64_t*>(ads_smem + ADS_STAGES * 3 * TILE);
  uint64_t* empty = full + ADS_STAGES;
  float (*red)[DHP + 2] = reinterpret_cast<float (*)[DHP + 2]>(empty + ADS_STAGES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane & 3, kslot = lane >> 2;
  if (threadIdx.x == 0) {
    for (int s = 0; s < ADS_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], AD_WARPS);
    }
    fence_mbar_init();
  }
  pdl_launch_dependents();
  __syncthreads();
  pdl_wait();
  const int t = p.pos_dev ? *p.pos_dev : p.q_pos0;
  const int nkeys = t + 1, nchunks = (nkeys + ADS_KEYS - 1) / ADS_KEYS;
  const bf16* Eb = E + static_cast<int64_t>(p.max_seq - 1 - t) * DH;  // row for key j is Eb + j*DH
  // Persistent: CTA c walks the (batch, head) items c, c + grid, ...; the chunks of all its items form one stream
  // G = 0, 1, ... (stage G % STAGES), so the loads of the next item are in flight while this one is merged.
  const int my_items = (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int total = my_items * nchunks;

  auto issue = [&](int G) {   // (thread 0) chunk G of the stream -> stage G % STAGES
    const int s = G % ADS_STAGES;
    const int item = blockIdx.x + (G / nchunks) * gridDim.x, c = G % nchunks;
    const int b = item / p.H, h = item - b * p.H;
    const int rows = min(ADS_KEYS, nkeys - c * ADS_KEYS);
    const uint32_t bytes = static_cast<uint32_t>(rows) * ROW;
    uint8_t* st = ads_smem + s * 3 * TILE;
    mbar_arrive_expect_tx(&full[s], 3 * bytes);
    const int64_t off = static_cast<int64_t>(c) * ADS_KEYS * DH;
    bulk_load_1d(st, k + b * p.k_sb + h * p.k_sh + off, bytes, &full[s]);
    bulk_load_1d(st + TILE, v + b * p.v_sb + h * p.v_sh + off, bytes, &full[s]);
    bulk_load_1d(st + 2 * TILE, Eb + off, bytes, &full[s]);
  };
  if (warp == AD_WARPS) {   // producer warp: keeps the ring full, never in the way of the compute warps
    if (lane == 0) {
      for (int G = 0; G < total; ++G) {
        if (G >= ADS_STAGES) mbar_wait(&empty[G % ADS_STAGES], ((G / ADS_STAGES) - 1) & 1);
        issue(G);
      }
    }
    return;
  }

  // Chunk cc of this lane is chunk g + 4 (cc ^ sw) of the row: the odd key slots take their two chunks in the other
  // order, so that the eight lanes of a quarter-warp (two key slots, rows 128 bytes apart) read 128 distinct bytes
  // per load instead of hitting the same banks twice.
  const int sw = (NC == 2) ? (kslot & 1) : 0;
  int cix[NC];
  bool live[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    cix[c] = g + 4 * (c ^ sw);
    live[c] = 8 * cix[c] < DH;
  }

  int G = 0;
  for (int it = 0; it < my_items; ++it) {
    const int item = blockIdx.x + it * gridDim.x;
    const int b = item / p.H, h = item - b * p.H;
    const uint8_t* kp = p.keypad ? p.keypad + b * p.keypad_ld : nullptr;
    float qf[ND];
    {
      const bf16* qrow = q + b * p.q_sb + h * p.q_sh;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        uint4 u = make_uint4(0, 0, 0, 0);
        if (live[c]) u = __ldg(reinterpret_cast<const uint4*>(qrow) + cix[c]);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          qf[c * 8 + 2 * e] = bf_lo(w[e]) * p.scale_log2;
          qf[c * 8 + 2 * e + 1] = bf_hi(w[e]) * p.scale_log2;
        }
      }
    }
    float m = -INFINITY, l = 0.f;
    float o[ND];
#pragma unroll
    for (int c = 0; c < ND; ++c) o[c] = 0.f;

    for (int c = 0; c < nchunks; ++c, ++G) {
      const int s = G % ADS_STAGES;
      const uint8_t* st = ads_smem + s * 3 * TILE;
      const int jr = warp * 8 + kslot;               // row of this lane group inside the chunk
      const int j = c * ADS_KEYS + jr;
      uint32_t pad = 1u;
      if (j <= t) pad = kp ? static_cast<uint32_t>(__ldg(kp + j)) : 0u;
      mbar_wait(&full[s], (G / ADS_STAGES) & 1);
      {
        // (every lane goes through the shuffles; the lanes past the last key carry zeros and pad = 1)
        uint4 kr[NC], er[NC], vr[NC];
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) {
          kr[cc] = er[cc] = vr[cc] = make_uint4(0, 0, 0, 0);
          if (live[cc] && j <= t) {
            const int off = jr * ROW + cix[cc] * 16;
            kr[cc] = *reinterpret_cast<const uint4*>(st + off);
            er[cc] = *reinterpret_cast<const uint4*>(st + 2 * TILE + off);
            vr[cc] = *reinterpret_cast<const uint4*>(st + TILE + off);
          }
        }
        float x = dot_chunks<NC>(qf, kr) + dot_chunks<NC>(qf, er);
        x += __shfl_xor_sync(0xffffffffu, x, 1);
        x += __shfl_xor_sync(0xffffffffu, x, 2);
        if (pad == 0u) {
          if (x > m) {
            const float a = fast_exp2(m - x);
            l *= a;
#pragma unroll
            for (int cc = 0; cc < ND; ++cc) o[cc] *= a;
            m = x;
          }
          const float pj = fast_exp2(x - m);
          l += pj;
#pragma unroll
          for (int cc = 0; cc < NC; ++cc) {
            const uint32_t w[4] = {vr[cc].x, vr[cc].y, vr[cc].z, vr[cc].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              o[cc * 8 + 2 * e] = fmaf(pj, bf_lo(w[e]), o[cc * 8 + 2 * e]);
              o[cc * 8 + 2 * e + 1] = fmaf(pj, bf_hi(w[e]), o[cc * 8 + 2 * e + 1]);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }

    // merge the 8 key slots of the warp (lanes with equal g), then the warps of the block
    if (NC == 2) {   // back to the common chunk order first
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float lo = o[c], hi = o[8 + c];
        o[c] = sw ? hi : lo;
        o[8 + c] = sw ? lo : hi;
      }
    }
    float wm = m;
#pragma unroll
    for (int sft = 4; sft <= 16; sft <<= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, sft));
    const float sc = (m == -INFINITY) ? 0.f : fast_exp2(m - wm);
    l *= sc;
#pragma unroll
    for (int c = 0; c < ND; ++c) o[c] *= sc;
#pragma unroll
    for (int sft = 4; sft <= 16; sft <<= 1) {
      l += __shfl_xor_sync(0xffffffffu, l, sft);
#pragma unroll
      for (int c = 0; c < ND; ++c) o[c] += __shfl_xor_sync(0xffffffffu, o[c], sft);
    }
    if (kslot == 0) {
      if (g == 0) {
        red[warp][DHP] = wm;
        red[warp][DHP + 1] = l;
      }
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int e = 0; e < 8; ++e) red[warp][8 * (g + 4 * c) + e] = o[c * 8 + e];
    }
    named_bar_sync(1, AD_WARPS * 32);
    if (warp == 0) {
      float gm = -INFINITY;
#pragma unroll
      for (int w = 0; w < AD_WARPS; ++w) gm = fmaxf(gm, red[w][DHP]);
      float gl = 0.f;
      float acc[DHP / 32];
#pragma unroll
      for (int c = 0; c < DHP / 32; ++c) acc[c] = 0.f;
#pragma unroll
      for (int w = 0; w < AD_WARPS; ++w) {
        const float wmx = red[w][DHP];
        const float s2 = (wmx == -INFINITY) ? 0.f : fast_exp2(wmx - gm);
        gl += red[w][DHP + 1] * s2;
#pragma unroll
        for (int c = 0; c < DHP / 32; ++c) acc[c] += red[w][lane + 32 * c] * s2;
      }
      const float inv = gl > 0.f ? 1.f / gl : 0.f;  // fully masked row -> 0
      bf16* orow = out + b * p.o_sb + h * DH;
#pragma unroll
      for (int c = 0; c < DHP / 32; ++c)
        if (lane + 32 * c < DH) orow[lane + 32 * c] = __float2bfloat16_rn(acc[c] * inv);
    }
    named_bar_sync(1, AD_WARPS * 32);   // `red` is free again
  }
}

template <int DHP, int DH>
static int launch_decode_stream(const bf16* q, const bf16* k, const bf16* v, const bf16* E, bf16* out, const AdParams& p,
                                int n_items, cudaStream_t st) {
  constexpr int SMEM = ADS_STAGES * 3 * ADS_KEYS * DH * 2 + 2 * ADS_STAGES * 8 + AD_WARPS * (DHP + 2) * 4;
  auto kern = attn_decode_stream_kernel<DHP, DH>;
  static bool configured = false;
  if (!configured) {
    ME_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  const int grid = std::min(n_items, 2 * sm_count());
  ME_CUDA(launch_pdl(kern, dim3(grid), dim3(AD_WARPS * 32 + 32), SMEM, st, q, k, v, E, out, p, n_items));
  ME_LAUNCH_CHECK();
  return 0;
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
  // contiguous cache rows, 16-byte aligned bases: the bulk-copy pipeline (ME_ATTN_DECODE=1 keeps the register one)
  static const bool reg_pipe = [] { const char* e = getenv("ME_ATTN_DECODE"); return e != nullptr && e[0] == '1'; }();
  const bool contiguous = a->k_sj == a->dh && a->v_sj == a->dh && a->k_sh % 8 == 0 && a->k_sb % 8 == 0 &&
                          a->v_sh % 8 == 0 && a->v_sb % 8 == 0 && (reinterpret_cast<uintptr_t>(a->k) & 15) == 0 &&
                          (reinterpret_cast<uintptr_t>(a->v) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->E) & 15) == 0;
  if (contiguous && !reg_pipe) {
    if (a->dh == 64) return launch_decode_stream<64, 64>(q, k, v, E, out, p, a->H * a->B, st);
    if (a->dh == 48) return launch_decode_stream<64, 48>(q, k, v, E, out, p, a->H * a->B, st);
    return launch_decode_stream<32, 32>(q, k, v, E, out, p, a->H * a->B, st);
  }
  if (a->dh == 64) attn_decode_kernel<64, 64><<<grid, AD_WARPS * 32, 0, st>>>(q, k, v, E, out, p);
  else if (a->dh == 48) attn_decode_kernel<64, 48><<<grid, AD_WARPS * 32, 0, st>>>(q, k, v, E, out, p);
  else attn_decode_kernel<32, 32><<<grid, AD_WARPS * 32, 0, st>>>(q, k, v, E, out, p);
  ME_LAUNCH_CHECK();
  return 0;
}

}  // namespace me
