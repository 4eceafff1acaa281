// Relative global attention on the 5th-generation tensor cores -- backward.
//
// One CTA owns a tile of 64 keys of one (batch, head) and walks the query tiles (128 rows) from the
// diagonal down.  Per step, with thread a == query row a == TMEM lane a:
//
//   MMA 1   S  = Q K^T          [128 x 64]      dP = dO V^T      [128 x 64]
//           R  = Q Eband^T      [128 x 192]     (same band of E as the forward kernel)
//   threads x  = S + skew(R)                    Srel[a, b] = R[a, 127 - a + b]
//           P  = exp2(x c - lse)                dS = P (dP - D) / sqrt(dh)
//           P, dS -> shared memory (bf16, UMMA layouts);  dSb = dS in band coordinates
//                                               dSb[a, 127 - a + b] = dS[a, b]  ("unskew" = a shifted store)
//   MMA 2   dV += P^T dO        dK += dS^T Q    dQ_tile = dS K + dSb Eband      dE_tile = dSb^T Q  [192 x dh]
//   threads dQ_tile and dE_tile: TMEM -> shared memory -> fp32 reduce-add into global memory by the TMA
//           unit (cp.reduce.async.bulk), so no thread ever issues an atomic.
//
// dK / dV stay in TMEM for the whole CTA and are written once.  Every MMA runs with M = 128: where the
// operand has only 64 valid rows (P^T, dS^T, the upper part of dSb^T) the second 64-row block is
// whatever follows in shared memory and the corresponding accumulator lanes are never read.
#include "attention_tc.cuh"

namespace me {

constexpr int FB_BM = 128;            // query rows per step
constexpr int FB_BN = 64;             // keys per CTA
constexpr int FB_EROWS = 192;
constexpr int FB_THREADS = 128;
constexpr int FB_STG_STRIDE = 272;    // 256 B of fp32 row + 16 B pad: conflict-free 16-byte stores
constexpr int FB_OFF_K = 0;
constexpr int FB_OFF_V = FB_OFF_K + 8192;
constexpr int FB_OFF_Q = FB_OFF_V + 8192;
constexpr int FB_OFF_DO = FB_OFF_Q + 16384;
constexpr int FB_OFF_E = FB_OFF_DO + 16384;
constexpr int FB_OFF_P = FB_OFF_E + 24576;
constexpr int FB_OFF_DS = FB_OFF_P + 16384;
constexpr int FB_OFF_DSB = FB_OFF_DS + 16384;
constexpr int FB_OFF_STG0 = FB_OFF_DSB + 49152;
constexpr int FB_OFF_STG1 = FB_OFF_STG0 + 128 * FB_STG_STRIDE;
constexpr int FB_OFF_BAR = FB_OFF_STG1 + 128 * FB_STG_STRIDE;
constexpr int FB_SMEM = FB_OFF_BAR + 128;
static_assert(FB_OFF_STG0 % 1024 == 0 && FB_OFF_BAR % 1024 == 0, "tile alignment");
static_assert(FB_SMEM <= 227 * 1024, "shared memory budget");
constexpr uint32_t FB_TMEM_COLS = 512;
constexpr uint32_t FB_COL_S = 0, FB_COL_DP = 64, FB_COL_R = 128, FB_COL_DK = 320, FB_COL_DV = 384, FB_COL_DQ = 448;
constexpr uint32_t FB_COL_DE_LO = 128, FB_COL_DE_HI = 192;  // alias R once it has been consumed

struct FbParams {
  int B, H, L, max_seq;
  int64_t k_sb, k_sh, k_sj, v_sb, v_sh, v_sj, keypad_ld;
  int64_t dq_sb, dq_si;  // fp32 dq accumulator [B, L, H*dh] addressing (elements)
  const uint8_t* keypad;
  const float* lse;
  const float* dsum;
  float* dq_acc;
  float* dE;
  bf16* dk;
  bf16* dv;
  float scale_log2, scale;
};

template <int DH>
__global__ void __launch_bounds__(FB_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                   const __grid_constant__ CUtensorMap tmE, FbParams p) {
  extern __shared__ __align__(1024) uint8_t fb_smem[];
  uint8_t* sK = fb_smem + FB_OFF_K;
  uint8_t* sV = fb_smem + FB_OFF_V;
  uint8_t* sQ = fb_smem + FB_OFF_Q;
  uint8_t* sdO = fb_smem + FB_OFF_DO;
  uint8_t* sE = fb_smem + FB_OFF_E;
  uint8_t* sP = fb_smem + FB_OFF_P;
  uint8_t* sdS = fb_smem + FB_OFF_DS;
  uint8_t* sdSb = fb_smem + FB_OFF_DSB;
  uint8_t* stg0 = fb_smem + FB_OFF_STG0;
  uint8_t* stg1 = fb_smem + FB_OFF_STG1;
  uint64_t* bars = reinterpret_cast<uint64_t*>(fb_smem + FB_OFF_BAR);
  uint64_t* kv_full = bars + 0;
  uint64_t* ld_full = bars + 1;
  uint64_t* m1_done = bars + 2;
  uint64_t* m2_done = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int j0 = kt * FB_BN;
  const int nq = (p.L + FB_BM - 1) / FB_BM;
  const int qi0 = j0 / FB_BM;
  const int nsteps = nq - qi0;

  auto load_step = [&](int st) {
    const int i0 = (qi0 + st) * FB_BM;
    mbar_arrive_expect_tx(ld_full, 16384 + 16384 + 24576);
    tma_load_4d(&tmQ, ld_full, sQ, 0, h, i0, b);
    tma_load_4d(&tmdO, ld_full, sdO, 0, h, i0, b);
    tma_load_2d(&tmE, ld_full, sE, 0, p.max_seq - FB_BM - (i0 - j0));
  };

  if (tid == 0) {
    if ((smem_u32(fb_smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmE);
    mbar_init(kv_full, 1);
    mbar_init(ld_full, 1);
    mbar_init(m1_done, 1);
    mbar_init(m2_done, 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(kv_full, 16384);
    tma_load_4d(&tmK, kv_full, sK, 0, h, j0, b);
    tma_load_4d(&tmV, kv_full, sV, 0, h, j0, b);
    load_step(0);
  }
  // dSb starts as zeros; every step rewrites only the 9 chunks around each row's window
  {
    uint4* z = reinterpret_cast<uint4*>(sdSb);
    for (int c = tid; c < 49152 / 16; c += FB_THREADS) z[c] = make_uint4(0, 0, 0, 0);
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, FB_TMEM_COLS);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  constexpr uint32_t idesc_s = make_idesc_bf16(128, FB_BN, 0, 0);     // S, dP : K-major x K-major
  constexpr uint32_t idesc_r = make_idesc_bf16(128, FB_EROWS, 0, 0);  // R
  constexpr uint32_t idesc_tt = make_idesc_bf16(128, DH, 1, 1);       // dV, dK, dE : A^T (MN-major) x B (MN-major)
  constexpr uint32_t idesc_nt = make_idesc_bf16(128, DH, 0, 1);       // dQ : A (K-major) x B (MN-major)
  const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), q_addr = smem_u32(sQ), do_addr = smem_u32(sdO);
  const uint32_t e_addr = smem_u32(sE), p_addr = smem_u32(sP), ds_addr = smem_u32(sdS), dsb_addr = smem_u32(sdSb);

  const int a = tid;
  const int shift = 31 - lane;
  const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  const uint8_t* kp = p.keypad ? p.keypad + static_cast<int64_t>(b) * p.keypad_ld : nullptr;
  uint32_t kp0 = 0, kp1 = 0;  // key-pad bits of this CTA's 64 keys
  if (kp) {
    const int ja = j0 + lane, jb = j0 + 32 + lane;
    kp0 = __ballot_sync(0xffffffffu, ja < p.L && kp[ja] != 0);
    kp1 = __ballot_sync(0xffffffffu, jb < p.L && kp[jb] != 0);
  }
  const float cs = p.scale_log2;
  // window of row a inside the 192-column band: c = 127 - a + b
  const int win_q0 = (127 - a) >> 3, win_o = (127 - a) & 7;

  for (int st = 0; st < nsteps; ++st) {
    const int i0 = (qi0 + st) * FB_BM;
    const int i = i0 + a;
    const bool row_ok = i < p.L;
    const int64_t stat = (static_cast<int64_t>(b) * p.H + h) * p.L + i;
    float lse2 = INFINITY, Di = 0.f;
    if (row_ok) {
      const float l_nat = p.lse[stat];
      lse2 = (l_nat == -INFINITY) ? INFINITY : l_nat * 1.4426950408889634f;
      Di = p.dsum[stat];
    }
    if (tid == 0) {
      if (st == 0) mbar_wait(kv_full, 0);
      mbar_wait(ld_full, st & 1);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < DH / 16; ++k)
        umma_bf16(tmem_base + FB_COL_S, make_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                  make_smem_desc_sw128(k_addr + k * 32, 16, 1024), idesc_s, k > 0);
#pragma unroll
      for (int k = 0; k < DH / 16; ++k)
        umma_bf16(tmem_base + FB_COL_DP, make_smem_desc_sw128(do_addr + k * 32, 16, 1024),
                  make_smem_desc_sw128(v_addr + k * 32, 16, 1024), idesc_s, k > 0);
#pragma unroll
      for (int k = 0; k < DH / 16; ++k)
        umma_bf16(tmem_base + FB_COL_R, make_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                  make_smem_desc_sw128(e_addr + k * 32, 16, 1024), idesc_r, k > 0);
      umma_commit(m1_done);
    }
    __syncwarp();

    const int lim = i - j0;
    uint32_t v0 = lim >= 31 ? 0xffffffffu : (lim < 0 ? 0u : ((2u << lim) - 1u));
    uint32_t v1 = lim >= 63 ? 0xffffffffu : (lim < 32 ? 0u : ((2u << (lim - 32)) - 1u));
    v0 &= ~kp0;
    v1 &= ~kp1;

    mbar_wait(m1_done, st & 1);
    tc_fence_after();

    uint32_t pw[32], dw[36];  // bf16x2 words of P[a, :] and dS[a, :] (+ 4 zero words for the band shift)
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      uint32_t sv[32], dpv[32], rv[64];
      tmem_ld32(t_lane + FB_COL_S + 32 * ch, sv);
      tmem_ld32(t_lane + FB_COL_DP + 32 * ch, dpv);
      tmem_ld64(t_lane + FB_COL_R + 96 - 32 * warp + 32 * ch, rv);
      tc_wait_ld();
      skew_select(rv, shift);
      const uint32_t vm = ch == 0 ? v0 : v1;
#pragma unroll
      for (int bb = 0; bb < 32; bb += 2) {
        float pr[2], dr[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float x = __uint_as_float(sv[bb + e]) + __uint_as_float(rv[bb + e]);
          float pe = fast_exp2(fmaf(x, cs, -lse2));
          if (!((vm >> (bb + e)) & 1u)) pe = 0.f;
          pr[e] = pe;
          dr[e] = pe * (__uint_as_float(dpv[bb + e]) - Di) * p.scale;
        }
        __nv_bfloat162 ph2 = __floats2bfloat162_rn(pr[0], pr[1]);
        __nv_bfloat162 dh2 = __floats2bfloat162_rn(dr[0], dr[1]);
        pw[16 * ch + bb / 2] = *reinterpret_cast<uint32_t*>(&ph2);
        dw[16 * ch + bb / 2] = *reinterpret_cast<uint32_t*>(&dh2);
      }
    }
    // P and dS rows: UMMA SWIZZLE_128B rows of 128 B (chunk kc of row a at position kc ^ (a & 7))
    {
      uint8_t* prow = sP + a * 128;
      uint8_t* drow = sdS + a * 128;
#pragma unroll
      for (int kc = 0; kc < 8; ++kc) {
        const int pos = (kc ^ (a & 7)) << 4;
        *reinterpret_cast<uint4*>(prow + pos) = make_uint4(pw[4 * kc], pw[4 * kc + 1], pw[4 * kc + 2], pw[4 * kc + 3]);
        *reinterpret_cast<uint4*>(drow + pos) = make_uint4(dw[4 * kc], dw[4 * kc + 1], dw[4 * kc + 2], dw[4 * kc + 3]);
      }
    }
    // dS row in band coordinates: shift right by win_o (0..7) elements inside a 72-element span
    {
      dw[32] = dw[33] = dw[34] = dw[35] = 0u;
      const uint32_t on4 = win_o & 4, on2 = win_o & 2;
#pragma unroll
      for (int w = 35; w >= 0; --w) dw[w] = sel_b32(w >= 2 ? dw[w - 2] : 0u, dw[w], on4);
#pragma unroll
      for (int w = 35; w >= 0; --w) dw[w] = sel_b32(w >= 1 ? dw[w - 1] : 0u, dw[w], on2);
      const uint32_t hs = (win_o & 1) ? 16u : 0u;
#pragma unroll
      for (int w = 35; w >= 0; --w) dw[w] = __funnelshift_l(w >= 1 ? dw[w - 1] : 0u, dw[w], hs);
#pragma unroll
      for (int n = 0; n < 9; ++n) {
        const int q = win_q0 + n;  // 16-byte chunk index inside the 192-column band row (0..23)
        uint8_t* dst = sdSb + (q >> 3) * 16384 + a * 128 + (((q & 7) ^ (a & 7)) << 4);
        *reinterpret_cast<uint4*>(dst) = make_uint4(dw[4 * n], dw[4 * n + 1], dw[4 * n + 2], dw[4 * n + 3]);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t acc0 = st > 0 ? 1u : 0u;
#pragma unroll
      for (int k = 0; k < 8; ++k)  // dV += P^T dO
        umma_bf16(tmem_base + FB_COL_DV, make_smem_desc_sw128(p_addr + k * 2048, 16384, 1024),
                  make_smem_desc_sw128(do_addr + k * 2048, 8192, 1024), idesc_tt, (k > 0) ? 1u : acc0);
#pragma unroll
      for (int k = 0; k < 8; ++k)  // dK += dS^T Q
        umma_bf16(tmem_base + FB_COL_DK, make_smem_desc_sw128(ds_addr + k * 2048, 16384, 1024),
                  make_smem_desc_sw128(q_addr + k * 2048, 8192, 1024), idesc_tt, (k > 0) ? 1u : acc0);
#pragma unroll
      for (int k = 0; k < 4; ++k)  // dQ_tile = dS K
        umma_bf16(tmem_base + FB_COL_DQ, make_smem_desc_sw128(ds_addr + k * 32, 16, 1024),
                  make_smem_desc_sw128(k_addr + k * 2048, 8192, 1024), idesc_nt, k > 0);
#pragma unroll
      for (int k = 0; k < 12; ++k)  // dQ_tile += dSb Eband
        umma_bf16(tmem_base + FB_COL_DQ,
                  make_smem_desc_sw128(dsb_addr + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                  make_smem_desc_sw128(e_addr + k * 2048, 8192, 1024), idesc_nt, 1u);
#pragma unroll
      for (int k = 0; k < 8; ++k)  // dE_tile[0:128] = dSb[:, 0:128]^T Q
        umma_bf16(tmem_base + FB_COL_DE_LO, make_smem_desc_sw128(dsb_addr + k * 2048, 16384, 1024),
                  make_smem_desc_sw128(q_addr + k * 2048, 8192, 1024), idesc_tt, k > 0);
#pragma unroll
      for (int k = 0; k < 8; ++k)  // dE_tile[128:192] = dSb[:, 128:192]^T Q (lanes 64..127 unused)
        umma_bf16(tmem_base + FB_COL_DE_HI, make_smem_desc_sw128(dsb_addr + 32768 + k * 2048, 16384, 1024),
                  make_smem_desc_sw128(q_addr + k * 2048, 8192, 1024), idesc_tt, k > 0);
      umma_commit(m2_done);
    }
    __syncwarp();
    mbar_wait(m2_done, st & 1);
    tc_fence_after();
    if (tid == 0 && st + 1 < nsteps) load_step(st + 1);  // Q / dO / E buffers are free again
    __syncwarp();

    // dQ tile and dE tile: TMEM -> padded fp32 rows in shared memory -> TMA reduce-add
    const int e0 = p.max_seq - FB_BM - (i0 - j0);
    {
      float* row = reinterpret_cast<float*>(stg0 + a * FB_STG_STRIDE);
#pragma unroll
      for (int c0 = 0; c0 < DH; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(t_lane + FB_COL_DQ + c0, v);
        tc_wait_ld();
#pragma unroll
        for (int c = 0; c < 16; c += 4)
          *reinterpret_cast<uint4*>(row + c0 + c) = make_uint4(v[c], v[c + 1], v[c + 2], v[c + 3]);
      }
      float* row1 = reinterpret_cast<float*>(stg1 + a * FB_STG_STRIDE);
#pragma unroll
      for (int c0 = 0; c0 < DH; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(t_lane + FB_COL_DE_LO + c0, v);
        tc_wait_ld();
#pragma unroll
        for (int c = 0; c < 16; c += 4)
          *reinterpret_cast<uint4*>(row1 + c0 + c) = make_uint4(v[c], v[c + 1], v[c + 2], v[c + 3]);
      }
      float* row2 = reinterpret_cast<float*>(sP + a * FB_STG_STRIDE);  // P / dS are free after MMA 2
      if (warp < 2) {
#pragma unroll
        for (int c0 = 0; c0 < DH; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(t_lane + FB_COL_DE_HI + c0, v);
          tc_wait_ld();
#pragma unroll
          for (int c = 0; c < 16; c += 4)
            *reinterpret_cast<uint4*>(row2 + c0 + c) = make_uint4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        }
      }
      fence_proxy_async_smem();
      if (row_ok)
        bulk_reduce_add_f32(p.dq_acc + static_cast<int64_t>(b) * p.dq_sb + static_cast<int64_t>(i) * p.dq_si + h * DH,
                            row, DH * 4);
      if (e0 + a < p.max_seq) bulk_reduce_add_f32(p.dE + static_cast<int64_t>(e0 + a) * DH, row1, DH * 4);
      if (warp < 2 && e0 + 128 + a < p.max_seq)
        bulk_reduce_add_f32(p.dE + static_cast<int64_t>(e0 + 128 + a) * DH, row2, DH * 4);
      bulk_commit();
      bulk_wait_read_all();  // staging rows (and the P/dS area) may be overwritten afterwards
    }
    tc_fence_before();
    __syncthreads();  // all TMEM tiles of this step have been read; shared staging is reusable
  }

  // dK / dV: rows 0..63 of the accumulators (lanes 0..63 = warps 0, 1)
  tc_fence_after();
  if (warp < 2) {
    const int j = j0 + a;
    const bool key_ok = j < p.L;
    bf16* dkrow = p.dk + static_cast<int64_t>(b) * p.k_sb + static_cast<int64_t>(j) * p.k_sj + h * p.k_sh;
    bf16* dvrow = p.dv + static_cast<int64_t>(b) * p.v_sb + static_cast<int64_t>(j) * p.v_sj + h * p.v_sh;
#pragma unroll
    for (int c0 = 0; c0 < DH; c0 += 16) {
      uint32_t vk[16], vv[16];
      tmem_ld16(t_lane + FB_COL_DK + c0, vk);  // .sync.aligned: every lane of the warp takes part
      tmem_ld16(t_lane + FB_COL_DV + c0, vv);
      tc_wait_ld();
      if (key_ok) {
#pragma unroll
        for (int c = 0; c < 16; c += 8) {
          uint4 uk, uv;
          __nv_bfloat162* hk = reinterpret_cast<__nv_bfloat162*>(&uk);
          __nv_bfloat162* hv = reinterpret_cast<__nv_bfloat162*>(&uv);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            hk[e] = __floats2bfloat162_rn(__uint_as_float(vk[c + 2 * e]), __uint_as_float(vk[c + 2 * e + 1]));
            hv[e] = __floats2bfloat162_rn(__uint_as_float(vv[c + 2 * e]), __uint_as_float(vv[c + 2 * e + 1]));
          }
          *reinterpret_cast<uint4*>(dkrow + c0 + c) = uk;
          *reinterpret_cast<uint4*>(dvrow + c0 + c) = uv;
        }
      }
    }
  }
  bulk_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem_base, FB_TMEM_COLS);
  }
}

// dsum[b, h, i] = sum_c dO[b, i, h, c] * O[b, i, h, c]   (the "D" term of the softmax backward)
template <int DH>
__global__ void attn_bwd_prep_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout, int64_t o_sb,
                                     int64_t o_si, int B, int H, int L, float* __restrict__ dsum) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(B) * L * H;
  if (idx >= total) return;
  const int h = static_cast<int>(idx % H);
  const int64_t bi = idx / H;
  const int i = static_cast<int>(bi % L);
  const int b = static_cast<int>(bi / L);
  const int64_t off = b * o_sb + i * o_si + h * DH;
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < DH; c += 8) {
    const uint4 uo = *reinterpret_cast<const uint4*>(out + off + c);
    const uint4 ug = *reinterpret_cast<const uint4*>(dout + off + c);
    const __nv_bfloat162* ho = reinterpret_cast<const __nv_bfloat162*>(&uo);
    const __nv_bfloat162* hg = reinterpret_cast<const __nv_bfloat162*>(&ug);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fo = __bfloat1622float2(ho[e]), fg = __bfloat1622float2(hg[e]);
      acc = fmaf(fo.x, fg.x, acc);
      acc = fmaf(fo.y, fg.y, acc);
    }
  }
  dsum[(static_cast<int64_t>(b) * H + h) * L + i] = acc;
}

// dq (bf16, strided) = dq_acc (fp32 [B, L, H*dh])
__global__ void attn_bwd_dq_convert_kernel(const float* __restrict__ acc, bf16* __restrict__ dq, int64_t q_sb,
                                           int64_t q_si, int L, int d, int64_t total4) {
  for (int64_t i4 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i4 < total4;
       i4 += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t e = i4 * 4;
    const int c = static_cast<int>(e % d);
    const int64_t row = e / d;
    const int i = static_cast<int>(row % L);
    const int64_t b = row / L;
    const float4 v = *reinterpret_cast<const float4*>(acc + e);
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(dq + b * q_sb + i * q_si + c) = u;
  }
}

template <int DH>
static int launch_bwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
                      const CUtensorMap& te, const FbParams& p, dim3 grid, cudaStream_t st) {
  auto kern = attn_bwd_tc_kernel<DH>;
  static bool configured = false;
  if (!configured) {
    ME_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM));
    configured = true;
  }
  kern<<<grid, FB_THREADS, FB_SMEM, st>>>(tq, tk, tv, tdo, te, p);
  ME_LAUNCH_CHECK();
  return 0;
}

int launch_attn_bwd_tc(const me_attn_bwd_args* ba) {
  const me_attn_args* a = &ba->f;
  ME_CHECK(me_device_is_sm100(), "me_attention_backward: the tensor-core path needs an sm_100 device");
  ME_CHECK(a->dtype == ME_BF16, "me_attention_backward: ME_ATTN_TENSOR computes in bf16 only");
  ME_CHECK(a->dh == 32 || a->dh == 48 || a->dh == 64, "me_attention_backward: ME_ATTN_TENSOR supports head dim 32/48/64 (got %d)", a->dh);
  ME_CHECK(a->q_pos0 == 0 && a->Lq == a->Lk && a->pos_dev == nullptr, "me_attention_backward: self-attention only");
  ME_CHECK(a->lse && ba->dsum && ba->dE && ba->dq_acc, "me_attention_backward: lse/dsum/dE/dq_acc required");
  ME_CHECK(a->q_sh == a->dh && a->k_sh == a->dh && a->v_sh == a->dh,
           "me_attention_backward: ME_ATTN_TENSOR expects heads packed along the feature axis (stride dh)");
  const int B = a->B, H = a->H, L = a->Lq, dh = a->dh, d = H * dh;
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  CUtensorMap tq, tk, tv, tdo, te;
  if (qkv_map(&tq, a->q, dh, H, L, B, a->q_sh, a->q_si, a->q_sb, FB_BM)) return 1;
  if (qkv_map(&tk, a->k, dh, H, L, B, a->k_sh, a->k_sj, a->k_sb, FB_BN)) return 1;
  if (qkv_map(&tv, a->v, dh, H, L, B, a->v_sh, a->v_sj, a->v_sb, FB_BN)) return 1;
  if (qkv_map(&tdo, ba->dout, dh, H, L, B, dh, a->o_si, a->o_sb, FB_BM)) return 1;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(dh), static_cast<uint64_t>(a->max_seq)};
    const uint64_t strides[1] = {static_cast<uint64_t>(dh)};
    const uint32_t box[2] = {64, FB_EROWS};
    if (make_tmap_nd_bf16(&te, a->E, 2, dims, strides, box)) return 1;
  }
  // D = rowsum(dO * O), zero the fp32 dq accumulator
  {
    const int64_t total = static_cast<int64_t>(B) * L * H;
    const int blocks = static_cast<int>((total + 127) / 128);
    const bf16* o = static_cast<const bf16*>(a->out);
    const bf16* g = static_cast<const bf16*>(ba->dout);
    if (dh == 64) attn_bwd_prep_kernel<64><<<blocks, 128, 0, st>>>(o, g, a->o_sb, a->o_si, B, H, L, ba->dsum);
    else if (dh == 48) attn_bwd_prep_kernel<48><<<blocks, 128, 0, st>>>(o, g, a->o_sb, a->o_si, B, H, L, ba->dsum);
    else attn_bwd_prep_kernel<32><<<blocks, 128, 0, st>>>(o, g, a->o_sb, a->o_si, B, H, L, ba->dsum);
    ME_LAUNCH_CHECK();
    ME_CUDA(cudaMemsetAsync(ba->dq_acc, 0, static_cast<size_t>(B) * L * d * sizeof(float), st));
  }
  FbParams p;
  p.B = B; p.H = H; p.L = L; p.max_seq = a->max_seq;
  p.k_sb = a->k_sb; p.k_sh = a->k_sh; p.k_sj = a->k_sj;
  p.v_sb = a->v_sb; p.v_sh = a->v_sh; p.v_sj = a->v_sj;
  p.keypad_ld = a->keypad_ld; p.keypad = a->keypad;
  p.dq_sb = static_cast<int64_t>(L) * d; p.dq_si = d;
  p.lse = a->lse; p.dsum = ba->dsum; p.dq_acc = ba->dq_acc; p.dE = ba->dE;
  p.dk = static_cast<bf16*>(ba->dk); p.dv = static_cast<bf16*>(ba->dv);
  p.scale = 1.f / sqrtf(static_cast<float>(dh));
  p.scale_log2 = 1.4426950408889634f * p.scale;
  dim3 grid((L + FB_BN - 1) / FB_BN, H, B);
  int rc;
  if (dh == 64) rc = launch_bwd<64>(tq, tk, tv, tdo, te, p, grid, st);
  else if (dh == 48) rc = launch_bwd<48>(tq, tk, tv, tdo, te, p, grid, st);
  else rc = launch_bwd<32>(tq, tk, tv, tdo, te, p, grid, st);
  if (rc) return rc;
  {
    const int64_t total4 = static_cast<int64_t>(B) * L * d / 4;
    const int64_t want = (total4 + 255) / 256;
    const int blocks = static_cast<int>(want < 148 * 16 ? want : 148 * 16);
    attn_bwd_dq_convert_kernel<<<blocks, 256, 0, st>>>(ba->dq_acc, static_cast<bf16*>(ba->dq), a->q_sb, a->q_si, L, d,
                                                      total4);
    ME_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace me
