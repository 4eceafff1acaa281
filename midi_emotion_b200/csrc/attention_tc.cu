// Relative global attention (Music Transformer) on the 5th-generation tensor cores -- forward.
//
//   S[i,j] = (q_i.k_j + q_i.E[max_seq-1-(i-j)]) / sqrt(dh),  j <= i, key j not pad;   O = softmax(S) V
//
// One CTA owns 128 query rows of one (batch, head) and walks the key tiles (64 keys) up to the
// diagonal.  Two CTAs are resident per SM (256 TMEM columns, ~112 KB shared memory each) so that one
// CTA's softmax overlaps the other's MMAs.
//
//   * Q, K, V and the needed band of E are brought in by TMA (4-D maps over the strided q/k/v views,
//     SWIZZLE_128B; rows past the sequence / past max_seq are zero-filled by the TMA unit);
//   * tcgen05.mma (M=128): [S | R] = Q [K ; Eband]^T (N=256, K and the band of E back to back in shared
//     memory), i.e. S = Q K^T (N=64) and the relative band R = Q Eband^T (N=192), where Eband is
//     the 191 consecutive rows of E a 128x64 tile can touch: R[a, c] = q_a . E[e0 + c],
//     e0 = max_seq - 128 - (i0 - j0).  The skew of the reference (pad/reshape/slice,
//     music_multi.py:245-262) is the identity  Srel[a, b] = R[a, 127 - a + b]: thread a (TMEM lane a)
//     reads its row of R with a warp-uniform column base and applies the per-lane part of the shift
//     (0..31 columns) with a 5-stage select network in registers -- nothing L^2-sized exists anywhere;
//   * online softmax in fp32 (exp2, running max/sum per row = per thread, no shuffles needed);
//   * P (bf16) goes to shared memory in the UMMA K-major layout, O_tile = P V (N = dh) lands in TMEM
//     and is folded into the fp32 output accumulator held in registers.
#include "attention_tc.cuh"

namespace me {

constexpr int FA_BM = 128;      // query rows per CTA
constexpr int FA_BN = 64;       // keys per tile
constexpr int FA_EROWS = 192;   // rows of E per tile (191 needed)
constexpr int FA_THREADS = 128; // thread = query row (TMEM lane); thread 0 also issues TMA and MMA
constexpr int FA_Q_BYTES = FA_BM * 128;
constexpr int FA_K_BYTES = FA_BN * 128;
constexpr int FA_E_BYTES = FA_EROWS * 128;
constexpr int FA_P_BYTES = FA_BM * 128;
constexpr int FA_STAGE_BYTES = 2 * FA_K_BYTES + FA_E_BYTES;
constexpr int FA_SMEM = FA_Q_BYTES + 2 * FA_STAGE_BYTES + FA_P_BYTES + 128;
constexpr uint32_t FA_TMEM_COLS = 256;
constexpr uint32_t FA_COL_S = 0, FA_COL_R = 64;

struct FaParams {
  int B, H, L, max_seq;
  int64_t o_sb, o_si, keypad_ld;
  const uint8_t* keypad;
  bf16* out;
  float* lse;
  float scale_log2;  // log2(e) / sqrt(dh)
  float sqrt_dh;
  float* m_tiles;    // with the tmP map: the probability tiles and their exponent offsets are saved for backward
  int tiles_per_head;
  int noncausal;     // ME_ATTN_NONCAUSAL: every key < L is visible (the band of E zero-fills itself above the diagonal)
};

// fp32 value rounded to the nearest bf16 (the reference's rounding points under autocast)
__device__ __forceinline__ float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// RR = ME_ATTN_REF_ROUNDING: QK^T, Srel, their sum and the scaled logits are rounded to bf16 where the reference
// rounds them under autocast (music_multi.py:215-222: einsum -> bf16, matmul -> bf16, bf16 + bf16, bf16 / sqrt(dh)).
template <int DH, bool RR>
__global__ void __launch_bounds__(FA_THREADS, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmE,
                   const __grid_constant__ CUtensorMap tmP, FaParams p) {
  extern __shared__ __align__(1024) uint8_t fa_smem[];
  uint8_t* sQ = fa_smem;
  uint8_t* sStage = sQ + FA_Q_BYTES;  // per stage: K | E | V  ([K ; Eband] is one 256-row B operand)
  uint8_t* sP = sStage + 2 * FA_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + FA_P_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_free = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* o_full = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qi = gridDim.x - 1 - blockIdx.x;  // heavy (late) query tiles first
  const int h = blockIdx.y, b = blockIdx.z;
  const int i0 = qi * FA_BM;
  const int kmax = p.noncausal ? p.L : min(i0 + FA_BM, p.L);  // keys 0 .. kmax-1 can be visible
  const int nt = (kmax + FA_BN - 1) / FA_BN;
  const bool save = p.m_tiles != nullptr;
  const int64_t tile0 = (static_cast<int64_t>(b) * p.H + h) * p.tiles_per_head +
                        (p.noncausal ? static_cast<int64_t>(qi) * ((p.L + FA_BN - 1) / FA_BN) : static_cast<int64_t>(qi) * (qi + 1));

  // Thread 0 is also the TMA producer and the MMA issuer: the per-tile schedule is a strict sequence
  // (S/R MMAs -> softmax -> P.V MMA -> accumulate), the overlap comes from the second CTA on the SM.
  auto load_tile = [&](int t) {
    const int s = t & 1, j0 = t * FA_BN;
    uint8_t* st = sStage + s * FA_STAGE_BYTES;
    mbar_arrive_expect_tx(&kv_full[s], FA_STAGE_BYTES);
    tma_load_4d(&tmK, &kv_full[s], st, 0, h, j0, b);
    tma_load_2d(&tmE, &kv_full[s], st + FA_K_BYTES, 0, p.max_seq - FA_BM - (i0 - j0));
    tma_load_4d(&tmV, &kv_full[s], st + FA_K_BYTES + FA_E_BYTES, 0, h, j0, b);
  };

  if (tid == 0) {
    if ((smem_u32(fa_smem) & 1023u) != 0) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmE);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_free[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(o_full, 1);
    fence_mbar_init();
    mbar_arrive_expect_tx(q_full, FA_Q_BYTES);
    tma_load_4d(&tmQ, q_full, sQ, 0, h, i0, b);
    load_tile(0);
    if (nt > 1) load_tile(1);
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, FA_TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  constexpr uint32_t idesc_sr = make_idesc_bf16(FA_BM, FA_BN + FA_EROWS, 0, 0);
  constexpr uint32_t idesc_o = make_idesc_bf16(FA_BM, DH, 0, 1);
  const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);

  const int a = tid;                    // query row inside the tile == TMEM lane
  const int i = i0 + a;
  const int shift = 31 - lane;          // per-lane part of the skew
  const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  const uint8_t* kp = p.keypad ? p.keypad + static_cast<int64_t>(b) * p.keypad_ld : nullptr;
  float O[DH];
#pragma unroll
  for (int c = 0; c < DH; ++c) O[c] = 0.f;
  float m = -INFINITY, l = 0.f;
  // RR: x is already divided by sqrt(dh) (and rounded), only log2(e) remains
  const float cs = RR ? 1.4426950408889634f : p.scale_log2;
  const bool any_kp = kp != nullptr || p.noncausal;

  for (int t = 0; t < nt; ++t) {
    const int s = t & 1;
    const uint32_t ph = (t >> 1) & 1;
    const uint32_t k_addr = smem_u32(sStage + s * FA_STAGE_BYTES);
    const uint32_t v_addr = k_addr + FA_K_BYTES + FA_E_BYTES;
    // Warp 0 issues converged with one elected lane: from a divergent `if (tid == 0)` the compiler wraps
    // every tcgen05.mma in a serialising loop (~90 cycles per instruction).
    if (warp == 0) {
      if (t == 0) mbar_wait(q_full, 0);
      mbar_wait(&kv_full[s], ph);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)  // [S | R] = Q [K ; Eband]^T, one N = 256 instruction per 16 head dims
          umma_bf16(tmem_base + FA_COL_S, make_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                    make_smem_desc_sw128(k_addr + k * 32, 16, 1024), idesc_sr, k > 0);
        umma_commit(s_full);
      }
      __syncwarp();
    }

    const int j0 = t * FA_BN;
    uint32_t kp0 = 0, kp1 = 0;
    if (any_kp) {  // (keys past the sequence only matter without the causal predicate)
      const int ja = j0 + lane, jb = j0 + 32 + lane;
      kp0 = __ballot_sync(0xffffffffu, ja >= p.L || (kp && kp[ja] != 0));
      kp1 = __ballot_sync(0xffffffffu, jb >= p.L || (kp && kp[jb] != 0));
    }
    const int lim = p.noncausal ? 63 : i - j0;  // columns b <= lim are causal-visible
    uint32_t v0 = lim >= 31 ? 0xffffffffu : (lim < 0 ? 0u : ((2u << lim) - 1u));
    uint32_t v1 = lim >= 63 ? 0xffffffffu : (lim < 32 ? 0u : ((2u << (lim - 32)) - 1u));
    v0 &= ~kp0;
    v1 &= ~kp1;
    const bool need_mask = !__all_sync(0xffffffffu, (v0 & v1) == 0xffffffffu);

    mbar_wait(s_full, t & 1);
    tc_fence_after();

    float x0[32], x1[32];
    {
      uint32_t sv[32], rv[64];
      tmem_ld32(t_lane + FA_COL_S, sv);
      tmem_ld64(t_lane + FA_COL_R + 96 - 32 * warp, rv);
      tc_wait_ld();
      skew_select(rv, shift);
#pragma unroll
      for (int bb = 0; bb < 32; ++bb) {
        if (RR)
          x0[bb] = bf16r(__fdiv_rn(bf16r(bf16r(__uint_as_float(sv[bb])) + bf16r(__uint_as_float(rv[bb]))), p.sqrt_dh));
        else
          x0[bb] = __uint_as_float(sv[bb]) + __uint_as_float(rv[bb]);
      }
    }
    {
      uint32_t sv[32], rv[64];
      tmem_ld32(t_lane + FA_COL_S + 32, sv);
      tmem_ld64(t_lane + FA_COL_R + 128 - 32 * warp, rv);
      tc_wait_ld();
      skew_select(rv, shift);
#pragma unroll
      for (int bb = 0; bb < 32; ++bb) {
        if (RR)
          x1[bb] = bf16r(__fdiv_rn(bf16r(bf16r(__uint_as_float(sv[bb])) + bf16r(__uint_as_float(rv[bb]))), p.sqrt_dh));
        else
          x1[bb] = __uint_as_float(sv[bb]) + __uint_as_float(rv[bb]);
      }
    }
    if (need_mask) {
#pragma unroll
      for (int bb = 0; bb < 32; ++bb) {
        if (!((v0 >> bb) & 1u)) x0[bb] = -INFINITY;
        if (!((v1 >> bb) & 1u)) x1[bb] = -INFINITY;
      }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int bb = 0; bb < 32; ++bb) mx = fmaxf(mx, fmaxf(x0[bb], x1[bb]));
    const float m_new = fmaxf(m, mx * cs);
    const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
    const float alpha = fast_exp2(m - m_use);  // m = -inf -> 0
    float rs = 0.f;
    // P row -> shared memory, UMMA K-major SWIZZLE_128B: 16-byte chunk kc of row a lives at chunk kc ^ (a & 7)
    uint8_t* prow = sP + a * 128;
#pragma unroll
    for (int kc = 0; kc < 8; ++kc) {
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int bb = (kc & 3) * 8 + 2 * e;
        const float xa = (kc < 4) ? x0[bb] : x1[bb];
        const float xb = (kc < 4) ? x0[bb + 1] : x1[bb + 1];
        const float pa = fast_exp2(fmaf(xa, cs, -m_use));
        const float pb = fast_exp2(fmaf(xb, cs, -m_use));
        rs += pa + pb;
        __nv_bfloat162 h2 = __floats2bfloat162_rn(pa, pb);
        w[e] = *reinterpret_cast<uint32_t*>(&h2);
      }
      *reinterpret_cast<uint4*>(prow + ((kc ^ (a & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    l = l * alpha + rs;
    m = m_new;
    if (save) p.m_tiles[(tile0 + t) * FA_BM + a] = m_use;   // P[a, :] of this tile = exp2(x c - m_use)
    fence_proxy_async_smem();  // generic-proxy writes of P -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();           // P complete, S/R consumed by every row
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < FA_BN / 16; ++k)
          umma_bf16(tmem_base + FA_COL_S, make_smem_desc_sw128(p_addr + k * 32, 16, 1024),
                    make_smem_desc_sw128(v_addr + k * 2048, 8192, 1024), idesc_o, k > 0);
        umma_commit(o_full);
        umma_commit(&kv_free[s]);
        if (save) {   // the tile of P, as the tensor core reads it, goes to the saved-activation tensor
          tma_store_2d(&tmP, sP, 0, static_cast<int>((tile0 + t) * FA_BM));
          bulk_commit();
        }
      }
      __syncwarp();
    }
    mbar_wait(o_full, t & 1);
    tc_fence_after();
#pragma unroll
    for (int c0 = 0; c0 < DH; c0 += 16) {
      uint32_t ov[16];
      tmem_ld16(t_lane + FA_COL_S + c0, ov);
      tc_wait_ld();
#pragma unroll
      for (int c = 0; c < 16; ++c) O[c0 + c] = fmaf(O[c0 + c], alpha, __uint_as_float(ov[c]));
    }
    tc_fence_before();
    if (warp == 0 && t + 2 < nt) {
      mbar_wait(&kv_free[s], ph);
      if (elect_one()) load_tile(t + 2);
      __syncwarp();
    }
    if (warp == 0 && save) {   // (elect.sync names the same lane every time: the one that committed the store)
      if (elect_one()) bulk_wait_read_all();
      __syncwarp();
    }
    __syncthreads();           // P.V tile read out of TMEM by every row before the next S/R MMAs
  }

  if (i < p.L) {
    const float inv = l > 0.f ? 1.f / l : 0.f;  // fully masked row -> 0 (reference: NaN)
    bf16* orow = p.out + static_cast<int64_t>(b) * p.o_sb + static_cast<int64_t>(i) * p.o_si + h * DH;
#pragma unroll
    for (int c0 = 0; c0 < DH; c0 += 8) {
      uint4 u;
      __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(O[c0 + 2 * e] * inv, O[c0 + 2 * e + 1] * inv);
      *reinterpret_cast<uint4*>(orow + c0) = u;
    }
    if (p.lse)
      p.lse[(static_cast<int64_t>(b) * p.H + h) * p.L + i] =
          l > 0.f ? (m + log2f(l)) * 0.69314718055994530942f : -INFINITY;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem_base, FA_TMEM_COLS);
  }
}

template <int DH, bool RR>
static int launch_fwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& te,
                      const CUtensorMap& tp, const FaParams& p, dim3 grid, cudaStream_t st) {
  auto kern = attn_fwd_tc_kernel<DH, RR>;
  static bool configured = false;
  if (!configured) {
    ME_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
    ME_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    configured = true;
  }
  cudaEvent_t pe = prof_begin(3.0 * attn_unit_flops(p.B, p.H, p.L, DH), st, 1);   // QK^T, QE^T, PV
  kern<<<grid, FA_THREADS, FA_SMEM, st>>>(tq, tk, tv, te, tp, p);
  prof_end(pe, st);
  ME_LAUNCH_CHECK();
  return 0;
}

int launch_attn_fwd2_tc(const me_attn_args* a);   // attention_tc_fwd2.cu: the pipelined one-CTA-per-SM schedule

int launch_attn_fwd_tc(const me_attn_args* a) {
  ME_CHECK(me_device_is_sm100(), "me_attention_forward: the tensor-core path needs an sm_100 device");
  ME_CHECK(a->dtype == ME_BF16, "me_attention_forward: ME_ATTN_TENSOR computes in bf16 only");
  ME_CHECK(a->dh == 32 || a->dh == 48 || a->dh == 64, "me_attention_forward: ME_ATTN_TENSOR supports head dim 32/48/64 (got %d)", a->dh);
  ME_CHECK(a->q_pos0 == 0 && a->Lq == a->Lk && a->pos_dev == nullptr,
           "me_attention_forward: ME_ATTN_TENSOR handles full self-attention (Lq == Lk, q_pos0 == 0)");
  ME_CHECK(a->Lq > 0 && a->Lq <= a->max_seq, "me_attention_forward: bad sequence length %d", a->Lq);
  ME_CHECK(a->q_sh > 0 && a->q_si > 0 && a->q_sb > 0 && a->k_sh > 0 && a->v_sh > 0, "me_attention_forward: bad strides");
  ME_CHECK(a->o_si % 8 == 0 && a->o_sb % 8 == 0 && (reinterpret_cast<uintptr_t>(a->out) & 15) == 0,
           "me_attention_forward: output rows must be 16-byte aligned");
  {
    static const bool two_cta = [] { const char* e = getenv("ME_ATTN_FWD"); return e != nullptr && e[0] == '1'; }();
    if (!two_cta) {
      const int rc = launch_attn_fwd2_tc(a);
      if (rc >= 0) return rc;
    }
  }
  CUtensorMap tq, tk, tv, te;
  if (qkv_map(&tq, a->q, a->dh, a->H, a->Lq, a->B, a->q_sh, a->q_si, a->q_sb, FA_BM)) return 1;
  if (qkv_map(&tk, a->k, a->dh, a->H, a->Lk, a->B, a->k_sh, a->k_sj, a->k_sb, FA_BN)) return 1;
  if (qkv_map(&tv, a->v, a->dh, a->H, a->Lk, a->B, a->v_sh, a->v_sj, a->v_sb, FA_BN)) return 1;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(a->dh), static_cast<uint64_t>(a->max_seq)};
    const uint64_t strides[1] = {static_cast<uint64_t>(a->dh)};
    const uint32_t box[2] = {64, FA_EROWS};
    if (make_tmap_nd_bf16(&te, a->E, 2, dims, strides, box)) return 1;
  }
  FaParams p;
  p.B = a->B; p.H = a->H; p.L = a->Lq; p.max_seq = a->max_seq;
  p.o_sb = a->o_sb; p.o_si = a->o_si; p.keypad_ld = a->keypad_ld; p.keypad = a->keypad;
  p.out = static_cast<bf16*>(a->out);
  p.lse = a->lse;
  p.sqrt_dh = sqrtf(static_cast<float>(a->dh));
  p.scale_log2 = 1.4426950408889634f / p.sqrt_dh;
  p.noncausal = (a->flags & ME_ATTN_NONCAUSAL) ? 1 : 0;
  p.m_tiles = nullptr;
  p.tiles_per_head = static_cast<int>(me_attention_saved_tiles(a->Lq, a->flags));
  CUtensorMap tp = te;   // (unused unless the tiles are saved)
  if (a->p_tiles != nullptr && a->m_tiles != nullptr) {
    const uint64_t rows = static_cast<uint64_t>(a->B) * a->H * p.tiles_per_head * FA_BM;
    ME_CHECK(rows < (1ull << 31), "me_attention_forward: saved-tile tensor too large");
    const uint64_t dims[2] = {64, rows};
    const uint64_t strides[1] = {64};
    const uint32_t box[2] = {64, FA_BM};
    if (make_tmap_nd_bf16(&tp, a->p_tiles, 2, dims, strides, box)) return 1;
    p.m_tiles = a->m_tiles;
  }
  dim3 grid((a->Lq + FA_BM - 1) / FA_BM, a->H, a->B);
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  if (a->flags & ME_ATTN_REF_ROUNDING) {
    if (a->dh == 64) return launch_fwd<64, true>(tq, tk, tv, te, tp, p, grid, st);
    if (a->dh == 48) return launch_fwd<48, true>(tq, tk, tv, te, tp, p, grid, st);
    return launch_fwd<32, true>(tq, tk, tv, te, tp, p, grid, st);
  }
  if (a->dh == 64) return launch_fwd<64, false>(tq, tk, tv, te, tp, p, grid, st);
  if (a->dh == 48) return launch_fwd<48, false>(tq, tk, tv, te, tp, p, grid, st);
  return launch_fwd<32, false>(tq, tk, tv, te, tp, p, grid, st);
}

}  // namespace me

extern "C" int64_t me_attention_saved_tiles(int L, int flags) {
  const int64_t nq = (L + me::FA_BM - 1) / me::FA_BM, nkt = (L + me::FA_BN - 1) / me::FA_BN;
  return (flags & ME_ATTN_NONCAUSAL) ? nq * nkt : nq * (nq + 1);
}
