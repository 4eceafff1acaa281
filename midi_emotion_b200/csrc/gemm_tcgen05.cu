// bf16 GEMM on the 5th-generation tensor cores:  D[M,N] = A . B^T (+ epilogue), fp32 accumulate.
//
//   * operands are moved global -> shared by TMA (cp.async.bulk.tensor, SWIZZLE_128B) through a
//     STAGES-deep mbarrier ring (one producer thread);
//   * tcgen05.mma (cta_group::1, M=128, N=BN, K=16) is issued by one thread; accumulators live in
//     TMEM, double buffered (2 x BN columns) so the epilogue of tile i overlaps the main loop of
//     tile i+1;
//   * four epilogue warps read TMEM with tcgen05.ld (thread == accumulator row), apply
//     bias / ReLU / residual add / ReLU-mask, and write bf16 or fp32 rows straight to global memory;
//   * persistent: grid = min(#tiles, #SMs), static round-robin tile schedule, optional split-K
//     (fp32 atomics) so that weight-gradient GEMMs (tiny M x N, huge K) still fill 148 SMs.
//
// Both operands may be K-major ([rows, K], K contiguous) or MN-major ([K, rows], rows contiguous);
// the latter feeds dgrad (B = W as stored) and wgrad (A = dY^T, B = X^T) without transposes.
#include "gemm_common.cuh"

namespace me {

constexpr int G_BM = 128;
constexpr int G_BK = 64;
constexpr int G_EPI_THREADS = 256;  // warps 2..9: warp w reads TMEM lanes 32*(w&3).., column halves split by (w-2)/4
constexpr int G_THREADS = 64 + G_EPI_THREADS;  // warp0: TMA, warp1: MMA + TMEM alloc, warps 2..9: epilogue

template <int BN>
struct GemmSmem {
  static constexpr int A_BYTES = G_BM * G_BK * 2;
  static constexpr int B_BYTES = BN * G_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int BIAS_BYTES = 2 * BN * 4;  // bias slice of the tile, double buffered with the accumulator
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16 + BIAS_BYTES;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // + alignment slack
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmParams p) {
  using S = GemmSmem<BN>;
  constexpr int STAGES = S::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + STAGES * S::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);  // [2][BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], G_EPI_THREADS);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  pdl_launch_dependents();   // the next kernel on the stream may set itself up now ...
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                // ... and this one touches global memory only once its predecessor has completed

  const int total_tiles = p.num_m_tiles * p.num_n_tiles * p.splits;

  // Producer and issuer warps run their loops converged and one elected lane issues: from a divergent
  // single-thread branch the compiler wraps every TMA / tcgen05.mma instruction in a serialising loop.
  if (warp == 0) {
    // ============================== TMA producer ==============================
    int s = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int ks = tile % p.splits;
      const int mn = tile / p.splits;
      const int m0 = (mn / p.num_n_tiles) * G_BM;
      const int n0 = (mn % p.num_n_tiles) * BN;
      const int kb0 = ks * p.kb_per_split;
      const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[s], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full_bar[s], S::STAGE_BYTES);
          uint8_t* a_dst = smA + s * S::A_BYTES;
          uint8_t* b_dst = smB + s * S::B_BYTES;
          if (!A_MN) {
            tma_load_2d(&tmA, &full_bar[s], a_dst, kb * G_BK, m0);
          } else {
#pragma unroll
            for (int blk = 0; blk < G_BM / 64; ++blk)
              tma_load_2d(&tmA, &full_bar[s], a_dst + blk * (G_BK * 128), m0 + blk * 64, kb * G_BK);
          }
          if (!B_MN) {
            tma_load_2d(&tmB, &full_bar[s], b_dst, kb * G_BK, n0);
          } else {
#pragma unroll
            for (int blk = 0; blk < BN / 64; ++blk)
              tma_load_2d(&tmB, &full_bar[s], b_dst + blk * (G_BK * 128), n0 + blk * 64, kb * G_BK);
          }
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    constexpr uint32_t idesc = make_idesc_bf16(G_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int s = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int ks = tile % p.splits;
      const int kb0 = ks * p.kb_per_split;
      const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[s], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smA + s * S::A_BYTES);
        const uint32_t b_addr = smem_u32(smB + s * S::B_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < G_BK / 16; ++k) {
            const uint64_t ad = A_MN ? make_smem_desc_sw128(a_addr + k * 2048, G_BK * 128, 1024)
                                     : make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t bd = B_MN ? make_smem_desc_sw128(b_addr + k * 2048, G_BK * 128, 1024)
                                     : make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            umma_bf16(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);                     // frees the smem stage when the MMAs retire
          if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 2) {
    // ============================== epilogue ==============================
    const int quarter = warp & 3;        // TMEM lane quarter this warp may access
    const int chalf = (warp - 2) >> 2;   // which half of the tile's columns
    const int et = threadIdx.x - 64;     // 0..255
    const int row_in_tile = quarter * 32 + lane;
    const bool out_bf16 = p.out_dtype == ME_BF16;
    const bool vec_ok = out_bf16 ? (p.ldd % 8 == 0) : (p.ldd % 4 == 0);
    const bool add_vec_ok = (p.ldd % 4 == 0);
    const bool mask_vec_ok = (p.ldmask % 8 == 0);
    constexpr int HALF = BN / 2;
    constexpr int CW = HALF >= 32 ? 32 : HALF;  // columns per TMEM load (BN = 32 -> 16)
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int ks = tile % p.splits;
      const int mn = tile / p.splits;
      const int m0 = (mn / p.num_n_tiles) * G_BM;
      const int n0 = (mn % p.num_n_tiles) * BN;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const bool first_split = (ks == 0);
      const bool use_bias = (p.flags & ME_EPI_BIAS) && first_split;
      // the tile's bias slice goes through shared memory (one global load per thread instead of one per element)
      float* bs = bias_s + acc * BN;
      if (use_bias) {
        for (int c = et; c < BN; c += G_EPI_THREADS) bs[c] = (n0 + c < p.N) ? __ldg(p.bias + n0 + c) : 0.f;
      }
      named_bar_sync(1, G_EPI_THREADS);
      const int m = m0 + row_in_tile;
      const bool row_ok = m < p.M;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = chalf * HALF; c0 < (chalf + 1) * HALF; c0 += CW) {
        if (n0 + c0 >= p.N) break;  // warp-uniform
        const int nb = n0 + c0;
        const bool full = nb + CW <= p.N;
        uint32_t r[CW];
        if (CW == 32) tmem_ld32(t_row + c0, *reinterpret_cast<uint32_t(*)[32]>(r));
        else tmem_ld16(t_row + c0, *reinterpret_cast<uint32_t(*)[16]>(r));
        // operands of the epilogue are fetched while the TMEM load is in flight
        float addv[CW];
        uint32_t maskw[CW / 2];
        const bool do_add = (p.flags & ME_EPI_ADD_F32) && first_split && row_ok;
        const bool do_mask = (p.flags & ME_EPI_RELU_MASK) && row_ok;
        if (do_add) {
          const float* ap = p.addend + static_cast<int64_t>(m) * p.ldd + nb;
          if (add_vec_ok && full) {
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
          if (mask_vec_ok && full) {
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
        tc_wait_ld();
        if (row_ok) {
          float v[CW];
#pragma unroll
          for (int j = 0; j < CW; ++j) v[j] = __uint_as_float(r[j]);
          if (use_bias) {
#pragma unroll
            for (int j = 0; j < CW; ++j) v[j] += bs[c0 + j];
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
              // bf16 > 0  <=>  sign bit clear and magnitude bits non-zero
              const uint32_t w = maskw[j];
              if (!((w & 0x8000u) == 0u && (w & 0x7FFFu) != 0u)) v[2 * j] = 0.f;
              if (!((w & 0x80000000u) == 0u && (w & 0x7FFF0000u) != 0u)) v[2 * j + 1] = 0.f;
            }
          }
          if (p.splits > 1) {
            float* dp = static_cast<float*>(p.D) + static_cast<int64_t>(m) * p.ldd + nb;
#pragma unroll
            for (int j = 0; j < CW; ++j)
              if (nb + j < p.N) atomicAdd(dp + j, v[j]);
          } else if (out_bf16) {
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
        }  // row_ok
      }
      __syncwarp();
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// -------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------
template <int BN, bool A_MN, bool B_MN>
static int launch_one(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int grid,
                      cudaStream_t st) {
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN>;
  static bool configured = false;
  if (!configured) {
    ME_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<BN>::TOTAL));
    configured = true;
  }
  ME_CUDA(launch_pdl(kern, dim3(grid), dim3(G_THREADS), GemmSmem<BN>::TOTAL, st, tmA, tmB, p));
  ME_LAUNCH_CHECK();
  return 0;
}

int launch_gemm_bf16_pair(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb, int ldd,
                          int a_mn, int b_mn, int out_dtype, int flags, const float* bias, const float* addend,
                          const void* relu_mask, int ldmask, int force_splits, cudaStream_t st, float* colsum_out,
                          bool* colsum_done);

int launch_gemm_bf16(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb, int ldd,
                     int a_mn, int b_mn, int out_dtype, int flags, const float* bias, const float* addend,
                     const void* relu_mask, int ldmask, int force_bn, int force_splits, cudaStream_t st,
                     float* colsum_out, bool* colsum_done) {
  if (colsum_done) *colsum_done = false;
  ME_CHECK(M > 0 && N > 0 && K > 0, "me_gemm_bf16: bad dims M=%d N=%d K=%d", M, N, K);
  ME_CHECK(lda % 8 == 0 && ldb % 8 == 0, "me_gemm_bf16: operand row pitches must be multiples of 8 elements (lda=%d ldb=%d)", lda, ldb);
  ME_CHECK((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
           "me_gemm_bf16: operands must be 16-byte aligned");
  // the epilogues use 16-byte stores / loads / reductions whenever the row pitch allows them
  const int d_vec = out_dtype == ME_BF16 ? 8 : 4;
  ME_CHECK(ldd % d_vec != 0 || (reinterpret_cast<uintptr_t>(D) & 15) == 0,
           "me_gemm_bf16: output must be 16-byte aligned (row pitch %d allows vector stores)", ldd);
  ME_CHECK(!(flags & ME_EPI_ADD_F32) || ldd % 4 != 0 || (reinterpret_cast<uintptr_t>(addend) & 15) == 0,
           "me_gemm_bf16: addend must be 16-byte aligned");
  ME_CHECK(!(flags & ME_EPI_RELU_MASK) || ldmask % 8 != 0 || (reinterpret_cast<uintptr_t>(relu_mask) & 15) == 0,
           "me_gemm_bf16: ReLU mask must be 16-byte aligned");
  ME_CHECK(!(flags & ME_EPI_BIAS) || bias, "me_gemm_bf16: bias flag without pointer");
  ME_CHECK(!(flags & ME_EPI_ADD_F32) || addend, "me_gemm_bf16: addend flag without pointer");
  ME_CHECK(!(flags & ME_EPI_RELU_MASK) || relu_mask, "me_gemm_bf16: mask flag without pointer");
  ME_CHECK(!(a_mn && !b_mn), "me_gemm_bf16: (A MN-major, B K-major) is not instantiated");
  ME_CHECK(me_device_is_sm100(), "me_gemm_bf16: tcgen05 path needs an sm_100 device");

  // CTA-pair kernel (256 x 256 tiles, cta_group::2) whenever the problem fills the machine with them
  if (force_bn == 0 || force_bn == 512) {
    const int rc = launch_gemm_bf16_pair(A, B, D, M, N, K, lda, ldb, ldd, a_mn, b_mn, out_dtype, flags, bias, addend,
                                         relu_mask, ldmask, force_bn == 512 ? (force_splits > 0 ? force_splits : -1) : 0, st,
                                         colsum_out, colsum_done);
    if (rc >= 0) return rc;
    ME_CHECK(force_bn != 512, "me_gemm_bf16: the CTA-pair kernel does not take this shape");
  }
  const int sms = sm_count();
  const int num_m_tiles = (M + G_BM - 1) / G_BM;
  // tile width: widest tile that still yields >= ~1 wave of CTAs
  int bn = 256;
  if (force_bn) bn = force_bn;
  else {
    const int min_bn = b_mn ? 128 : 32;
    if (a_mn && b_mn) {
      bn = N > 128 ? 256 : 128;  // weight gradients: wide tiles + split-K beat narrow tiles (gemm_sweep.py)
    } else {
      while (bn > min_bn && num_m_tiles * ((N + bn - 1) / bn) < sms) bn >>= 1;
      if (b_mn && bn < 128) bn = 128;
    }
  }
  ME_CHECK(bn == 32 || bn == 64 || bn == 128 || bn == 256, "me_gemm_bf16: bad tile width %d", bn);
  ME_CHECK(!(b_mn && bn < 128), "me_gemm_bf16: MN-major B needs tile width >= 128");
  const int num_n_tiles = (N + bn - 1) / bn;
  const int num_kb = (K + G_BK - 1) / G_BK;
  int splits = 1;
  if (force_splits > 0) splits = force_splits;
  else if (out_dtype == ME_F32 && !(flags & (ME_EPI_RELU | ME_EPI_RELU_MASK)) && num_kb >= 32) {
    // split-K only where the epilogue is linear (weight gradients): pick the split count whose CTA count
    // fills whole waves of SMs best (each split keeps at least 8 k-blocks)
    // measured on B200 (scripts/gemm_sweep.py): one wave of CTAs, as full as the split count allows
    const int tiles = num_m_tiles * num_n_tiles;
    if (tiles < sms) {
      splits = sms / tiles;
      if (splits > num_kb / 8) splits = num_kb / 8;
      if (splits < 1) splits = 1;
    }
  }
  if (splits > num_kb) splits = num_kb;
  ME_CHECK(splits == 1 || (out_dtype == ME_F32 && !(flags & (ME_EPI_RELU | ME_EPI_RELU_MASK))),
           "me_gemm_bf16: split-K needs fp32 output and a linear epilogue");
  const int kb_per = (num_kb + splits - 1) / splits;
  splits = (num_kb + kb_per - 1) / kb_per;

  CUtensorMap tmA, tmB;
  if (!a_mn) { if (make_tmap_2d_bf16(&tmA, A, K, M, lda, G_BK, G_BM)) return 1; }
  else       { if (make_tmap_2d_bf16(&tmA, A, M, K, lda, 64, G_BK)) return 1; }
  if (!b_mn) { if (make_tmap_2d_bf16(&tmB, B, K, N, ldb, G_BK, bn)) return 1; }
  else       { if (make_tmap_2d_bf16(&tmB, B, N, K, ldb, 64, G_BK)) return 1; }

  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.ldd = ldd; p.ldmask = ldmask; p.flags = flags; p.out_dtype = out_dtype;
  p.num_m_tiles = num_m_tiles; p.num_n_tiles = num_n_tiles; p.splits = splits; p.kb_per_split = kb_per;
  p.num_kb = num_kb; p.bias = bias; p.addend = addend; p.relu_mask = relu_mask; p.D = D;
  p.colsum = nullptr;
  if (splits > 1) {
    ME_CUDA(cudaMemsetAsync(D, 0, static_cast<size_t>(M) * ldd * sizeof(float), st));
  }
  const int total = num_m_tiles * num_n_tiles * splits;
  const int grid = total < sms ? total : sms;

  cudaEvent_t pe = prof_begin(2.0 * M * N * K, st);
#define ME_GEMM_CASE(BNV, AMN, BMN)                                                   \
  if (bn == BNV && a_mn == AMN && b_mn == BMN) {                                      \
    const int rc = launch_one<BNV, (AMN != 0), (BMN != 0)>(tmA, tmB, p, grid, st);    \
    prof_end(pe, st);                                                                 \
    return rc;                                                                        \
  }
  ME_GEMM_CASE(256, 0, 0)
  ME_GEMM_CASE(128, 0, 0)
  ME_GEMM_CASE(64, 0, 0)
  ME_GEMM_CASE(32, 0, 0)
  ME_GEMM_CASE(256, 0, 1)
  ME_GEMM_CASE(128, 0, 1)
  ME_GEMM_CASE(256, 1, 1)
  ME_GEMM_CASE(128, 1, 1)
#undef ME_GEMM_CASE
  set_error("me_gemm_bf16: no kernel for bn=%d a_mn=%d b_mn=%d", bn, a_mn, b_mn);
  return 1;
}

}  // namespace me

extern "C" int me_gemm_bf16(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb, int ldd,
                            int a_mn, int b_mn, int out_dtype, int epi_flags, const float* bias,
                            const float* addend, const void* relu_mask, int ldmask, void* stream) {
  return me::launch_gemm_bf16(A, B, D, M, N, K, lda, ldb, ldd, a_mn, b_mn, out_dtype, epi_flags, bias, addend,
                              relu_mask, ldmask, 0, 0, static_cast<cudaStream_t>(stream), nullptr, nullptr);
}

// test/tuning hook: force the tile width and split count
extern "C" int me_gemm_bf16_ex(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb,
                               int ldd, int a_mn, int b_mn, int out_dtype, int epi_flags, const float* bias,
                               const float* addend, const void* relu_mask, int ldmask, int tile_n, int splits,
                               void* stream) {
  return me::launch_gemm_bf16(A, B, D, M, N, K, lda, ldb, ldd, a_mn, b_mn, out_dtype, epi_flags, bias, addend,
                              relu_mask, ldmask, tile_n, splits, static_cast<cudaStream_t>(stream), nullptr, nullptr);
}
