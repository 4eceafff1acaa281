// bf16 GEMM on CTA pairs (tcgen05 cta_group::2):  D[M,N] = A . B^T (+ epilogue), 256 x 256 tile per pair.
//
// The two CTAs of a cluster sit on the two SMs of a TPC.  Each loads its own 128 rows of A and HALF of the
// B tile (128 of the 256 N-rows), so every SM pulls 32 KB per 64-wide k-block instead of 48 KB -- the
// 1-CTA kernel is bound by the per-SM L2->SM path, not by the tensor pipe.  The leader CTA (rank 0)
// issues tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16); each CTA ends up with its 128 accumulator
// rows x 256 columns in its own tensor memory, double buffered (2 x 256 columns) against the epilogue.
//
//   warp 0  TMA producer (both CTAs; cta_group::2 loads signal the LEADER's full barrier)
//   warp 1  TMEM alloc (both CTAs); MMA issue (leader only); tcgen05.commit multicast to both CTAs
//   warps 2..9  epilogue (tcgen05.ld, bias/ReLU/residual/mask, 16-byte stores or split-K atomics);
//               the peer's epilogue threads arrive remotely on the leader's "accumulator free" barrier.
#include "gemm_common.cuh"

namespace me {

constexpr int G2_BM = 256, G2_BN = 256, G2_BK = 64;  // pair tile
constexpr int G2_STAGES = 6;          // direct-store epilogue
constexpr int G2_STAGES_STAGED = 4;   // staged epilogue: 64 KB of the ring become the output staging tile
constexpr int G2_STG_BYTES = 128 * G2_BN * 2;  // bf16 [128 rows x 256 cols] as 4 SWIZZLE_128B blocks of 64 columns
constexpr int G2_A_BYTES = 128 * G2_BK * 2;          // per CTA: its 128 rows of A
constexpr int G2_B_BYTES = 128 * G2_BK * 2;          // per CTA: its half of the B tile
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
constexpr int G2_EPI_THREADS = 256;
constexpr int G2_THREADS = 64 + G2_EPI_THREADS;
constexpr int G2_TRANSPOSE_BYTES = (G2_EPI_THREADS / 32) * 2048;  // per-warp scratch of the coalesced fp32 epilogue
constexpr int G2_SMEM = G2_STAGES * G2_STAGE_BYTES + (2 * G2_STAGES + 5) * 8 + 16 + 2 * G2_BN * 4 + 16 +
                        G2_TRANSPOSE_BYTES + 1024;
constexpr uint32_t G2_PEER_MASK = 0xFEFFFFFFu;       // clears the CTA-rank bit of a shared::cluster address

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* m, uint64_t* leader_bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(leader_bar) & G2_PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the same-offset mbarrier of both CTAs once the MMAs issued so far have retired
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(smem_u32(bar)));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

// STAGED: bf16 output tile (and, for the ReLU-mask epilogue, the mask tile) goes through a swizzled
// shared-memory tile and the TMA unit, so global traffic of the epilogue is fully coalesced.
template <bool A_MN, bool B_MN, bool STAGED>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmMask, GemmParams p) {
  constexpr int G2_STAGES = STAGED ? me::G2_STAGES_STAGED : me::G2_STAGES;
  extern __shared__ uint8_t smem_raw2[];
  // identical carve-up in both CTAs (same offsets: the MMA and the multicast commit address both by offset)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw2) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + G2_STAGES * G2_A_BYTES;
  uint8_t* stg = smem + G2_STAGES * G2_STAGE_BYTES;  // STAGED only
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + G2_STAGES * G2_STAGE_BYTES + (STAGED ? G2_STG_BYTES : 0));
  uint64_t* empty_bar = full_bar + G2_STAGES;
  uint64_t* tfull_bar = empty_bar + G2_STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;          // [2] (used on the leader; both CTAs' epilogues arrive there)
  uint64_t* mask_bar = tempty_bar + 2;           // mask tile landed in the staging buffer (STAGED + mask)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mask_bar + 1);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);  // [2][G2_BN]
  uint8_t* transpose_s = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(bias_s + 2 * G2_BN) + 15) & ~uintptr_t(15));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  constexpr uint32_t TMEM_COLS = 512;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < G2_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);   // leader's producer arrives (expect_tx covers both CTAs' bytes)
      mbar_init(&empty_bar[s], 1);  // multicast commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * G2_EPI_THREADS);
    }
    mbar_init(mask_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // barriers of both CTAs initialised, TMEM allocated on both SMs
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.num_m_tiles * p.num_n_tiles * p.splits;  // in units of 256 x 256 pair tiles

  // Producer and issuer warps run their loops converged and one elected lane issues: from a divergent
  // single-thread branch the compiler wraps every TMA / tcgen05.mma instruction in a serialising loop.
  if (warp == 0) {
    // ============================== TMA producer (both CTAs) ==============================
    int s = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < total_tiles; tile += num_pairs) {
      const int ks = tile % p.splits;
      const int mn = tile / p.splits;
      const int m0 = (mn / p.num_n_tiles) * G2_BM + rank * 128;
      const int n0 = (mn % p.num_n_tiles) * G2_BN + rank * 128;
      const int kb0 = ks * p.kb_per_split;
      const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty_bar[s], phase ^ 1);
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * G2_STAGE_BYTES);
          uint8_t* a_dst = smA + s * G2_A_BYTES;
          uint8_t* b_dst = smB + s * G2_B_BYTES;
          if (!A_MN) {
            tma_load_2d_pair(&tmA, &full_bar[s], a_dst, kb * G2_BK, m0);
          } else {
#pragma unroll
            for (int blk = 0; blk < 2; ++blk)
              tma_load_2d_pair(&tmA, &full_bar[s], a_dst + blk * (G2_BK * 128), m0 + blk * 64, kb * G2_BK);
          }
          if (!B_MN) {
            tma_load_2d_pair(&tmB, &full_bar[s], b_dst, kb * G2_BK, n0);
          } else {
#pragma unroll
            for (int blk = 0; blk < 2; ++blk)
              tma_load_2d_pair(&tmB, &full_bar[s], b_dst + blk * (G2_BK * 128), n0 + blk * 64, kb * G2_BK);
          }
        }
        __syncwarp();
        if (++s == G2_STAGES) { s = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && leader) {
    // ============================== MMA issuer (leader CTA) ==============================
    constexpr uint32_t idesc = make_idesc_bf16(G2_BM, G2_BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int s = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = pair; tile < total_tiles; tile += num_pairs, ++it) {
      const int ks = tile % p.splits;
      const int kb0 = ks * p.kb_per_split;
      const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * G2_BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[s], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smA + s * G2_A_BYTES);
        const uint32_t b_addr = smem_u32(smB + s * G2_B_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < G2_BK / 16; ++k) {
            const uint64_t ad = A_MN ? make_smem_desc_sw128(a_addr + k * 2048, G2_BK * 128, 1024)
                                     : make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t bd = B_MN ? make_smem_desc_sw128(b_addr + k * 2048, G2_BK * 128, 1024)
                                     : make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            umma_bf16_pair(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_pair(&empty_bar[s]);                      // frees the stage in both CTAs
          if (kb == kb1 - 1) umma_commit_pair(&tfull_bar[acc]);  // accumulators complete in both CTAs
        }
        __syncwarp();
        if (++s == G2_STAGES) { s = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 2) {
    // ============================== epilogue (both CTAs, own 128 rows) ==============================
    const int quarter = warp & 3;
    const int chalf = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;
    const int row_in_tile = quarter * 32 + lane;
    const bool vec_ok = (p.out_dtype == ME_BF16) ? (p.ldd % 8 == 0) : (p.ldd % 4 == 0);
    constexpr int HALF = G2_BN / 2, CW = 32;
    int it = 0;
    for (int tile = pair; tile < total_tiles; tile += num_pairs, ++it) {
      const int ks = tile % p.splits;
      const int mn = tile / p.splits;
      const int m0 = (mn / p.num_n_tiles) * G2_BM + rank * 128;
      const int n0 = (mn % p.num_n_tiles) * G2_BN;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const bool first_split = (ks == 0);
      const bool use_bias = (p.flags & ME_EPI_BIAS) && first_split;
      float* bs = bias_s + acc * G2_BN;
      if (use_bias) {
        for (int c = et; c < G2_BN; c += G2_EPI_THREADS) bs[c] = (n0 + c < p.N) ? __ldg(p.bias + n0 + c) : 0.f;
      }
      const bool mask_tile = STAGED && (p.flags & ME_EPI_RELU_MASK);
      if (STAGED && et == 0) {
        bulk_wait_read_all();  // the previous tile's TMA stores have finished reading the staging tile
        if (mask_tile) {
          mbar_arrive_expect_tx(mask_bar, G2_STG_BYTES);
#pragma unroll
          for (int blk = 0; blk < 4; ++blk) tma_load_2d(&tmMask, mask_bar, stg + blk * 16384, n0 + blk * 64, m0);
        }
      }
      named_bar_sync(1, G2_EPI_THREADS);
      const int m = m0 + row_in_tile;
      const bool row_ok = m < p.M;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * G2_BN;
      mbar_wait(&tfull_bar[acc], acc_phase);
      if (mask_tile) mbar_wait(mask_bar, it & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = chalf * HALF; c0 < (chalf + 1) * HALF; c0 += CW) {
        if (n0 + c0 >= p.N) break;  // warp-uniform
        const int nb = n0 + c0;
        const bool full = nb + CW <= p.N;
        uint32_t r[CW];
        tmem_ld32(t_row + c0, r);
        float addv[CW];
        uint32_t maskw[CW / 2];
        const bool do_add = (p.flags & ME_EPI_ADD_F32) && first_split && row_ok;
        const bool do_mask = (p.flags & ME_EPI_RELU_MASK) && row_ok;
        if (STAGED) {
          // this thread's 64 bytes inside the swizzled staging tile: block c0/64, row, four 16-byte chunks
          uint8_t* srow = stg + (c0 >> 6) * 16384 + row_in_tile * 128;
          const int q0 = (c0 & 63) >> 3;
          if (mask_tile) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 u = *reinterpret_cast<const uint4*>(srow + (((q0 + j) ^ (row_in_tile & 7)) << 4));
              maskw[4 * j] = u.x; maskw[4 * j + 1] = u.y; maskw[4 * j + 2] = u.z; maskw[4 * j + 3] = u.w;
            }
          }
          tc_wait_ld();
          float bsl[CW], v[CW];
          if (use_bias) {
#pragma unroll
            for (int j = 0; j < CW; ++j) bsl[j] = bs[c0 + j];
          }
          gemm_epilogue_math<CW>(p, r, bsl, addv, maskw, use_bias, false, mask_tile, v);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 u;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
            *reinterpret_cast<uint4*>(srow + (((q0 + j) ^ (row_in_tile & 7)) << 4)) = u;
          }
        } else if (p.out_dtype == ME_F32 && p.splits == 1 && full && (p.ldd % 4 == 0)) {
          // fp32 rows (residual-stream gradients, weight gradients without split-K): full-sector stores
          float4 addc[8];
          gemm_addend_coalesced(p, (p.flags & ME_EPI_ADD_F32) != 0, m0 + quarter * 32, nb, lane, addc);
          gemm_epilogue_prefetch<CW>(p, addv, maskw, false, do_mask, m, nb, full);
          tc_wait_ld();
          float bsl[CW], v[CW];
          if (use_bias) {
#pragma unroll
            for (int j = 0; j < CW; ++j) bsl[j] = bs[c0 + j];
          }
          gemm_epilogue_math<CW>(p, r, bsl, addv, maskw, use_bias, false, do_mask, v);
          gemm_store_f32_coalesced(p, v, addc, m0 + quarter * 32, nb, transpose_s + (warp - 2) * 2048, lane);
        } else {
          gemm_epilogue_prefetch<CW>(p, addv, maskw, do_add, do_mask, m, nb, full);
          tc_wait_ld();
          if (row_ok) {
            float bsl[CW];
            if (use_bias) {
#pragma unroll
              for (int j = 0; j < CW; ++j) bsl[j] = bs[c0 + j];
            }
            gemm_epilogue_chunk<CW>(p, r, bsl, addv, maskw, use_bias, do_add, do_mask, m, nb, full, vec_ok);
          }
        }
      }
      __syncwarp();
      tc_fence_before();
      if (leader) mbar_arrive(&tempty_bar[acc]);
      else mbar_arrive_leader(&tempty_bar[acc]);
      if (STAGED) {
        fence_proxy_async_smem();
        named_bar_sync(1, G2_EPI_THREADS);  // the whole tile is in shared memory
        if (et == 0) {
#pragma unroll
          for (int blk = 0; blk < 4; ++blk)
            if (n0 + blk * 64 < p.N && m0 < p.M) tma_store_2d(&tmD, stg + blk * 16384, n0 + blk * 64, m0);
          bulk_commit();
        }
        if (p.colsum != nullptr && n0 + et < p.N) {
          // column sums of the tile as stored (bf16-rounded): the bias gradient of the Linear whose output gradient
          // this GEMM produces -- thread = column, rows straight out of the staging tile (a warp reads 64 contiguous
          // bytes of a row: no bank conflicts), one atomic per column and tile instead of a second pass over [M, N]
          const uint8_t* col = stg + (et >> 6) * 16384 + (et & 7) * 2;
          const int q = (et & 63) >> 3;
          const int rows = min(128, p.M - m0);
          float sum = 0.f;
#pragma unroll 8
          for (int r = 0; r < rows; ++r)
            sum += __bfloat162float(*reinterpret_cast<const bf16*>(col + r * 128 + ((q ^ (r & 7)) << 4)));
          atomicAdd(&p.colsum[n0 + et], sum);
        }
      }
    }
    if (STAGED && et == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <bool A_MN, bool B_MN, bool STAGED>
static int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD, const CUtensorMap& tmMask,
                       const GemmParams& p, int grid, cudaStream_t st) {
  auto kern = gemm_tc2_kernel<A_MN, B_MN, STAGED>;
  static bool configured = false;
  if (!configured) {
    ME_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM));
    configured = true;
  }
  kern<<<grid, G2_THREADS, G2_SMEM, st>>>(tmA, tmB, tmD, tmMask, p);
  ME_LAUNCH_CHECK();
  return 0;
}

int g_pair_force_direct = 0;  // test hook: 1 = never use the staged epilogue

// Returns 0 on success, 1 on error, -1 when the shape is better served by the 1-CTA kernel.
int launch_gemm_bf16_pair(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb, int ldd,
                          int a_mn, int b_mn, int out_dtype, int flags, const float* bias, const float* addend,
                          const void* relu_mask, int ldmask, int force_splits, cudaStream_t st, float* colsum_out,
                          bool* colsum_done) {
  if (colsum_done) *colsum_done = false;
  if (a_mn && !b_mn) return -1;
  const int sms = sm_count();
  const int pairs = sms / 2;
  const int num_m_tiles = (M + G2_BM - 1) / G2_BM;
  const int num_n_tiles = (N + G2_BN - 1) / G2_BN;
  const int num_kb = (K + G2_BK - 1) / G2_BK;
  const int tiles = num_m_tiles * num_n_tiles;
  int splits = 1;
  const bool linear = out_dtype == ME_F32 && !(flags & (ME_EPI_RELU | ME_EPI_RELU_MASK));
  const bool forced = force_splits != 0;
  if (force_splits > 0) splits = force_splits;
  else if (linear && num_kb >= 32 && tiles < pairs) {
    splits = pairs / tiles;
    if (splits > num_kb / 8) splits = num_kb / 8;
    if (splits < 1) splits = 1;
  }
  if (splits > 1 && !linear) return -1;
  // wide pair tiles only pay off when they fill most of the machine
  if (!forced && tiles * splits * 10 < pairs * 7) return -1;
  if (!forced && N < 192) return -1;
  const int kb_per = (num_kb + splits - 1) / splits;
  splits = (num_kb + kb_per - 1) / kb_per;

  CUtensorMap tmA, tmB;
  if (!a_mn) { if (make_tmap_2d_bf16(&tmA, A, K, M, lda, G2_BK, 128)) return 1; }
  else       { if (make_tmap_2d_bf16(&tmA, A, M, K, lda, 64, G2_BK)) return 1; }
  if (!b_mn) { if (make_tmap_2d_bf16(&tmB, B, K, N, ldb, G2_BK, 128)) return 1; }
  else       { if (make_tmap_2d_bf16(&tmB, B, N, K, ldb, 64, G2_BK)) return 1; }

  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.ldd = ldd; p.ldmask = ldmask; p.flags = flags; p.out_dtype = out_dtype;
  p.num_m_tiles = num_m_tiles; p.num_n_tiles = num_n_tiles; p.splits = splits; p.kb_per_split = kb_per;
  p.num_kb = num_kb; p.bias = bias; p.addend = addend; p.relu_mask = relu_mask; p.D = D;
  p.colsum = nullptr;
  if (splits > 1) ME_CUDA(cudaMemsetAsync(D, 0, static_cast<size_t>(M) * ldd * sizeof(float), st));
  const int total = tiles * splits;
  const int grid = 2 * (total < pairs ? total : pairs);
  // staged (TMA-store) epilogue: bf16 output without residual add / split-K and with 16-byte aligned rows
  bool staged = !a_mn && out_dtype == ME_BF16 && splits == 1 && ldd % 8 == 0 && !(flags & ME_EPI_ADD_F32) &&
                (reinterpret_cast<uintptr_t>(D) & 15) == 0;
  if ((flags & ME_EPI_RELU_MASK) && (ldmask % 8 != 0 || (reinterpret_cast<uintptr_t>(relu_mask) & 15) != 0)) staged = false;
  if (g_pair_force_direct) staged = false;
  CUtensorMap tmD = tmA, tmMask = tmA;  // placeholders when unused
  if (staged) {
    if (make_tmap_2d_bf16(&tmD, D, N, M, ldd, 64, 128)) return 1;
    if (flags & ME_EPI_RELU_MASK) {
      if (make_tmap_2d_bf16(&tmMask, relu_mask, N, M, ldmask, 64, 128)) return 1;
    }
  }
  if (staged && colsum_out != nullptr) {
    p.colsum = colsum_out;
    if (colsum_done) *colsum_done = true;
  }
  cudaEvent_t pe = prof_begin(2.0 * M * N * K, st);
  int rc;
  if (!a_mn && !b_mn) rc = staged ? launch_pair<false, false, true>(tmA, tmB, tmD, tmMask, p, grid, st)
                                  : launch_pair<false, false, false>(tmA, tmB, tmD, tmMask, p, grid, st);
  else if (!a_mn && b_mn) rc = staged ? launch_pair<false, true, true>(tmA, tmB, tmD, tmMask, p, grid, st)
                                      : launch_pair<false, true, false>(tmA, tmB, tmD, tmMask, p, grid, st);
  else rc = launch_pair<true, true, false>(tmA, tmB, tmD, tmMask, p, grid, st);
  prof_end(pe, st);
  return rc;
}

}  // namespace me
