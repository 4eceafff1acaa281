// Relative global attention on the 5th-generation tensor cores -- backward.
//
// One CTA owns a tile of 64 keys of one (batch, head) and walks the query tiles (128 rows) from the
// diagonal down.  Per step, with thread a == query row a == TMEM lane a:
//
//   MMA 1   [S | R] = Q [K ; Eband]^T   [128 x (64 + 192)], one N = 256 instruction per 16 head dims (K and the
//           band of E sit back to back in shared memory; same band as the forward kernel)
//           dP = dO V^T                 [128 x 64]
//   threads x  = S + skew(R)                    Srel[a, b] = R[a, 127 - a + b]
//           P  = exp2(x c - lse)                dS = P (dP - D) / sqrt(dh)
//           P, dS -> shared memory (bf16, UMMA layouts);  dSb = dS in band coordinates
//                                               dSb[a, 127 - a + b] = dS[a, b]  ("unskew" = a shifted store)
//   MMA 2   dV += P^T dO        dK += dS^T Q    dQ_tile = dS K + dSb Eband      dE_tile = dSb^T Q  [192 x dh]
//   threads dQ_tile and dE_tile: TMEM -> shared memory; the control warp then issues fp32 reduce-adds into
//           global memory on the TMA unit (cp.reduce.async.bulk), so no thread ever issues an atomic and the
//           compute warps never wait on a bulk group (mbarrier handshakes stg_full / stg_free / p_free).
//
// dK / dV stay in TMEM for the whole CTA and are written once.  Every MMA runs with M = 128.  dK (64 keys)
// and rows 128..191 of the dE tile share one MMA, [dS | dSb_hi]^T Q: lanes 0..63 of its accumulator keep
// summing dK over the steps, lanes 64..127 hold this step's dE rows and are read out and zeroed (tcgen05.st)
// by their threads every step.  P^T has only 64 valid rows: the other accumulator lanes of dV are never read.
#include "attention_tc.cuh"

namespace me {

constexpr int FB_BM = 128;            // query rows per step
constexpr int FB_BN = 64;             // keys per CTA
constexpr int FB_EROWS = 192;
constexpr int FB_COMPUTE_THREADS = 256;  // 8 warps: warp w and w+4 share TMEM lanes 32*(w&3).. and split the key columns
constexpr int FB_CONTROL_WARP = 8;       // TMA producer + MMA issuer, TMEM alloc
constexpr int FB_REDUCE_WARP = 9;        // issues the fp32 reduce-adds of the staged dQ / dE tiles
constexpr int FB_LOAD_WARP = 10;         // TMA loads
constexpr int FB_THREADS = FB_COMPUTE_THREADS + 96;
constexpr int FB_DE_COPIES = 32;      // private dE accumulators: concurrently running CTAs walk the same bands of E
                                      // in lockstep, and same-address reduce-adds serialise in the L2 slices
constexpr int FB_STG_MAX = 256;       // staging row pitch for dh = 64: dh floats, 16-byte chunks XOR-swizzled by the row
constexpr int FB_OFF_K = 0;
constexpr int FB_OFF_E = FB_OFF_K + 8192;   // E band directly behind K: [K ; Eband] is one 256-row B operand
constexpr int FB_OFF_V = FB_OFF_E + 24576;
constexpr int FB_OFF_Q = FB_OFF_V + 8192;
constexpr int FB_OFF_DO = FB_OFF_Q + 16384;
constexpr int FB_OFF_P = FB_OFF_DO + 16384;
constexpr int FB_OFF_DS = FB_OFF_P + 16384;
constexpr int FB_OFF_DSB = FB_OFF_DS + 16384;
constexpr int FB_OFF_STG0 = FB_OFF_DSB + 49152;
constexpr int FB_OFF_STG1 = FB_OFF_STG0 + 128 * FB_STG_MAX;
constexpr int FB_OFF_BAR = FB_OFF_STG1 + 128 * FB_STG_MAX;
constexpr int FB_SMEM = FB_OFF_BAR + 128;
static_assert(FB_OFF_STG0 % 1024 == 0 && FB_OFF_BAR % 1024 == 0, "tile alignment");
static_assert(FB_SMEM <= 227 * 1024, "shared memory budget");
constexpr uint32_t FB_TMEM_COLS = 512;
constexpr uint32_t FB_COL_S = 0, FB_COL_R = 64, FB_COL_DP = 256, FB_COL_DK = 320, FB_COL_DV = 384, FB_COL_DQ = 448;
constexpr uint32_t FB_COL_DE_LO = 64;  // rows 0..127 of the dE tile alias R once it has been consumed; rows 128..191
                                       // are lanes 64..127 of the dK columns (one MMA: [dS | dSb_hi]^T Q)

struct FbParams {
  int B, H, L, max_seq;
  int64_t k_sb, k_sh, k_sj, v_sb, v_sh, v_sj, keypad_ld;
  const uint8_t* keypad;
  const float* lse;
  const float* dsum;
  float* dq_ws;  // fp32 [B, H, L, dh]: dq accumulated across key tiles (chunks swizzled like the staging rows)
  float* dE_ws;  // fp32 [FB_DE_COPIES, max_seq, dh]
  bf16* dk;
  bf16* dv;
  float scale_log2, scale;
  int noncausal;     // ME_ATTN_NONCAUSAL: every query tile, every key < L visible
  long long* trace;  // debugging: per-phase clock64() stamps of one CTA (me_debug_trace_set), else NULL
};
static long long* g_attn_bwd_trace = nullptr;
// stamps: [role (0 = warp 0, 1 = warp 7, 2 = control warp)][step < 16][event < 16].  Compiled in only with
// -DME_ATTN_BWD_TRACE (ME_TRACE=1 python -m midi_emotion_b200.build): the predicated stamps cost registers.
#ifdef ME_ATTN_BWD_TRACE
#define FB_TRACE(role, st, k)                                                                  \
  do {                                                                                         \
    if (tr && (st) < 16) p.trace[((role) * 16 + (st)) * 16 + (k)] = clock64();                 \
  } while (0)
#else
#define FB_TRACE(role, st, k) do { } while (0)
#endif

// TMEM -> shared staging: NCOLS (multiple of 8) accumulator columns of this thread's lane
template <int NCOLS>
__device__ __forceinline__ void stage_row(uint32_t taddr, float* dst) {
  static_assert(NCOLS % 8 == 0, "column groups of 8");
  uint32_t v[NCOLS];
#pragma unroll
  for (int c0 = 0; c0 < NCOLS; c0 += 8) {
    uint32_t t[8];
    tmem_ld8(taddr + c0, t);
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c0 + c] = t[c];
  }
  tc_wait_ld();
#pragma unroll
  for (int c = 0; c < NCOLS; c += 4)
    *reinterpret_cast<uint4*>(dst + c) = make_uint4(v[c], v[c + 1], v[c + 2], v[c + 3]);
}

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
template <int NCOLS>
__device__ __forceinline__ void tmem_zero_cols(uint32_t taddr) {  // no wait (tcgen05.wait::st)
#pragma unroll
  for (int c0 = 0; c0 < NCOLS; c0 += 8)
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr + c0), "r"(0u)
                 : "memory");
}
// chunk-swizzle key of the staging rows: 8 rows when the row is a multiple of 128 bytes, else 4
__host__ __device__ constexpr int fb_swizzle_mask(int dh) { return dh % 32 == 0 ? 7 : 3; }
// NCOLS floats of row r, starting at 16-byte chunk `chunk0` of the (unswizzled) row
template <int NCOLS, int SWZ>
__device__ __forceinline__ void sts_row_swz(uint8_t* row, int r, int chunk0, const uint32_t (&v)[NCOLS]) {
#pragma unroll
  for (int c = 0; c < NCOLS / 4; ++c)
    *reinterpret_cast<uint4*>(row + (((chunk0 + c) ^ (r & SWZ)) << 4)) =
        make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}

template <int DH>
__global__ void __launch_bounds__(FB_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                   const __grid_constant__ CUtensorMap tmE, FbParams p) {
  // Staging / workspace rows are dh floats with no padding; the 16-byte chunk c of row r sits at position
  // c ^ (r & SWZ) (conflict-free 16-byte stores by eight consecutive rows).  Tiles start at multiples of 64 rows
  // in the global workspaces, so the key is the same function of the global row (the finish kernel undoes it).
  constexpr int STG = DH * 4;
  constexpr int SWZ = fb_swizzle_mask(DH);
  constexpr int HC = DH / 2;          // accumulator columns handled by each of the two threads of a row
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
  uint8_t* stg2 = sP;  // P / dS are free between MMA 2 and the next step
  uint64_t* bars = reinterpret_cast<uint64_t*>(fb_smem + FB_OFF_BAR);
  uint64_t* kv_full = bars + 0;
  uint64_t* q_full = bars + 1;    // per step loads: Q, dO and the band of E on their own barriers
  uint64_t* do_full = bars + 2;
  uint64_t* e_full = bars + 3;
  uint64_t* m1_done = bars + 4;   // S, R ready
  uint64_t* a_done = bars + 5;    // P, dS, dSb written (256 arrivals)
  uint64_t* dp_done = bars + 6;   // dP ready
  uint64_t* q_free = bars + 7;    // dK accumulated, dE tile ready: Q is free
  uint64_t* e_free = bars + 8;    // the relative part of dQ is done: the band of E is free
  uint64_t* m2_done = bars + 9;   // dV accumulated, dQ tile ready: every MMA of the step has retired (dO, P free)
  uint64_t* b_done = bars + 10;   // dQ tile (and the dE rows in the dK columns) read out of TMEM (256 arrivals)
  uint64_t* stg_full = bars + 11; // dQ / dE tiles staged in shared memory (256 arrivals)
  uint64_t* stg_free = bars + 12; // the reduce-adds of the step have read the staging buffers
  uint64_t* p_free = bars + 13;   // ... the part of them that aliases P
  uint64_t* de_read = bars + 14;  // dE rows 0..127 (aliasing R) read out of TMEM (256 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int j0 = kt * FB_BN;
  const int nq = (p.L + FB_BM - 1) / FB_BM;
  const int qi0 = p.noncausal ? 0 : j0 / FB_BM;
  const int nsteps = nq - qi0;
  // consecutive block ids (the CTAs resident at the same time) accumulate dE into different copies
  float* const dE_mine =
      p.dE_ws + static_cast<int64_t>((blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) % FB_DE_COPIES) *
                    p.max_seq * DH;

  const bool tr = p.trace != nullptr && kt == 2 && h == p.H / 2 && b == p.B / 2 && lane == 0 &&
                  (warp == 0 || warp == 7 || warp == FB_CONTROL_WARP);  // (reduce warp: not traced)
  const int trole = warp == 0 ? 0 : (warp == 7 ? 1 : 2);

  if (tid == 0) {
    if ((smem_u32(fb_smem) & 1023u) != 0) __trap();
    mbar_init(kv_full, 1);
    mbar_init(q_full, 1);
    mbar_init(do_full, 1);
    mbar_init(e_full, 1);
    mbar_init(m1_done, 1);
    mbar_init(a_done, FB_COMPUTE_THREADS);
    mbar_init(dp_done, 1);
    mbar_init(q_free, 1);
    mbar_init(e_free, 1);
    mbar_init(m2_done, 1);
    mbar_init(b_done, FB_COMPUTE_THREADS);
    mbar_init(de_read, FB_COMPUTE_THREADS);
    mbar_init(stg_full, FB_COMPUTE_THREADS);
    mbar_init(stg_free, 1);
    mbar_init(p_free, 1);
    fence_mbar_init();
  }
  // dSb starts as zeros; every step rewrites only the chunks around each row's window
  {
    uint4* z = reinterpret_cast<uint4*>(sdSb);
    for (int c = tid; c < 49152 / 16; c += FB_THREADS) z[c] = make_uint4(0, 0, 0, 0);
  }
  if (warp == FB_CONTROL_WARP) tmem_alloc(tmem_slot, FB_TMEM_COLS);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == FB_CONTROL_WARP) {
    // ======================= MMA issuer =======================
    // The whole warp runs this loop converged and one elected lane issues: issued from a divergent
    // `if (lane == 0)` every tcgen05.mma is wrapped by the compiler in a serialising loop (~90 cycles per
    // instruction, three times the execution time of an N = 64 MMA).  Issue is not fire-and-forget either:
    // it blocks once a few MMAs are queued, so nothing else (loads, bulk waits) lives in this warp.
    // Order inside a step (the tensor pipe executes in issue order), chosen so that the per-step operands
    // are released early and their reloads hide behind the remaining MMAs:
    //   [S | R](st)                      needs Q, E of the step and the TMEM tiles of step st-1 read out
    //   dP(st)                           needs dO; the threads pick it up after they have formed P
    //   dQ(st)  = dSb Eband          ->  E free
    //   dK(st), dE(st)               ->  Q free, dE tile ready
    //   dV(st), dQ(st) += dS K       ->  dO, P free, dQ tile ready (K is resident)
    constexpr uint32_t idesc_s = make_idesc_bf16(128, FB_BN, 0, 0);     // dP : K-major x K-major
    constexpr uint32_t idesc_sr = make_idesc_bf16(128, FB_BN + FB_EROWS, 0, 0);  // [S | R]
    constexpr uint32_t idesc_tt = make_idesc_bf16(128, DH, 1, 1);       // dV, dK, dE : A^T (MN-major) x B (MN-major)
    constexpr uint32_t idesc_nt = make_idesc_bf16(128, DH, 0, 1);       // dQ : A (K-major) x B (MN-major)
    const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), q_addr = smem_u32(sQ), do_addr = smem_u32(sdO);
    const uint32_t e_addr = smem_u32(sE), p_addr = smem_u32(sP), ds_addr = smem_u32(sdS), dsb_addr = smem_u32(sdSb);
    auto issue_dp = [&]() {
#pragma unroll
      for (int k = 0; k < DH / 16; ++k)
        umma_bf16(tmem_base + FB_COL_DP, make_smem_desc_sw128(do_addr + k * 32, 16, 1024),
                  make_smem_desc_sw128(v_addr + k * 32, 16, 1024), idesc_s, k > 0);
    };
    // [S | R] of step st only overwrites TMEM that the threads release early (S, R in phase A; the dE rows
    // aliasing R at q_free time), so it is queued directly behind the MMAs of step st-1: the tensor pipe
    // never waits for the threads to drain the dQ tile.
    auto issue_sr = [&](int st) {
      mbar_wait(q_full, st & 1);
      mbar_wait(e_full, st & 1);
      if (st > 0) mbar_wait(de_read, (st - 1) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)  // [S | R] = Q [K ; Eband]^T
          umma_bf16(tmem_base + FB_COL_S, make_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                    make_smem_desc_sw128(k_addr + k * 32, 16, 1024), idesc_sr, k > 0);
        umma_commit(m1_done);
      }
      __syncwarp();
      mbar_wait(do_full, st & 1);
      tc_fence_after();
      if (elect_one()) {
        issue_dp();  // (its TMEM columns were consumed before a_done of the previous step)
        umma_commit(dp_done);
      }
      __syncwarp();
    };
    mbar_wait(kv_full, 0);
    issue_sr(0);
    for (int st = 0; st < nsteps; ++st) {
      const uint32_t ph = st & 1;
      FB_TRACE(2, st, 3);
      mbar_wait(a_done, ph);
      if (st > 0) mbar_wait(b_done, (st - 1) & 1);  // the dQ tile and the dE rows in the dK columns were read out
      tc_fence_after();
      FB_TRACE(2, st, 4);
      const uint32_t acc0 = st > 0 ? 1u : 0u;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 12; ++k)  // dQ_tile = dSb Eband
          umma_bf16(tmem_base + FB_COL_DQ,
                    make_smem_desc_sw128(dsb_addr + ((k >> 2) == 2 ? 0 : (k >> 2) + 1) * 16384 + (k & 3) * 32, 16, 1024),
                    make_smem_desc_sw128(e_addr + k * 2048, 8192, 1024), idesc_nt, k > 0);
        umma_commit(e_free);
#pragma unroll
        for (int k = 0; k < 8; ++k)  // lanes 0..63: dK += dS^T Q;  lanes 64..127 (zeroed): dE_tile[128:192] = dSb_hi^T Q
          umma_bf16(tmem_base + FB_COL_DK, make_smem_desc_sw128(ds_addr + k * 2048, 16384, 1024),
                    make_smem_desc_sw128(q_addr + k * 2048, 8192, 1024), idesc_tt, (k > 0) ? 1u : acc0);
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dE_tile[0:128] = dSb[:, 0:128]^T Q
          umma_bf16(tmem_base + FB_COL_DE_LO, make_smem_desc_sw128(dsb_addr + 16384 + k * 2048, 16384, 1024),
                    make_smem_desc_sw128(q_addr + k * 2048, 8192, 1024), idesc_tt, k > 0);
        umma_commit(q_free);
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dV += P^T dO
          umma_bf16(tmem_base + FB_COL_DV, make_smem_desc_sw128(p_addr + k * 2048, 16384, 1024),
                    make_smem_desc_sw128(do_addr + k * 2048, 8192, 1024), idesc_tt, (k > 0) ? 1u : acc0);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // dQ_tile += dS K
          umma_bf16(tmem_base + FB_COL_DQ, make_smem_desc_sw128(ds_addr + k * 32, 16, 1024),
                    make_smem_desc_sw128(k_addr + k * 2048, 8192, 1024), idesc_nt, 1u);
        umma_commit(m2_done);
      }
      __syncwarp();
      FB_TRACE(2, st, 5);
      if (st + 1 < nsteps) issue_sr(st + 1);
      FB_TRACE(2, st, 6);
    }
  } else if (warp == FB_LOAD_WARP) {
    // ======================= TMA loads =======================
    auto load_q = [&](int st) {
      mbar_arrive_expect_tx(q_full, 16384);
      tma_load_4d(&tmQ, q_full, sQ, 0, h, (qi0 + st) * FB_BM, b);
    };
    auto load_do = [&](int st) {
      mbar_arrive_expect_tx(do_full, 16384);
      tma_load_4d(&tmdO, do_full, sdO, 0, h, (qi0 + st) * FB_BM, b);
    };
    auto load_e = [&](int st) {
      mbar_arrive_expect_tx(e_full, 24576);
      tma_load_2d(&tmE, e_full, sE, 0, p.max_seq - FB_BM - ((qi0 + st) * FB_BM - j0));
    };
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      tma_prefetch_desc(&tmdO);
      tma_prefetch_desc(&tmE);
      mbar_arrive_expect_tx(kv_full, 16384);
      tma_load_4d(&tmK, kv_full, sK, 0, h, j0, b);
      tma_load_4d(&tmV, kv_full, sV, 0, h, j0, b);
      load_do(0);
      load_q(0);
      load_e(0);
    }
    __syncwarp();
    for (int st = 0; st + 1 < nsteps; ++st) {
      const uint32_t ph = st & 1;
      mbar_wait(e_free, ph);
      if (elect_one()) load_e(st + 1);
      __syncwarp();
      mbar_wait(q_free, ph);
      if (elect_one()) load_q(st + 1);
      __syncwarp();
      mbar_wait(m2_done, ph);
      if (elect_one()) load_do(st + 1);
      __syncwarp();
    }
  } else if (warp == FB_REDUCE_WARP) {
    // ======================= reduce-add issuer =======================
    // fp32 reduce-add of the staged tiles into global memory by the TMA unit; the part that aliases P goes
    // first, as its own bulk group, so that P may be rewritten early.  Waiting on bulk groups happens here,
    // never in the compute warps or the MMA issuer.
    for (int st = 0; st < nsteps; ++st) {
      mbar_wait(stg_full, st & 1);
      // The L2 executes about 30 B/cycle/SM of fp32 reduce-adds and the TMA unit is a FIFO: loads issued
      // behind a step's reduce-adds would wait ~3000 cycles, so the next step's dO load (the last of its
      // loads) has to land before they are queued.
      if (st + 1 < nsteps) mbar_wait(do_full, (st + 1) & 1);
      if (elect_one()) {
        const int i0 = (qi0 + st) * FB_BM;
        const int e0 = p.max_seq - FB_BM - (i0 - j0);
        const int n2 = min(64, p.max_seq - (e0 + 128));
        if (n2 > 0) bulk_reduce_add_f32(dE_mine + static_cast<int64_t>(e0 + 128) * DH, stg2, n2 * STG);
        bulk_commit();
        const int nq_rows = min(128, p.L - i0);
        bulk_reduce_add_f32(p.dq_ws + ((static_cast<int64_t>(b) * p.H + h) * p.L + i0) * DH, stg0, nq_rows * STG);
        const int n1 = min(128, p.max_seq - e0);
        if (n1 > 0) bulk_reduce_add_f32(dE_mine + static_cast<int64_t>(e0) * DH, stg1, n1 * STG);
        bulk_commit();
        bulk_wait_read_1();
        mbar_arrive(p_free);
        bulk_wait_read_all();
        mbar_arrive(stg_free);
      }
      __syncwarp();
    }
    if (elect_one()) bulk_wait_all();
    __syncwarp();
  } else {
    // ================================ compute warps ================================
    const int half = warp >> 2, quarter = warp & 3;
    const int a = quarter * 32 + lane;   // query row inside the tile == TMEM lane
    const int shift = 31 - lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint8_t* kp = p.keypad ? p.keypad + static_cast<int64_t>(b) * p.keypad_ld : nullptr;
    uint32_t kpm = 0;  // key-pad bits of this thread's 32 keys
    if (kp || p.noncausal) {  // (keys past the sequence only matter without the causal predicate)
      const int j = j0 + 32 * half + lane;
      kpm = __ballot_sync(0xffffffffu, j >= p.L || (kp && kp[j] != 0));
    }
    const float cs = p.scale_log2;
    // this thread's 32 values sit at band columns c = 127 - a + 32*half + bb
    const int win_base = 127 - a + 32 * half;
    const int win_q0 = win_base >> 3, win_o = win_base & 7;

    for (int st = 0; st < nsteps; ++st) {
      const uint32_t ph = st & 1;
      const int i0 = (qi0 + st) * FB_BM;
      const int i = i0 + a;
      const bool row_ok = i < p.L;
      const int64_t stat = (static_cast<int64_t>(b) * p.H + h) * p.L + i;
      float l_nat = -INFINITY, Di = 0.f;
      if (row_ok) {
        l_nat = p.lse[stat];
        Di = p.dsum[stat];
      }
      const int lim = p.noncausal ? 31 : i - j0 - 32 * half;  // this thread's columns bb <= lim are causal-visible
      uint32_t vm = lim >= 31 ? 0xffffffffu : (lim < 0 ? 0u : ((2u << lim) - 1u));
      vm &= ~kpm;

      FB_TRACE(trole, st, 0);
      mbar_wait(m1_done, ph);
      tc_fence_after();
      FB_TRACE(trole, st, 1);
      const float lse2 = (!row_ok || l_nat == -INFINITY) ? INFINITY : l_nat * 1.4426950408889634f;

      uint32_t pw[16], dw[20];  // bf16x2 words of P[a, 32h..] and dS[a, 32h..] (+ 4 zero words for the band shift)
      {
        float pe[32];
        {
          uint32_t sv[32], rv[64];
          tmem_ld32(t_lane + FB_COL_S + 32 * half, sv);
          tmem_ld64(t_lane + FB_COL_R + 96 - 32 * quarter + 32 * half, rv);
          tc_wait_ld();
          skew_select(rv, shift);
          // only the tiles on the diagonal (and rows past the sequence / pad keys) need the mask: a warp-uniform
          // branch keeps 96 ALU-pipe instructions out of the common case
          if (__all_sync(0xffffffffu, vm == 0xffffffffu)) {
#pragma unroll
            for (int bb = 0; bb < 32; ++bb)
              pe[bb] = fast_exp2(fmaf(__uint_as_float(sv[bb]) + __uint_as_float(rv[bb]), cs, -lse2));
          } else {
#pragma unroll
            for (int bb = 0; bb < 32; ++bb) {
              const float x = __uint_as_float(sv[bb]) + __uint_as_float(rv[bb]);
              const float e = fast_exp2(fmaf(x, cs, -lse2));
              pe[bb] = ((vm >> bb) & 1u) ? e : 0.f;
            }
          }
#pragma unroll
          for (int bb = 0; bb < 32; bb += 2) {
            __nv_bfloat162 ph2 = __floats2bfloat162_rn(pe[bb], pe[bb + 1]);
            pw[bb / 2] = *reinterpret_cast<uint32_t*>(&ph2);
          }
        }
        mbar_wait(dp_done, ph);  // dP was issued behind [S | R]: it lands while P is being formed
        tc_fence_after();
        {
          uint32_t dpv[32];
          tmem_ld32(t_lane + FB_COL_DP + 32 * half, dpv);
          tc_wait_ld();
          const float nds = -Di * p.scale;
#pragma unroll
          for (int bb = 0; bb < 32; bb += 2) {
            const float d0 = pe[bb] * fmaf(__uint_as_float(dpv[bb]), p.scale, nds);
            const float d1 = pe[bb + 1] * fmaf(__uint_as_float(dpv[bb + 1]), p.scale, nds);
            __nv_bfloat162 dh2 = __floats2bfloat162_rn(d0, d1);
            dw[bb / 2] = *reinterpret_cast<uint32_t*>(&dh2);
          }
        }
      }
      FB_TRACE(trole, st, 2);
      // the previous step's reduce-add must have left the staging rows that alias P / dS
      if (st > 0) mbar_wait(p_free, (st - 1) & 1);
      FB_TRACE(trole, st, 4);
      // P and dS rows: UMMA SWIZZLE_128B rows of 128 B (chunk kc of row a at position kc ^ (a & 7))
      {
        uint8_t* prow = sP + a * 128;
        uint8_t* drow = sdS + a * 128;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const int pos = (((4 * half + n) ^ (a & 7))) << 4;
          *reinterpret_cast<uint4*>(prow + pos) = make_uint4(pw[4 * n], pw[4 * n + 1], pw[4 * n + 2], pw[4 * n + 3]);
          *reinterpret_cast<uint4*>(drow + pos) = make_uint4(dw[4 * n], dw[4 * n + 1], dw[4 * n + 2], dw[4 * n + 3]);
        }
      }
      FB_TRACE(trole, st, 12);
      // dS in band coordinates: shift right by win_o (0..7) elements inside a 40-element span.  Chunk 4 of
      // the lower half and chunk 0 of the upper half are the same 16 bytes: both sides write only their own
      // elements there (2-byte stores), every other chunk is written whole.
      {
        dw[16] = dw[17] = dw[18] = dw[19] = 0u;
        const uint32_t on4 = win_o & 4, on2 = win_o & 2;
#pragma unroll
        for (int w = 19; w >= 0; --w) dw[w] = sel_b32(w >= 2 ? dw[w - 2] : 0u, dw[w], on4);
#pragma unroll
        for (int w = 19; w >= 0; --w) dw[w] = sel_b32(w >= 1 ? dw[w - 1] : 0u, dw[w], on2);
        const uint32_t hs = (win_o & 1) ? 16u : 0u;
#pragma unroll
        for (int w = 19; w >= 0; --w) dw[w] = __funnelshift_l(w >= 1 ? dw[w - 1] : 0u, dw[w], hs);
        auto chunk_ptr = [&](int q) -> uint8_t* {
          const int panel = q >> 3;  // 64 band columns each, stored in the order [2][0][1] behind dS
          return sdSb + (panel == 2 ? 0 : panel + 1) * 16384 + a * 128 + (((q & 7) ^ (a & 7)) << 4);
        };
        const int shared_n = half == 0 ? 4 : 0;  // this span's chunk that is shared with the other half
#pragma unroll
        for (int n = 0; n < 5; ++n) {
          uint8_t* dst = chunk_ptr(win_q0 + n);
          if (n != shared_n) {
            *reinterpret_cast<uint4*>(dst) = make_uint4(dw[4 * n], dw[4 * n + 1], dw[4 * n + 2], dw[4 * n + 3]);
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const bool mine = half == 0 ? (e < win_o) : (e >= win_o);
              const uint32_t word = dw[4 * n + e / 2];
              const uint16_t val = static_cast<uint16_t>((e & 1) ? (word >> 16) : (word & 0xFFFFu));
              if (mine) *reinterpret_cast<uint16_t*>(dst + 2 * e) = val;
            }
          }
        }
      }
      FB_TRACE(trole, st, 13);
      fence_proxy_async_smem();
      FB_TRACE(trole, st, 14);
      tc_fence_before();
      mbar_arrive(a_done);
      FB_TRACE(trole, st, 5);

      // dE tile, then dQ tile: TMEM -> registers -> fp32 rows in shared memory; the reduce warp issues one TMA
      // reduce-add per tile
      mbar_wait(q_free, ph);   // dK accumulated, dE tile ready
      tc_fence_after();
      if (st > 0) mbar_wait(stg_free, (st - 1) & 1);
      FB_TRACE(trole, st, 6);
      uint32_t hi[HC];  // rows 128..191 of the dE tile wait in registers: their staging rows alias P (dV reads it)
      {
        uint32_t lo[HC];
        tmem_ld_cols<HC>(t_lane + FB_COL_DE_LO + half * HC, lo);
        if (quarter >= 2) tmem_ld_cols<HC>(t_lane + FB_COL_DK + half * HC, hi);
        tc_wait_ld();
        tc_fence_before();
        mbar_arrive(de_read);   // the next [S | R] may overwrite the R columns
        if (quarter >= 2) tmem_zero_cols<HC>(t_lane + FB_COL_DK + half * HC);  // next step accumulates onto zeros
        sts_row_swz<HC, SWZ>(stg1 + a * STG, a, half * (HC / 4), lo);
      }
      FB_TRACE(trole, st, 7);
      mbar_wait(m2_done, ph);  // dV accumulated, dQ tile ready; every MMA of the step has retired
      tc_fence_after();
      FB_TRACE(trole, st, 8);
      {
        // release the dQ columns before the (slower) staging stores
        uint32_t dq[HC];
        tmem_ld_cols<HC>(t_lane + FB_COL_DQ + half * HC, dq);
        tc_wait_ld();
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(b_done);
        FB_TRACE(trole, st, 9);
        sts_row_swz<HC, SWZ>(stg0 + a * STG, a, half * (HC / 4), dq);
        if (quarter >= 2) sts_row_swz<HC, SWZ>(stg2 + (a - 64) * STG, a, half * (HC / 4), hi);
      }
      FB_TRACE(trole, st, 10);
      fence_proxy_async_smem();
      mbar_arrive(stg_full);  // the reduce warp issues the reduce-adds
      FB_TRACE(trole, st, 11);
    }

    // dK / dV: rows 0..63 of the accumulators (TMEM lanes 0..63: quarters 0 and 1), columns split by half
    // (every thread has waited for m2_done of the last step: dK / dV are complete)
    if (quarter < 2) {
      const int j = j0 + a;
      const bool key_ok = j < p.L;
      bf16* dkrow = p.dk + static_cast<int64_t>(b) * p.k_sb + static_cast<int64_t>(j) * p.k_sj + h * p.k_sh + half * HC;
      bf16* dvrow = p.dv + static_cast<int64_t>(b) * p.v_sb + static_cast<int64_t>(j) * p.v_sj + h * p.v_sh + half * HC;
#pragma unroll
      for (int c0 = 0; c0 < HC; c0 += 8) {
        uint32_t vk[8], vv[8];
        tmem_ld8(t_lane + FB_COL_DK + half * HC + c0, vk);
        tmem_ld8(t_lane + FB_COL_DV + half * HC + c0, vv);
        tc_wait_ld();
        if (key_ok) {
          uint4 uk, uv;
          __nv_bfloat162* hk = reinterpret_cast<__nv_bfloat162*>(&uk);
          __nv_bfloat162* hv = reinterpret_cast<__nv_bfloat162*>(&uv);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            hk[e] = __floats2bfloat162_rn(__uint_as_float(vk[2 * e]), __uint_as_float(vk[2 * e + 1]));
            hv[e] = __floats2bfloat162_rn(__uint_as_float(vv[2 * e]), __uint_as_float(vv[2 * e + 1]));
          }
          *reinterpret_cast<uint4*>(dkrow + c0) = uk;
          *reinterpret_cast<uint4*>(dvrow + c0) = uv;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == FB_CONTROL_WARP) tmem_dealloc(tmem_base, FB_TMEM_COLS);
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

// dq (bf16, strided) = dq_ws (fp32 [B, H, L, dh]);  dE[e, :] += sum over copies of dE_ws[., e, :]
// (both with the staging rows' chunk swizzle)
__global__ void attn_bwd_finish_kernel(const float* __restrict__ dq_ws, bf16* __restrict__ dq, int64_t q_sb,
                                       int64_t q_si, int64_t q_sh, int B, int H, int L, int dh,
                                       const float* __restrict__ dE_ws, float* __restrict__ dE, int max_seq,
                                       int de_blocks) {
  const int q4 = dh / 4;
  const int swz = fb_swizzle_mask(dh);
  if (static_cast<int>(blockIdx.x) < de_blocks) {
    // the first blocks fold the private dE copies (32 independent loads per thread, issued together)
    const int64_t n_de = static_cast<int64_t>(max_seq) * q4;
    for (int64_t u = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; u < n_de;
         u += static_cast<int64_t>(de_blocks) * blockDim.x) {
      const int c = static_cast<int>(u % q4) * 4;
      const int64_t e = u / q4;
      const float* src = dE_ws + e * dh + (((c >> 2) ^ (static_cast<int>(e) & swz)) << 2);
      float4 v[FB_DE_COPIES];
#pragma unroll
      for (int cp = 0; cp < FB_DE_COPIES; ++cp)
        v[cp] = *reinterpret_cast<const float4*>(src + static_cast<int64_t>(cp) * max_seq * dh);
      float4 o = *reinterpret_cast<float4*>(dE + e * dh + c);
#pragma unroll
      for (int cp = 0; cp < FB_DE_COPIES; ++cp) { o.x += v[cp].x; o.y += v[cp].y; o.z += v[cp].z; o.w += v[cp].w; }
      *reinterpret_cast<float4*>(dE + e * dh + c) = o;
    }
    return;
  }
  const int64_t n_dq = static_cast<int64_t>(B) * H * L * q4;
  const int64_t stride = static_cast<int64_t>(gridDim.x - de_blocks) * blockDim.x;
  for (int64_t t = static_cast<int64_t>(blockIdx.x - de_blocks) * blockDim.x + threadIdx.x; t < n_dq; t += stride) {
    const int c = static_cast<int>(t % q4) * 4;
    const int64_t row = t / q4;  // (b*H + h)*L + i
    const int i = static_cast<int>(row % L);
    const int64_t bh = row / L;
    const int h = static_cast<int>(bh % H);
    const int64_t b = bh / H;
    const float4 v = *reinterpret_cast<const float4*>(dq_ws + row * dh + (((c >> 2) ^ (i & swz)) << 2));
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(dq + b * q_sb + i * q_si + h * q_sh + c) = u;
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
  ME_CHECK(a->max_seq % 128 == 0 && a->Lq <= a->max_seq, "me_attention_backward: max_seq must be a multiple of 128 and >= L");
  ME_CHECK((reinterpret_cast<uintptr_t>(ba->dq_acc) & 15) == 0, "me_attention_backward: dq_acc must be 16-byte aligned");
  ME_CHECK(a->q_sh == a->dh && a->k_sh == a->dh && a->v_sh == a->dh,
           "me_attention_backward: ME_ATTN_TENSOR expects heads packed along the feature axis (stride dh)");
  const int B = a->B, H = a->H, L = a->Lq, dh = a->dh;
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
    ME_CUDA(cudaMemsetAsync(ba->dq_acc, 0, sizeof(float) * me_attention_backward_workspace_floats(B, H, L, dh, a->max_seq), st));
  }
  FbParams p;
  p.B = B; p.H = H; p.L = L; p.max_seq = a->max_seq;
  p.k_sb = a->k_sb; p.k_sh = a->k_sh; p.k_sj = a->k_sj;
  p.v_sb = a->v_sb; p.v_sh = a->v_sh; p.v_sj = a->v_sj;
  p.keypad_ld = a->keypad_ld; p.keypad = a->keypad;
  p.lse = a->lse; p.dsum = ba->dsum;
  p.dq_ws = ba->dq_acc;
  p.dE_ws = ba->dq_acc + static_cast<int64_t>(B) * H * L * dh;
  p.dk = static_cast<bf16*>(ba->dk); p.dv = static_cast<bf16*>(ba->dv);
  p.scale = 1.f / sqrtf(static_cast<float>(dh));
  p.scale_log2 = 1.4426950408889634f * p.scale;
  p.noncausal = (a->flags & ME_ATTN_NONCAUSAL) ? 1 : 0;
  p.trace = g_attn_bwd_trace;
  dim3 grid((L + FB_BN - 1) / FB_BN, H, B);
  int rc;
  if (dh == 64) rc = launch_bwd<64>(tq, tk, tv, tdo, te, p, grid, st);
  else if (dh == 48) rc = launch_bwd<48>(tq, tk, tv, tdo, te, p, grid, st);
  else rc = launch_bwd<32>(tq, tk, tv, tdo, te, p, grid, st);
  if (rc) return rc;
  {
    const int64_t total = static_cast<int64_t>(B) * H * L * (dh / 4);
    const int64_t want = (total + 255) / 256;
    const int de_blocks = static_cast<int>((static_cast<int64_t>(a->max_seq) * (dh / 4) + 255) / 256);
    const int blocks = static_cast<int>(want < 148 * 16 ? want : 148 * 16) + de_blocks;
    attn_bwd_finish_kernel<<<blocks, 256, 0, st>>>(p.dq_ws, static_cast<bf16*>(ba->dq), a->q_sb, a->q_si, a->q_sh, B, H,
                                                   L, dh, p.dE_ws, ba->dE, a->max_seq, de_blocks);
    ME_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace me

extern "C" int me_debug_trace_set(long long* device_buf) {
  me::g_attn_bwd_trace = device_buf;
  return 0;
}

extern "C" int64_t me_attention_backward_workspace_floats(int B, int H, int L, int dh, int max_seq) {
  return (static_cast<int64_t>(B) * H * L + static_cast<int64_t>(me::FB_DE_COPIES) * max_seq) * dh;
}
