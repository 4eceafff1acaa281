// Relative global attention on the 5th-generation tensor cores -- backward, query side: dQ and dE.
//
// One CTA owns 128 query rows of one (batch, head) and walks the key tiles (64 keys) up to the diagonal, so dQ
// accumulates in tensor memory for the whole CTA and is written once -- no cross-CTA reduction.  The softmax
// backward itself (P, dS) is done once, by the key-owning kernel in attention_tc_bwd.cu, which leaves every
// 128 x 64 tile of dS (bf16, already in the UMMA K-major swizzled layout) in a scratch tensor; this kernel only
// streams those tiles back through TMA and multiplies.  (Splitting the backward in two removes the fp32 reduce-add
// traffic that bounded the fused round-1 kernel -- 80 KB per tile against a measured chip-wide reduce-add throughput
// of 2.8 TB/s, profiles/r02_a_micro_reduce_bw.txt -- and handing dS over instead of recomputing it keeps the
// per-logit CUDA-core work, the real limit of these kernels, where it was.)
//
// Band coordinates.  For this CTA (rows i0 .. i0+127) define g = 127 - a + j for query row a and key j:
//     Srel[a, j] = q_a . E[e_base + g],   e_base = max_seq - 128 - i0.
// The 64 keys of step t touch g in [64 t, 64 t + 190], i.e. the three 64-wide chunks t, t+1, t+2 -- consecutive steps
// share two of them.  dS in band coordinates, T[a, g] = dS[a, j], is assembled chunk by chunk in a ring of six
// shared-memory panels (the "unskew" of the reference); chunk t is complete after step t.  Per step:
//   warps 0-3  (thread = query row) dS row of the tile (shared memory, as delivered by TMA) -> registers -> T ring.
//              A row's 16-byte chunks of T start at keys j = (a + 1) mod 8: the thread prepends the (127 - a) mod 8
//              values it carried over from the previous tile, writes eight WHOLE chunks and carries the tail on --
//              no partial stores, nothing to re-zero (every chunk of a live panel is written exactly once)
//   MMA        dQ += dS K_t     dQ += T_t E_t  (K = 64)     [odd t]  dE[128 rows] = [T_{t-1} | T_t]^T Q
//   warps 4-7  (thread = row of the dE pair) TMEM -> registers -> fp32 reduce-adds into the CTA's private copy of dE,
//              concurrently with the assembly of the following steps (16 KB per step instead of the 48 KB of a
//              192-row band per step).
// Three load stages cover the latency of the dS tiles.
#include "attention_tc.cuh"

namespace me {

constexpr int QB_BM = 128;
constexpr int QB_BN = 64;
constexpr int QB_ROW_THREADS = 128;       // warps 0-3 assemble T (thread = query row), warps 4-7 drain dE (thread = dE row)
constexpr int QB_COMPUTE_THREADS = 256;
constexpr int QB_MMA_WARP = 8;
constexpr int QB_LOAD_WARP = 9;
constexpr int QB_THREADS = QB_COMPUTE_THREADS + 64;
constexpr int QB_STAGES = 3;                     // load stages: dS tile | K tile | chunk of E of a step
constexpr int QB_STAGE_DS = 0, QB_STAGE_K = 16384, QB_STAGE_E = 16384 + 8192, QB_STAGE_BYTES = 32768;
constexpr int QB_TSLOTS = 6;                     // T ring: chunk c -> slot c % 6.  Six panels let the threads assemble step
                                                 // t+1 while the MMAs of step t still read theirs
constexpr int QB_OFF_Q = 0;
constexpr int QB_OFF_STAGE = QB_OFF_Q + 16384;
constexpr int QB_OFF_T = QB_OFF_STAGE + QB_STAGES * QB_STAGE_BYTES;   // 128 x 64 bf16 panels
constexpr int QB_OFF_BAR = QB_OFF_T + QB_TSLOTS * 16384;
constexpr int QB_SMEM = QB_OFF_BAR + 256;
static_assert(QB_OFF_STAGE % 1024 == 0 && QB_OFF_T % 1024 == 0 && QB_OFF_BAR % 1024 == 0, "tile alignment");
static_assert(QB_SMEM <= 227 * 1024, "shared memory budget");
constexpr uint32_t QB_TMEM_COLS = 256;
// two dQ accumulators (the dS K and the T E products): back-to-back MMAs into ONE accumulator are a dependent chain
// and N = 64 instructions are short, so the issuer interleaves independent chains; the threads add the two at the end
constexpr uint32_t QB_COL_DQ = 0, QB_COL_DQ2 = 64, QB_COL_DE = 128;

long long* g_attn_trace = nullptr;   // me_debug_trace_set

struct QbParams {
  int B, H, L, max_seq;
  int64_t q_sb, q_sh, q_si;
  float* dE_ws;  // fp32 [FB_DE_COPIES, max_seq, dh]
  int b0;        // first sequence of this launch's slice of the batch (blockIdx.z counts from it)
  long long* trace;  // tuning builds (-DME_ATTN_TRACE): clock64() stamps of one CTA, [role][step < 20][event < 8]
  bf16* dq;
  int noncausal;
  int tiles_per_head;  // dS scratch: tile (qi, kt) of head (b, h) starts at row ((b H + h) tiles_per_head + index) 128
};

#ifdef ME_ATTN_TRACE
#define QB_TRACE(role, st, k)                                                                         \
  do {                                                                                                \
    if (tr && (st) < 20) p.trace[((role) * 20 + (st)) * 8 + (k)] = clock64();                         \
  } while (0)
#else
#define QB_TRACE(role, st, k) do { } while (0)
#endif

template <int DH>
__global__ void __launch_bounds__(QB_THREADS, 1)
attn_bwd_q_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmdS, const __grid_constant__ CUtensorMap tmE64,
                     QbParams p) {
  extern __shared__ __align__(1024) uint8_t qb_smem[];
  uint8_t* sQ = qb_smem + QB_OFF_Q;
  uint8_t* sStage = qb_smem + QB_OFF_STAGE;
  uint8_t* sT = qb_smem + QB_OFF_T;
  uint64_t* bars = reinterpret_cast<uint64_t*>(qb_smem + QB_OFF_BAR);
  // per-step barriers come in threes (index t % 3): a waiter can then never be a whole phase behind the barrier
  // it polls, however far the loads run ahead
  uint64_t* q_full = bars + 0;
  uint64_t* ld_full = bars + 1;    // [3]: dS tile, K tile and the chunk of E of step t
  uint64_t* a_done = bars + 4;     // [3] the T chunks of step t are in shared memory (128 arrivals)
  uint64_t* mma_done = bars + 7;   // [3] every MMA of step t has retired
  uint64_t* de_read = bars + 10;   // the dE pair is out of tensor memory (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qi = gridDim.x - 1 - blockIdx.x;  // heavy (late) query tiles first
  const int h = blockIdx.y, bl = blockIdx.z, b = p.b0 + bl;   // bl: index inside the slice (dS scratch), b: sequence
  const int i0 = qi * QB_BM;
  const int kmax = p.noncausal ? p.L : min(i0 + QB_BM, p.L);
  const int nt = (kmax + QB_BN - 1) / QB_BN;        // real steps (key tiles)
  // chunks of T that can hold a non-zero value AND meet a row of E below max_seq: c <= min(nt + 1, 2 qi + 1);
  // steps nt .. t_end-1 only drain them.  t_end is even so that the last pair is flushed.
  const int c_last = max(nt - 1, min(nt + 1, 2 * qi + 1));
  const int t_end = (c_last + 2) & ~1;
  const int e_base = p.max_seq - QB_BM - i0;
  const int nkt = (p.L + QB_BN - 1) / QB_BN;
  const int64_t ds_row0 =
      ((static_cast<int64_t>(bl) * p.H + h) * p.tiles_per_head + (p.noncausal ? qi * nkt : qi * (qi + 1))) * QB_BM;
  float* const dE_mine =
      p.dE_ws + static_cast<int64_t>((blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) % FB_DE_COPIES) *
                    p.max_seq * DH;

  const bool tr = p.trace != nullptr && qi == static_cast<int>(gridDim.x) - 1 && h == 0 && bl == 0 && lane == 0;
  (void)tr;
  if (tid == 0) {
    if ((smem_u32(qb_smem) & 1023u) != 0) __trap();
    mbar_init(q_full, 1);
    for (int k = 0; k < 3; ++k) {
      mbar_init(&ld_full[k], 1);
      mbar_init(&a_done[k], QB_ROW_THREADS);
      mbar_init(&mma_done[k], 1);
    }
    mbar_init(de_read, QB_ROW_THREADS);
    fence_mbar_init();
  }
  if (warp == QB_MMA_WARP) tmem_alloc(tmem_slot, QB_TMEM_COLS);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == QB_MMA_WARP) {
    // ======================= MMA issuer =======================
    constexpr uint32_t idesc_tt = make_idesc_bf16(128, DH, 1, 1);    // dE : A^T (MN-major) x B (MN-major)
    constexpr uint32_t idesc_nt = make_idesc_bf16(128, DH, 0, 1);    // dQ : A (K-major) x B (MN-major)
    const uint32_t q_addr = smem_u32(sQ);
    mbar_wait(q_full, 0);
    uint32_t dq_acc = 0, dq2_acc = 0;
    for (int t = 0; t < t_end; ++t) {
      const int s3 = t % 3;
      const uint32_t ph3 = (t / 3) & 1;
      QB_TRACE(1, t, 0);
      mbar_wait(&ld_full[s3], ph3);   // dS, K, E chunk of the step (dS was also read by the threads)
      QB_TRACE(1, t, 1);
      mbar_wait(&a_done[s3], ph3);
      tc_fence_after();
      QB_TRACE(1, t, 2);
      const uint32_t st_addr = smem_u32(sStage + s3 * QB_STAGE_BYTES);
      const uint32_t k_addr = st_addr + QB_STAGE_K, ds_addr = st_addr + QB_STAGE_DS, e_addr = st_addr + QB_STAGE_E;
      const uint32_t t_addr = smem_u32(sT + (t % QB_TSLOTS) * 16384);
      if ((t & 1) && t >= 3) {
        mbar_wait(de_read, ((t - 3) >> 1) & 1);
        tc_fence_after();
      }
      if (elect_one()) {
        const bool real = t < nt;
        const uint32_t pair_addr = smem_u32(sT + ((t - 1) % QB_TSLOTS) * 16384);
        // three independent accumulation chains, issued round-robin:
        //   dQ  += dS K_t            (real steps)
        //   dQ2 += T_t E_t           (chunk t is complete)
        //   dE   = [T_{t-1} | T_t]^T Q   (odd steps: the pair of chunks t-1, t; two instructions per round)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (real)
            umma_bf16(tmem_base + QB_COL_DQ, make_smem_desc_sw128(ds_addr + k * 32, 16, 1024),
                      make_smem_desc_sw128(k_addr + k * 2048, 8192, 1024), idesc_nt, (k > 0) ? 1u : dq_acc);
          umma_bf16(tmem_base + QB_COL_DQ2, make_smem_desc_sw128(t_addr + k * 32, 16, 1024),
                    make_smem_desc_sw128(e_addr + k * 2048, 8192, 1024), idesc_nt, (k > 0) ? 1u : dq2_acc);
#if defined(ME_EXP) && ME_EXP == 3
          if (false) {
#else
          if (t & 1) {
#endif
#pragma unroll
            for (int kk = 2 * k; kk < 2 * k + 2; ++kk)
              umma_bf16(tmem_base + QB_COL_DE, make_smem_desc_sw128(pair_addr + kk * 2048, 16384, 1024),
                        make_smem_desc_sw128(q_addr + kk * 2048, 8192, 1024), idesc_tt, kk > 0);
          }
        }
        umma_commit(&mma_done[s3]);
      }
      if (t < nt) dq_acc = 1;
      dq2_acc = 1;
      __syncwarp();
      QB_TRACE(1, t, 3);
    }
  } else if (warp == QB_LOAD_WARP) {
    // ======================= TMA loads =======================
    auto load_step = [&](int t) {
      const int s3 = t % 3;
      uint8_t* st = sStage + s3 * QB_STAGE_BYTES;
      const bool real = t < nt;
      mbar_arrive_expect_tx(&ld_full[s3], 8192 + (real ? 16384 + 8192 : 0));
      if (real) {
        tma_load_2d(&tmdS, &ld_full[s3], st + QB_STAGE_DS, 0, static_cast<int>(ds_row0 + static_cast<int64_t>(t) * QB_BM));
        tma_load_4d(&tmK, &ld_full[s3], st + QB_STAGE_K, 0, h, t * QB_BN, b);
      }
      tma_load_2d(&tmE64, &ld_full[s3], st + QB_STAGE_E, 0, e_base + 64 * t);
    };
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmdS);
      tma_prefetch_desc(&tmE64);
      mbar_arrive_expect_tx(q_full, 16384);
      tma_load_4d(&tmQ, q_full, sQ, 0, h, i0, b);
      for (int t = 0; t < QB_STAGES && t < t_end; ++t) load_step(t);
    }
    __syncwarp();
    for (int t = QB_STAGES; t < t_end; ++t) {
      mbar_wait(&mma_done[t % 3], ((t - 3) / 3) & 1);   // the MMAs of step t-3 have read stage t % 3
      QB_TRACE(2, t, 0);
      if (elect_one()) load_step(t);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ================================ dE drain (warps 4-7) ================================
    const int a = tid - QB_ROW_THREADS;          // row of the pair == TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp - 4) * 32) << 16);
    for (int f = 0; f < t_end / 2; ++f) {
      const int t = 2 * f + 1;                   // the step whose MMAs produced the pair
      mbar_wait(&mma_done[t % 3], (t / 3) & 1);
      tc_fence_after();
      uint32_t v[DH];
      tmem_ld_cols<DH>(t_lane + QB_COL_DE, v);
      tc_wait_ld();
      tc_fence_before();
      mbar_arrive(de_read);
      const int row = e_base + 128 * f + a;
#if defined(ME_EXP) && ME_EXP == 1
      if (false) {
#else
      if (row < p.max_seq) {
#endif
        float* dst = dE_mine + static_cast<int64_t>(row) * DH;
#pragma unroll
        for (int c = 0; c < DH; c += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c), "f"(__uint_as_float(v[c])),
                       "f"(__uint_as_float(v[c + 1])), "f"(__uint_as_float(v[c + 2])), "f"(__uint_as_float(v[c + 3]))
                       : "memory");
      }
    }
  } else {
    // ================================ T assembly (warps 0-3) ================================
    const int a = tid;                           // query row inside the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const int i = i0 + a;
    const bool row_ok = i < p.L;
    // the row's window of step t covers the band columns [64 t + base_g, 64 t + base_g + 64): `cl` values carried over
    // from the previous tile followed by the first 64 - cl values of this one
    const int cl = (127 - a) & 7;
    const int base_q = (127 - a) >> 3;           // base_g / 8: first 16-byte chunk of the window (step 0)
    auto chunk_ptr = [&](int q) -> uint8_t* {    // q: 16-byte chunk index along the band (8 columns each)
      return sT + ((q >> 3) % QB_TSLOTS) * 16384 + a * 128 + (((q & 7) ^ (a & 7)) << 4);
    };
    uint32_t carry[4] = {0u, 0u, 0u, 0u};        // the previous tile's last eight values of this row
    // band columns below the row's first window never get a value: zero them once (panels 0 and 1; every later
    // panel is written whole before it is read, up to the end of the row, which the last step pads with zeros)
    for (int q = 0; q < base_q; ++q) *reinterpret_cast<uint4*>(chunk_ptr(q)) = make_uint4(0, 0, 0, 0);

    for (int t = 0; t < t_end; ++t) {
      const bool real = t < nt;
      if (warp == 0) QB_TRACE(0, t, 0);
      if (real) {
        mbar_wait(&ld_full[t % 3], (t / 3) & 1);
        if (warp == 0) QB_TRACE(0, t, 2);
        uint32_t w[36];
#pragma unroll
        for (int c = 0; c < 4; ++c) w[c] = carry[c];
        const uint8_t* drow = sStage + (t % 3) * QB_STAGE_BYTES + QB_STAGE_DS + a * 128;
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          const uint4 u = *reinterpret_cast<const uint4*>(drow + ((n ^ (a & 7)) << 4));
          w[4 + 4 * n] = u.x; w[5 + 4 * n] = u.y; w[6 + 4 * n] = u.z; w[7 + 4 * n] = u.w;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) carry[c] = w[32 + c];
        // shift right by cl (0..7) elements; what moves in at the low end are the carried values
        const uint32_t on4 = cl & 4, on2 = cl & 2;
#pragma unroll
        for (int k = 35; k >= 2; --k) w[k] = sel_b32(w[k - 2], w[k], on4);
#pragma unroll
        for (int k = 35; k >= 3; --k) w[k] = sel_b32(w[k - 1], w[k], on2);
        const uint32_t hs = (cl & 1) ? 16u : 0u;
#pragma unroll
        for (int k = 35; k >= 4; --k) w[k] = __funnelshift_l(w[k - 1], w[k], hs);
        const int q0 = 8 * t + base_q;
#if !(defined(ME_EXP) && ME_EXP == 2)
#pragma unroll
        for (int n = 0; n < 8; ++n)
          *reinterpret_cast<uint4*>(chunk_ptr(q0 + n)) = make_uint4(w[4 + 4 * n], w[5 + 4 * n], w[6 + 4 * n], w[7 + 4 * n]);
#endif
        if (t == nt - 1) {
          // end of the row: the carried tail (cl values, then zeros) and zeros up to the end of the last panel the
          // MMAs will read -- those panels still hold chunks of six steps ago
          uint32_t z[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) z[c] = carry[c];
          // the tail chunk holds the last cl values in its first cl elements: shift the carry words right by 8 - cl
          // elements (i.e. keep elements 8-cl..7 and move them to 0..cl-1)
          {
            uint32_t x[8] = {z[0], z[1], z[2], z[3], 0u, 0u, 0u, 0u};
            const int sh = 8 - cl;                 // 1..8 elements to the left
            const uint32_t l4 = sh & 4, l2 = sh & 2, l8 = sh & 8;
#pragma unroll
            for (int k = 0; k < 4; ++k) x[k] = sel_b32(x[k + 2], x[k], l4);
#pragma unroll
            for (int k = 0; k < 4; ++k) x[k] = sel_b32(x[k + 1], x[k], l2);
            const uint32_t hl = (sh & 1) ? 16u : 0u;
#pragma unroll
            for (int k = 0; k < 4; ++k) x[k] = __funnelshift_r(x[k], x[k + 1], hl);
#pragma unroll
            for (int k = 0; k < 4; ++k) z[k] = l8 ? 0u : x[k];
          }
          *reinterpret_cast<uint4*>(chunk_ptr(q0 + 8)) = make_uint4(z[0], z[1], z[2], z[3]);
          for (int q = q0 + 9; q < 8 * t_end; ++q) *reinterpret_cast<uint4*>(chunk_ptr(q)) = make_uint4(0, 0, 0, 0);
        }
        fence_proxy_async_smem();
      }
      // (a real step got here through ld_full, which the loader arms only after the MMAs of step t-3: the same bound
      // keeps a drain step from arriving on a_done[t % 3] before the issuer has consumed its previous phase)
      if (!real && t >= 3) mbar_wait(&mma_done[t % 3], ((t - 3) / 3) & 1);
      if (warp == 0) QB_TRACE(0, t, 3);
      mbar_arrive(&a_done[t % 3]);
    }

    // dQ: complete in tensor memory, written once
    mbar_wait(&mma_done[(t_end - 1) % 3], ((t_end - 1) / 3) & 1);
    tc_fence_after();
    {
      bf16* dqrow = p.dq + static_cast<int64_t>(b) * p.q_sb + static_cast<int64_t>(i) * p.q_si + h * p.q_sh;
#pragma unroll
      for (int c0 = 0; c0 < DH; c0 += 8) {
        uint32_t vq[8], vq2[8];
        tmem_ld8(t_lane + QB_COL_DQ + c0, vq);   // (warp-collective: outside the row predicate)
        tmem_ld8(t_lane + QB_COL_DQ2 + c0, vq2);
        tc_wait_ld();
        uint4 u;
        __nv_bfloat162* hq = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          hq[e] = __floats2bfloat162_rn(__uint_as_float(vq[2 * e]) + __uint_as_float(vq2[2 * e]),
                                        __uint_as_float(vq[2 * e + 1]) + __uint_as_float(vq2[2 * e + 1]));
        if (row_ok) *reinterpret_cast<uint4*>(dqrow + c0) = u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == QB_MMA_WARP) tmem_dealloc(tmem_base, QB_TMEM_COLS);
}

template <int DH>
static int launch_bwd_q(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tds, const CUtensorMap& te64,
                        const QbParams& p, dim3 grid, cudaStream_t st) {
  auto kern = attn_bwd_q_tc_kernel<DH>;
  static bool configured = false;
  if (!configured) {
    ME_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, QB_SMEM));
    configured = true;
  }
  cudaEvent_t pe = prof_begin(3.0 * attn_unit_flops(static_cast<int>(grid.z), p.H, p.L, DH), st, 3);   // dS K, T E, T^T Q
  kern<<<grid, QB_THREADS, QB_SMEM, st>>>(tq, tk, tds, te64, p);
  prof_end(pe, st);
  ME_LAUNCH_CHECK();
  return 0;
}

// dq (bf16, q strides) and the dE partial sums (into the private copies of dE_ws) for one attention call, from the
// dS tiles the key-side kernel left in `ds_scratch` ([tiles, 128, 64] bf16, described by `tds`)
int launch_attn_bwd_q_tc(const me_attn_bwd_args* ba, float* dE_ws, const CUtensorMap& tds, int tiles_per_head, int b0,
                         int nb) {
  const me_attn_args* a = &ba->f;
  const int B = a->B, H = a->H, L = a->Lq, dh = a->dh;
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  CUtensorMap tq, tk, te64;
  if (qkv_map(&tq, a->q, dh, H, L, B, a->q_sh, a->q_si, a->q_sb, QB_BM)) return 1;
  if (qkv_map(&tk, a->k, dh, H, L, B, a->k_sh, a->k_sj, a->k_sb, QB_BN)) return 1;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(dh), static_cast<uint64_t>(a->max_seq)};
    const uint64_t strides[1] = {static_cast<uint64_t>(dh)};
    const uint32_t box64[2] = {64, 64};
    if (make_tmap_nd_bf16(&te64, a->E, 2, dims, strides, box64)) return 1;
  }
  QbParams p;
  p.B = B; p.H = H; p.L = L; p.max_seq = a->max_seq;
  p.q_sb = a->q_sb; p.q_sh = a->q_sh; p.q_si = a->q_si;
  p.dE_ws = dE_ws;
  p.dq = static_cast<bf16*>(ba->dq);
  p.noncausal = (a->flags & ME_ATTN_NONCAUSAL) ? 1 : 0;
  p.tiles_per_head = tiles_per_head;
  p.b0 = b0;
  p.trace = g_attn_trace ? g_attn_trace + 3 * 20 * 8 : nullptr;
  dim3 grid((L + QB_BM - 1) / QB_BM, H, nb);
  if (dh == 64) return launch_bwd_q<64>(tq, tk, tds, te64, p, grid, st);
  if (dh == 48) return launch_bwd_q<48>(tq, tk, tds, te64, p, grid, st);
  return launch_bwd_q<32>(tq, tk, tds, te64, p, grid, st);
}

}  // namespace me
