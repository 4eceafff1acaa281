// Relative global attention on the 5th-generation tensor cores -- forward, persistent pipelined version.
//
//   S[i,j] = (q_i.k_j + q_i.E[max_seq-1-(i-j)]) / sqrt(dh),  j <= i, key j not pad;   O = softmax(S) V
//
// Same arithmetic as attention_tc.cu (which stays as the two-CTAs-per-SM variant, ME_ATTN_FWD=1), different
// schedule.  One persistent CTA per SM claims units (128 query rows of one (batch, head)) from a global counter,
// heavy units first inside groups of heads that keep K/V in L2; its twelve warps have five roles:
//
//   warps 0-3 / 4-7  two softmax teams.  The key tiles of the CTA form one stream g = 0, 1, 2, ... across units;
//                    team k takes the tiles with g & 1 == k and runs an independent online softmax over its
//                    tiles of a unit (thread = query row = TMEM lane); the two partial results (m, l, O) are
//                    merged once per unit.  While one team is in its exponentials the other one's S tile is
//                    being computed, so neither waits for a tensor-core round trip.
//   warp 8           issues the tcgen05.mma of S_k = Q K_t^T (N=64) and of one 64-row chunk of the relative band
//                    G = Q Eband^T per tile, running ahead into the next unit;
//   warp 11          issues O_k += P_k V_t (a warp of its own, so that neither issuer sits behind the other's wait).
//   warp 9           TMA producer (Q and the first two band chunks per unit, double buffered; K_t and the next
//                    chunk of E in one ring, V_t in another).
//   warp 10          stores the P tiles for the backward pass (TMA) when they are saved.
//
//   * The relative band rolls: band coordinate g = 127 - a + j is the same for every key tile of the unit
//     (E row = max_seq - 128 - i0 + g), tile t needs the chunks t, t+1, t+2 of 64 columns, so each tile adds
//     ONE new chunk to a four-slot ring in TMEM (the first tile three) instead of recomputing 192 columns.
//   * O lives in TMEM and is accumulated by the tensor core across tiles (no per-tile read-back).  The
//     running maximum is lazy: P = exp2(x c - m_run) with m_run raised only when the tile maximum exceeds it
//     by more than 2^8 (then the O accumulator is rescaled in TMEM), so P <= 256 and the rescale is rare.
//   * TMEM: S_A 0, S_B 64, O_A 128, O_B 192, ring 256..511.
#include "attention_tc.cuh"

namespace me {

constexpr int F2_BM = 128, F2_BN = 64;
constexpr int F2_THREADS = 384;
constexpr int F2_NKE = 4;                      // K_t | E chunk t+2 stages (free as soon as S_t is computed)
constexpr int F2_NV = 4;                       // V_t stages (free when P V_t has retired)
constexpr int F2_T = F2_BN * 128;              // one 64-row bf16 tile
constexpr int F2_Q_BYTES = F2_BM * 128;
constexpr int F2_KE = 2 * F2_T;
constexpr int F2_P_BYTES = F2_BM * 128;
constexpr int F2_MERGE_LD = 67;                // floats per row of the merge scratch: O[dh], m, l (odd: conflict-free)
constexpr int F2_MERGE_BYTES = F2_BM * F2_MERGE_LD * 4;
constexpr int F2_SMEM = 2 * F2_Q_BYTES + 4 * F2_T + F2_NKE * F2_KE + F2_NV * F2_T + 2 * F2_P_BYTES + F2_MERGE_BYTES + 512;
constexpr uint32_t F2_COL_S = 0, F2_COL_O = 128, F2_COL_G = 256;
#if defined(ME_EXP) && ME_EXP >= 100
constexpr int F2_PV_LEAD = ME_EXP - 100;
#else
constexpr int F2_PV_LEAD = 2;                  // S(g) is issued after P.V(g - 3)
#endif
constexpr float F2_TAU = 8.f;                  // log2 of the largest P value the lazy maximum allows
constexpr int F2_GROUP_HEADS = 74;             // (batch, head) pairs per scheduling group: their K/V (19 MB) stay in L2

struct F2Params {
  int B, H, L, max_seq;
  int64_t o_sb, o_si, keypad_ld;
  const uint8_t* keypad;
  bf16* out;
  float* lse;
  float scale_log2, sqrt_dh;
  float* m_tiles;
  int tiles_per_head;
  int noncausal;
  int nq, total_units;
  int* counters;         // [0] units claimed, [1] CTAs finished (the last one re-arms both)
  long long* trace;      // tuning builds (-DME_ATTN_TRACE)
};

extern long long* g_attn_trace;
#ifdef ME_ATTN_TRACE
#define F2_TRACE(role, st, k)                                                                        \
  do {                                                                                                \
    if (tr && (st) < 20) p.trace[((role) * 20 + (st)) * 8 + (k)] = clock64();                         \
  } while (0)
#else
#define F2_TRACE(role, st, k) do { } while (0)
#endif

__device__ __forceinline__ float f2_bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// TMEM column of band coordinate g (ring of four 64-column chunks)
__device__ __forceinline__ uint32_t f2_ring_col(int g) {
  return F2_COL_G + (static_cast<uint32_t>(g >> 6) & 3u) * 64u + static_cast<uint32_t>(g & 63);
}

template <int DH, bool RR>
__global__ void __launch_bounds__(F2_THREADS, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmE,
                 const __grid_constant__ CUtensorMap tmP, F2Params p) {
  extern __shared__ __align__(1024) uint8_t f2_smem[];
  uint8_t* sQ = f2_smem;                        // [2] per unit parity
  uint8_t* sE01 = sQ + 2 * F2_Q_BYTES;          // [2] chunks 0 and 1 of the band (first tile of a unit)
  uint8_t* sKE = sE01 + 4 * F2_T;               // per stage: K_t | E chunk t+2
  uint8_t* sV = sKE + F2_NKE * F2_KE;
  uint8_t* sP = sV + F2_NV * F2_T;              // one P tile per team
  float* sMerge = reinterpret_cast<float*>(sP + 2 * F2_P_BYTES);
  int4* slots = reinterpret_cast<int4*>(reinterpret_cast<uint8_t*>(sMerge) + F2_MERGE_BYTES);   // [2] {qi, h, b, nt}
  uint64_t* bars = reinterpret_cast<uint64_t*>(slots + 2);
  uint64_t* q_full = bars + 0;     // [2]   unit descriptor, Q and band chunks 0/1 of the unit are in shared memory
  uint64_t* q_free = bars + 2;     // [2]   every S/G product of the unit retired, every reader has the descriptor
  uint64_t* ke_full = bars + 4;    // [NKE]
  uint64_t* ke_free = bars + 8;    // [NKE] S and ring chunk of the tile computed
  uint64_t* v_full = bars + 12;    // [NV]
  uint64_t* v_free = bars + 16;    // [NV]  P.V of the tile retired
  uint64_t* s_full = bars + 20;    // [2]   S_k and the ring chunk of the tile are in TMEM
  uint64_t* sr_read = bars + 22;   // [2]   team k has its S row and band window in registers
  uint64_t* p_ready = bars + 24;   // [2]   P_k written (and O_k rescaled when the maximum moved)
  uint64_t* pv_done = bars + 26;   // [2]   O_k += P_k V retired
  uint64_t* p_free = bars + 28;    // [2]   saved-tile store has read P_k
  uint64_t* scr_full = bars + 30;  //       the early team's partial (O, m, l) of the unit is in the merge scratch
  uint64_t* scr_free = bars + 31;  //       ... and has been read by the merging team
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);
  volatile int* pv_issued = reinterpret_cast<volatile int*>(bars + 33);   // P.V products issued so far

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool save = p.m_tiles != nullptr;
  // output rows 32-byte aligned (the C-ABI only asks for 16): 256-bit stores
  const bool f2_rows32 = ((reinterpret_cast<uintptr_t>(p.out) | (p.o_si * 2) | (p.o_sb * 2) | (DH * 2)) & 31) == 0;
  const int nkt_all = (p.L + F2_BN - 1) / F2_BN;
  const bool tr = p.trace != nullptr && blockIdx.x == 0 && lane == 0;
  (void)tr;

  if (tid == 0) {
    if ((smem_u32(f2_smem) & 1023u) != 0) __trap();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmE);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_free[s], save ? 11 : 10);  // S-side MMA commit, P.V warp, one lane of each team warp (+ the store warp)
      mbar_init(&s_full[s], 1);
      mbar_init(&sr_read[s], 128);
      mbar_init(&p_ready[s], 128);
      mbar_init(&pv_done[s], 1);
      mbar_init(&p_free[s], 1);
    }
    *pv_issued = 0;
    mbar_init(scr_full, 128);
    mbar_init(scr_free, 128);
    for (int s = 0; s < F2_NKE; ++s) {
      mbar_init(&ke_full[s], 1);
      mbar_init(&ke_free[s], 1);
    }
    for (int s = 0; s < F2_NV; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_free[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 8) {
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 9) {
    // ------------------------------------------------------------------ unit claims + TMA producer
    if (elect_one()) {
      int g0 = 0;
      for (int ui = 0;; ++ui) {
        const int buf = ui & 1;
        const int u = atomicAdd(&p.counters[0], 1);
        if (ui >= 2) mbar_wait(&q_free[buf], ((ui >> 1) - 1) & 1);
        if (u >= p.total_units) {
          slots[buf] = make_int4(0, 0, 0, 0);   // nt = 0: no more work
          mbar_arrive(&q_full[buf]);
          break;
        }
        // heavy query tiles first inside a group of heads
        const int per_group = F2_GROUP_HEADS * p.nq;
        const int grp = u / per_group, r = u - grp * per_group;
        const int heads = min(F2_GROUP_HEADS, p.B * p.H - grp * F2_GROUP_HEADS);
        const int qi = p.nq - 1 - r / heads;
        const int bh = grp * F2_GROUP_HEADS + r % heads;
        const int b = bh / p.H, h = bh - b * p.H;
        const int i0 = qi * F2_BM;
        const int kmax = p.noncausal ? p.L : min(i0 + F2_BM, p.L);
        const int nt = (kmax + F2_BN - 1) / F2_BN;
        const int e_base = p.max_seq - F2_BM - i0;   // E row of band coordinate 0
        slots[buf] = make_int4(qi, h, b, nt);
        mbar_arrive_expect_tx(&q_full[buf], F2_Q_BYTES + 2 * F2_T);
        tma_load_4d(&tmQ, &q_full[buf], sQ + buf * F2_Q_BYTES, 0, h, i0, b);
        tma_load_2d(&tmE, &q_full[buf], sE01 + buf * 2 * F2_T, 0, e_base);
        tma_load_2d(&tmE, &q_full[buf], sE01 + buf * 2 * F2_T + F2_T, 0, e_base + 64);
        auto load_ke = [&](int t) {
          const int g = g0 + t, s = g % F2_NKE;
          if (g >= F2_NKE) mbar_wait(&ke_free[s], ((g / F2_NKE) - 1) & 1);
          uint8_t* st = sKE + s * F2_KE;
          mbar_arrive_expect_tx(&ke_full[s], F2_KE);
          tma_load_4d(&tmK, &ke_full[s], st, 0, h, t * F2_BN, b);
          tma_load_2d(&tmE, &ke_full[s], st + F2_T, 0, e_base + 64 * (t + 2));
        };
        load_ke(0);
        if (nt > 1) load_ke(1);
        for (int t = 0; t < nt; ++t) {   // (in the order the MMA warp releases the stages)
          if (t + 2 < nt) load_ke(t + 2);
          const int g = g0 + t, s = g % F2_NV;
          if (g >= F2_NV) mbar_wait(&v_free[s], ((g / F2_NV) - 1) & 1);
          mbar_arrive_expect_tx(&v_full[s], F2_T);
          tma_load_4d(&tmV, &v_full[s], sV + s * F2_T, 0, h, t * F2_BN, b);
        }
        g0 += nt;
      }
    }
    __syncwarp();
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer: S and band chunks
    constexpr uint32_t idesc_s = make_idesc_bf16(F2_BM, 64, 0, 0);
    constexpr uint32_t idesc_g2 = make_idesc_bf16(F2_BM, 128, 0, 0);
    int g = 0;
    int sr_waited[2] = {0, 0};     // phases of sr_read[k] already consumed
    auto wait_read = [&](int gg) { // tile gg has been read out of TMEM by its team
      if (gg < 0) return;
      const int k = gg & 1, n = gg >> 1;
      while (sr_waited[k] <= n) {
        mbar_wait(&sr_read[k], sr_waited[k] & 1);
        ++sr_waited[k];
      }
    };
    for (int ui = 0;; ++ui) {
      const int buf = ui & 1;
      mbar_wait(&q_full[buf], (ui >> 1) & 1);
      const int nt = slots[buf].w;
      if (nt == 0) break;
      const uint32_t q_addr = smem_u32(sQ + buf * F2_Q_BYTES);
      for (int t = 0; t < nt; ++t, ++g) {
        // the S buffer and the ring slot are free once tile g-2 has been read; a new unit rewrites the whole
        // ring, so its first tile also waits for tile g-1
        wait_read(g - 2);
        if (t == 0) wait_read(g - 1);
        const int s = g % F2_NKE;
        mbar_wait(&ke_full[s], (g / F2_NKE) & 1);
        // The tensor pipe is a FIFO: S(g) is not needed before its team has finished tile g-2, P.V(g-3) is needed
        // now -- let it go first.
        while (*pv_issued < g - F2_PV_LEAD) { }
        tc_fence_after();
        if (elect_one()) {
          const uint32_t st = smem_u32(sKE + s * F2_KE);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)
            umma_bf16(tmem_base + F2_COL_S + 64u * (g & 1), make_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                      make_smem_desc_sw128(st + k * 32, 16, 1024), idesc_s, k > 0);
          if (t == 0) {
            const uint32_t e01 = smem_u32(sE01 + buf * 2 * F2_T);
#pragma unroll
            for (int k = 0; k < DH / 16; ++k)
              umma_bf16(tmem_base + F2_COL_G, make_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                        make_smem_desc_sw128(e01 + k * 32, 16, 1024), idesc_g2, k > 0);
          }
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)
            umma_bf16(tmem_base + F2_COL_G + 64u * ((t + 2) & 3), make_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                      make_smem_desc_sw128(st + F2_T + k * 32, 16, 1024), idesc_s, k > 0);
          umma_commit(&s_full[g & 1]);
          umma_commit(&ke_free[s]);
          if (t == nt - 1) umma_commit(&q_free[buf]);
        }
        __syncwarp();
        F2_TRACE(2, g, 0);
      }
    }
  } else if (warp == 11) {
    // ------------------------------------------------------------------ MMA issuer: O_k (+)= P_k V_t
    // (a second issuing warp: each of the two blocks on its own barriers, neither sits behind the other's wait)
    constexpr uint32_t idesc_o = make_idesc_bf16(F2_BM, DH, 0, 1);
    int g = 0;
    for (int ui = 0;; ++ui) {
      const int buf = ui & 1;
      mbar_wait(&q_full[buf], (ui >> 1) & 1);
      const int nt = slots[buf].w;
      __syncwarp();
      if (lane == 0) mbar_arrive(&q_free[buf]);
      if (nt == 0) break;
      for (int t = 0; t < nt; ++t, ++g) {
        const int k = g & 1, n = g >> 1, s = g % F2_NV;
        mbar_wait(&v_full[s], (g / F2_NV) & 1);
        mbar_wait(&p_ready[k], n & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t pa = smem_u32(sP + k * F2_P_BYTES);
          const uint32_t va = smem_u32(sV + s * F2_T);
#pragma unroll
          for (int kk = 0; kk < F2_BN / 16; ++kk)   // the first tile of a team in a unit starts its accumulator
            umma_bf16(tmem_base + F2_COL_O + 64u * k, make_smem_desc_sw128(pa + kk * 32, 16, 1024),
                      make_smem_desc_sw128(va + kk * 2048, 8192, 1024), idesc_o, (t >= 2 || kk > 0) ? 1u : 0u);
          umma_commit(&pv_done[k]);
          umma_commit(&v_free[s]);
        }
        __syncwarp();
        if (lane == 0) *pv_issued = g + 1;
        F2_TRACE(2, g, 1);
      }
    }
  } else if (warp == 10) {
    // ------------------------------------------------------------------ saved-tile stores
    // (a warp of its own: the bulk-group wait that releases the P buffer must not sit in the MMA issue order)
    if (save) {
      int g0 = 0;
      for (int ui = 0;; ++ui) {
        const int buf = ui & 1;
        mbar_wait(&q_full[buf], (ui >> 1) & 1);
        const int4 un = slots[buf];
        __syncwarp();
        if (lane == 0) mbar_arrive(&q_free[buf]);
        if (un.w == 0) break;
        const int64_t tile0 = (static_cast<int64_t>(un.z) * p.H + un.y) * p.tiles_per_head +
                              (p.noncausal ? static_cast<int64_t>(un.x) * nkt_all : static_cast<int64_t>(un.x) * (un.x + 1));
        for (int t = 0; t < un.w; ++t) {
          const int g = g0 + t, k = g & 1, n = g >> 1;
          mbar_wait(&p_ready[k], n & 1);
          if (elect_one()) {
            tma_store_2d(&tmP, sP + k * F2_P_BYTES, 0, static_cast<int>((tile0 + t) * F2_BM));
            bulk_commit();
            bulk_wait_read_all();
            mbar_arrive(&p_free[k]);
          }
          __syncwarp();
        }
        g0 += un.w;
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax teams
    const int k = warp >> 2, w = warp & 3;
    const int a = tid & 127;              // query row inside the tile == TMEM lane
    const int shift = 31 - lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(w * 32) << 16);
    const float cs = RR ? 1.4426950408889634f : p.scale_log2;
    uint8_t* prow = sP + k * F2_P_BYTES + a * 128;
    float* mrow = sMerge + a * F2_MERGE_LD;
    int g0 = 0;
    for (int ui = 0;; ++ui) {
      const int buf = ui & 1;
      mbar_wait(&q_full[buf], (ui >> 1) & 1);
      const int4 un = slots[buf];
      __syncwarp();
      if (lane == 0) mbar_arrive(&q_free[buf]);
      const int nt = un.w;
      if (nt == 0) break;
      const int qi = un.x, h = un.y, b = un.z;
      const int i0 = qi * F2_BM, i = i0 + a;
      const uint8_t* kp = p.keypad ? p.keypad + static_cast<int64_t>(b) * p.keypad_ld : nullptr;
      const bool any_kp = kp != nullptr || p.noncausal;
      const int64_t tile0 = (static_cast<int64_t>(b) * p.H + h) * p.tiles_per_head +
                            (p.noncausal ? static_cast<int64_t>(qi) * nkt_all : static_cast<int64_t>(qi) * (qi + 1));
      float m_run = -INFINITY, l = 0.f;
      int own = 0;                 // tiles of this unit taken by this team so far
      int n_last = -1;
      for (int t = (k ^ (g0 & 1)); t < nt; t += 2, ++own) {
        const int n = (g0 + t) >> 1;   // running tile count of this team: phase of its barriers
        n_last = n;
        const int j0 = t * F2_BN;
        uint32_t kp0 = 0, kp1 = 0;
        if (any_kp) {
          const int ja = j0 + lane, jb = j0 + 32 + lane;
          kp0 = __ballot_sync(0xffffffffu, ja >= p.L || (kp && kp[ja] != 0));
          kp1 = __ballot_sync(0xffffffffu, jb >= p.L || (kp && kp[jb] != 0));
        }
        const int lim = p.noncausal ? 63 : i - j0;
        uint32_t v0 = lim >= 31 ? 0xffffffffu : (lim < 0 ? 0u : ((2u << lim) - 1u));
        uint32_t v1 = lim >= 63 ? 0xffffffffu : (lim < 32 ? 0u : ((2u << (lim - 32)) - 1u));
        v0 &= ~kp0;
        v1 &= ~kp1;
        const bool need_mask = !__all_sync(0xffffffffu, (v0 & v1) == 0xffffffffu);

        mbar_wait(&s_full[k], n & 1);
        tc_fence_after();
        if (w == 0) F2_TRACE(k, n, 0);

        float x0[32], x1[32];
        const int gb = 64 * t + 96 - 32 * w;   // band coordinate of the window of this warp
        {
          uint32_t sv[32], rv[64];
          tmem_ld32(t_lane + F2_COL_S + 64u * k, sv);
          tmem_ld32(t_lane + f2_ring_col(gb), reinterpret_cast<uint32_t(&)[32]>(rv[0]));
          tmem_ld32(t_lane + f2_ring_col(gb + 32), reinterpret_cast<uint32_t(&)[32]>(rv[32]));
          tc_wait_ld();
          skew_select(rv, shift);
#pragma unroll
          for (int bb = 0; bb < 32; ++bb) {
            if (RR)
              x0[bb] = f2_bf16r(__fdiv_rn(f2_bf16r(f2_bf16r(__uint_as_float(sv[bb])) + f2_bf16r(__uint_as_float(rv[bb]))), p.sqrt_dh));
            else
              x0[bb] = __uint_as_float(sv[bb]) + __uint_as_float(rv[bb]);
          }
        }
        {
          uint32_t sv[32], rv[64];
          tmem_ld32(t_lane + F2_COL_S + 64u * k + 32, sv);
          tmem_ld32(t_lane + f2_ring_col(gb + 32), reinterpret_cast<uint32_t(&)[32]>(rv[0]));
          tmem_ld32(t_lane + f2_ring_col(gb + 64), reinterpret_cast<uint32_t(&)[32]>(rv[32]));
          tc_wait_ld();
          tc_fence_before();
          mbar_arrive(&sr_read[k]);          // S_k and the band window of this tile are in registers
          if (w == 0) F2_TRACE(k, n, 1);
          skew_select(rv, shift);
#pragma unroll
          for (int bb = 0; bb < 32; ++bb) {
            if (RR)
              x1[bb] = f2_bf16r(__fdiv_rn(f2_bf16r(f2_bf16r(__uint_as_float(sv[bb])) + f2_bf16r(__uint_as_float(rv[bb]))), p.sqrt_dh));
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
        const float m_tile = mx * cs;
        if (w == 0) F2_TRACE(k, n, 2);

        // P_k (and O_k) of the previous own tile are done with
        if (own > 0) mbar_wait(&pv_done[k], (n - 1) & 1);
        if (save && n > 0) mbar_wait(&p_free[k], (n - 1) & 1);
        tc_fence_after();
        if (w == 0) F2_TRACE(k, n, 3);
        const bool grow = m_run != -INFINITY && m_tile > m_run + F2_TAU;
        if (m_run == -INFINITY) m_run = m_tile;   // nothing accumulated yet for this row (O row and l are 0)
        if (__any_sync(0xffffffffu, grow)) {      // rare: rescale the accumulator of the rows whose maximum moved
          const float alpha = grow ? fast_exp2(m_run - m_tile) : 1.f;
          if (grow) {
            m_run = m_tile;
            l *= alpha;
          }
#pragma unroll
          for (int c0 = 0; c0 < DH; c0 += 16) {
            uint32_t ov[16];
            tmem_ld16(t_lane + F2_COL_O + 64u * k + c0, ov);
            tc_wait_ld();
#pragma unroll
            for (int c = 0; c < 16; ++c) ov[c] = __float_as_uint(__uint_as_float(ov[c]) * alpha);
            tmem_st16(t_lane + F2_COL_O + 64u * k + c0, ov);
          }
          tc_wait_st();
        }
        const float m_use = (m_run == -INFINITY) ? 0.f : m_run;
        float rs = 0.f;
#pragma unroll
        for (int kc = 0; kc < 8; ++kc) {
          uint32_t wv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int bb = (kc & 3) * 8 + 2 * e;
            const float xa = (kc < 4) ? x0[bb] : x1[bb];
            const float xb = (kc < 4) ? x0[bb + 1] : x1[bb + 1];
            const float pa = fast_exp2(fmaf(xa, cs, -m_use));
            const float pb = fast_exp2(fmaf(xb, cs, -m_use));
            rs += pa + pb;
            __nv_bfloat162 h2 = __floats2bfloat162_rn(pa, pb);
            wv[e] = *reinterpret_cast<uint32_t*>(&h2);
          }
          *reinterpret_cast<uint4*>(prow + ((kc ^ (a & 7)) << 4)) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        }
        l += rs;
        if (save) p.m_tiles[(tile0 + t) * F2_BM + a] = m_use;
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&p_ready[k]);
        if (w == 0) F2_TRACE(k, n, 4);
      }

      // ---- merge the two teams: O = (O_A 2^(m_A - m) + O_B 2^(m_B - m)) / (l_A 2^(m_A - m) + l_B 2^(m_B - m))
      float O[DH];
      if (own > 0) {
        mbar_wait(&pv_done[k], n_last & 1);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < DH; c0 += 16) {
          uint32_t ov[16];
          tmem_ld16(t_lane + F2_COL_O + 64u * k + c0, ov);
          tc_wait_ld();
#pragma unroll
          for (int c = 0; c < 16; ++c) O[c0 + c] = __uint_as_float(ov[c]);
        }
        tc_fence_before();   // (the next P.V into O_k is ordered behind this team's next p_ready arrival)
      } else {
#pragma unroll
        for (int c = 0; c < DH; ++c) O[c] = 0.f;
      }
      // The team that owns the last tile of the unit merges; the other one leaves its partial result in the
      // scratch and goes on to the next unit without waiting.
      const int late = (g0 + nt - 1) & 1;
      if (w == 0) F2_TRACE(k, n_last, 5);
      if (k != late) {
        if (ui > 0) mbar_wait(scr_free, (ui - 1) & 1);
#pragma unroll
        for (int c = 0; c < DH; ++c) mrow[c] = O[c];
        mrow[DH] = m_run;
        mrow[DH + 1] = l;
        mbar_arrive(scr_full);
      } else {
        mbar_wait(scr_full, ui & 1);
        if (w == 0) F2_TRACE(k, n_last, 3);
        const float mB = mrow[DH], lB = mrow[DH + 1];
        const float m = fmaxf(m_run, mB);
        const float wA = (m == -INFINITY) ? 0.f : fast_exp2(m_run - m);
        const float wB = (m == -INFINITY) ? 0.f : fast_exp2(mB - m);
        const float lt = l * wA + lB * wB;
        const float inv = lt > 0.f ? 1.f / lt : 0.f;   // fully masked row -> 0 (reference: NaN)
#pragma unroll
        for (int c = 0; c < DH; ++c) O[c] = (O[c] * wA + mrow[c] * wB) * inv;
        mbar_arrive(scr_free);
        if (w == 0) F2_TRACE(k, n_last, 7);
        if (i < p.L) {
          bf16* orow = p.out + static_cast<int64_t>(b) * p.o_sb + static_cast<int64_t>(i) * p.o_si + h * DH;
#pragma unroll
          for (int c0 = 0; c0 < DH; c0 += 16) {   // whole 32-byte sectors per store
            uint32_t u[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(O[c0 + 2 * e], O[c0 + 2 * e + 1]);
              u[e] = *reinterpret_cast<uint32_t*>(&h2);
            }
            if (f2_rows32) {
              asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(orow + c0), "r"(u[0]),
                           "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7])
                           : "memory");
            } else {
              *reinterpret_cast<uint4*>(orow + c0) = make_uint4(u[0], u[1], u[2], u[3]);
              *reinterpret_cast<uint4*>(orow + c0 + 8) = make_uint4(u[4], u[5], u[6], u[7]);
            }
          }
          if (p.lse)
            p.lse[(static_cast<int64_t>(b) * p.H + h) * p.L + i] =
                lt > 0.f ? (m + log2f(lt)) * 0.69314718055994530942f : -INFINITY;
        }
      }
      if (w == 0) F2_TRACE(k, n_last, 6);
      g0 += nt;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
  if (tid == 0) {   // the last CTA to finish re-arms the counters for the next launch
    __threadfence();
    if (atomicAdd(&p.counters[1], 1) == static_cast<int>(gridDim.x) - 1) {
      p.counters[0] = 0;
      p.counters[1] = 0;
      __threadfence();
    }
  }
}

template <int DH, bool RR>
static int launch_fwd2(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& te,
                       const CUtensorMap& tp, const F2Params& p, int grid, cudaStream_t st) {
  auto kern = attn_fwd2_kernel<DH, RR>;
  static bool configured = false;
  if (!configured) {
    ME_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM));
    configured = true;
  }
  cudaEvent_t pe = prof_begin(3.0 * attn_unit_flops(p.B, p.H, p.L, DH), st, 1);   // QK^T, QE^T, PV
  kern<<<grid, F2_THREADS, F2_SMEM, st>>>(tq, tk, tv, te, tp, p);
  prof_end(pe, st);
  ME_LAUNCH_CHECK();
  return 0;
}

// Unit-claim counters of the persistent kernel: one pair per device, zeroed once (the kernel re-arms them at its
// end).  Forward launches of one process are stream-ordered on one device, which is what the single pair assumes.
static int* fwd2_counters(cudaStream_t st, bool* capturing) {
  static int* ptr[64] = {};
  int dev = 0;
  *capturing = false;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    set_error("me_attention_forward: cudaGetDevice failed");
    return nullptr;
  }
  if (ptr[dev] == nullptr) {
    // (no allocation inside a stream capture: the caller falls back to the two-CTAs-per-SM kernel for that call)
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) {
      *capturing = true;
      return nullptr;
    }
    int* q = nullptr;
    if (cudaMalloc(&q, 2 * sizeof(int)) != cudaSuccess || cudaMemset(q, 0, 2 * sizeof(int)) != cudaSuccess) {
      set_error("me_attention_forward: cannot allocate the unit counters");
      return nullptr;
    }
    ptr[dev] = q;
  }
  return ptr[dev];
}

// (argument checks are done by the caller, launch_attn_fwd_tc; returns -1 when the caller should use the other kernel)
int launch_attn_fwd2_tc(const me_attn_args* a) {
  CUtensorMap tq, tk, tv, te;
  if (qkv_map(&tq, a->q, a->dh, a->H, a->Lq, a->B, a->q_sh, a->q_si, a->q_sb, F2_BM)) return 1;
  if (qkv_map(&tk, a->k, a->dh, a->H, a->Lk, a->B, a->k_sh, a->k_sj, a->k_sb, F2_BN)) return 1;
  if (qkv_map(&tv, a->v, a->dh, a->H, a->Lk, a->B, a->v_sh, a->v_sj, a->v_sb, F2_BN)) return 1;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(a->dh), static_cast<uint64_t>(a->max_seq)};
    const uint64_t strides[1] = {static_cast<uint64_t>(a->dh)};
    const uint32_t box[2] = {64, 64};
    if (make_tmap_nd_bf16(&te, a->E, 2, dims, strides, box)) return 1;
  }
  F2Params p;
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
    const uint64_t rows = static_cast<uint64_t>(a->B) * a->H * p.tiles_per_head * F2_BM;
    ME_CHECK(rows < (1ull << 31), "me_attention_forward: saved-tile tensor too large");
    const uint64_t dims[2] = {64, rows};
    const uint64_t strides[1] = {64};
    const uint32_t box[2] = {64, F2_BM};
    if (make_tmap_nd_bf16(&tp, a->p_tiles, 2, dims, strides, box)) return 1;
    p.m_tiles = a->m_tiles;
  }
  p.nq = (a->Lq + F2_BM - 1) / F2_BM;
  p.total_units = p.nq * a->B * a->H;
  bool capturing = false;
  p.counters = fwd2_counters(static_cast<cudaStream_t>(a->stream), &capturing);
  p.trace = g_attn_trace;
  if (p.counters == nullptr) return capturing ? -1 : 1;
  const int grid = std::min(sm_count(), p.total_units);
  cudaStream_t st = static_cast<cudaStream_t>(a->stream);
  if (a->flags & ME_ATTN_REF_ROUNDING) {
    if (a->dh == 64) return launch_fwd2<64, true>(tq, tk, tv, te, tp, p, grid, st);
    if (a->dh == 48) return launch_fwd2<48, true>(tq, tk, tv, te, tp, p, grid, st);
    return launch_fwd2<32, true>(tq, tk, tv, te, tp, p, grid, st);
  }
  if (a->dh == 64) return launch_fwd2<64, false>(tq, tk, tv, te, tp, p, grid, st);
  if (a->dh == 48) return launch_fwd2<48, false>(tq, tk, tv, te, tp, p, grid, st);
  return launch_fwd2<32, false>(tq, tk, tv, te, tp, p, grid, st);
}

}  // namespace me
