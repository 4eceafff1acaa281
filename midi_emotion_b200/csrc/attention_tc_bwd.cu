// Relative global attention on the 5th-generation tensor cores -- backward, key side: dK and dV
// (the query side, dQ and dE, is attention_tc_bwd_q.cu; the launcher at the end of this file runs both).
//
// One CTA owns a tile of 64 keys of one (batch, head) and walks the query tiles (128 rows) from the diagonal
// down; dK and dV stay in tensor memory for the whole CTA and are written once -- no cross-CTA reduction.
// Per step, with thread (a, half) owning 32 key columns of query row a == TMEM lane a:
//
//   MMA 1   [S | R] = Q [K ; Eband]^T   [128 x (64 + 192)], one N = 256 instruction per 16 head dims (K and the
//           band of E sit back to back in shared memory; same band as the forward kernel)
//           dP = dO V^T                 [128 x 64]
//   threads x  = S + skew(R)                    Srel[a, b] = R[a, 127 - a + b]
//           P  = exp2(x c - lse)                dS = P (dP - D) / sqrt(dh)       -> shared memory (bf16)
//   MMA 2   dV += P^T dO        dK += dS^T Q    (M = 128 with the 64 keys in lanes 0..63; the other lanes are
//           never read)
//   TMA     the dS tile (already in the UMMA K-major swizzled layout) -> scratch tensor, for the query-side kernel
//
// With the probability tiles the forward pass saved (me_attn_args.p_tiles / m_tiles; template SAVED) MMA 1 shrinks to
// dP and the threads only rescale the stored tile: P = p_saved * exp2(m_saved - lse) -- no QK^T, no band, no skew,
// no exponential per element.
//
// Q, dO and [K ; Eband] are double-buffered, so MMA 1 of step st+1 is issued as soon as the threads have S, R and
// dP of step st in registers and runs while they form P and dS; MMA 2 of step st then runs under the first half of
// step st+1's thread work.
#include <stdlib.h>

#include "attention_tc.cuh"

namespace me {

constexpr int FB_BM = 128;            // query rows per step
constexpr int FB_BN = 64;             // keys per CTA
constexpr int FB_EROWS = 192;
constexpr int FB_COMPUTE_THREADS = 256;  // 8 warps: warp w and w+4 share TMEM lanes 32*(w&3).. and split the key columns
constexpr int FB_CONTROL_WARP = 8;       // MMA issuer, TMEM alloc
constexpr int FB_LOAD_WARP = 9;          // TMA loads
constexpr int FB_THREADS = FB_COMPUTE_THREADS + 64;
constexpr int FB_OFF_V = 0;                     // two V tiles (item parity)
constexpr int FB_OFF_P = FB_OFF_V + 2 * 8192;       // two buffers of [P | dS] (step parity): dS directly behind P, so that
constexpr int FB_PDS_BYTES = 32768;              // [P | dS]^T is one 128-row A operand; the threads of step st+1 write
constexpr int FB_OFF_STAGE = FB_OFF_P + 2 * FB_PDS_BYTES;   // theirs while MMA 2 and the TMA store of step st read the other
// per stage, recompute variant: K | E band | Q | dO   ([K ; Eband] has to be one 256-row B operand, so each stage
//                                                     carries its own copy of the CTA's K tile)
//            saved-P variant:   P tile (as saved by the forward pass) | Q | dO
// [Q | dO] are adjacent in both: one MN-major B operand with N = 128 for the merged dK / dV product
template <bool SAVED> struct FbStage {
  static constexpr int K = 0, E = 8192, P = 0;
  static constexpr int Q = SAVED ? 16384 : 8192 + 24576;
  static constexpr int DO = Q + 16384;
  static constexpr int BYTES = DO + 16384;
  static constexpr int TX = BYTES;                       // bytes one step's loads bring in
  static constexpr int NS = SAVED ? 3 : 2;               // load stages (what fits next to V, P and dS)
  static constexpr int SMEM = FB_OFF_STAGE + NS * BYTES + 1024;   // + barriers
};
static_assert(FB_OFF_P % 1024 == 0 && FB_PDS_BYTES % 1024 == 0 && FB_OFF_STAGE % 1024 == 0 && FbStage<false>::BYTES % 1024 == 0 &&
              FbStage<true>::BYTES % 1024 == 0, "tile alignment");
static_assert(FbStage<false>::SMEM <= 227 * 1024 && FbStage<true>::SMEM <= 227 * 1024, "shared memory budget");
constexpr uint32_t FB_TMEM_COLS = 512;
// [dK | dV] come out of ONE N = 128 product [P | dS]^T [Q | dO]: dK in lanes 64..127 of the first 64 columns, dV in
// lanes 0..63 of the second 64 (N = 64 instructions cost as much as N = 128 ones, profiles/r02_b_micro_mma_rate.txt)
constexpr uint32_t FB_COL_S = 0, FB_COL_R = 64, FB_COL_DP = 256, FB_COL_DK = 320, FB_COL_DV = 384;

struct FbParams {
  int B, H, L, max_seq;
  int64_t k_sb, k_sh, k_sj, v_sb, v_sh, v_sj, keypad_ld;
  const uint8_t* keypad;
  const float* lse;
  const float* dsum;
  bf16* dk;
  bf16* dv;
  float scale_log2, scale;
  int noncausal;     // ME_ATTN_NONCAUSAL: every query tile, every key < L visible
  int tiles_per_head;  // dS scratch: tile (qi, kt) of head (b, h) starts at row ((b H + h) tiles_per_head + index) 128
  int b0, nb;          // this launch's slice of the batch: sequences b0 .. b0 + nb - 1
  const float* m_tiles;  // SAVED: exponent offsets of the saved probability tiles
  long long* trace;      // tuning builds (-DME_ATTN_TRACE)
  int saved_tiles_per_head;
};

extern long long* g_attn_trace;
#ifdef ME_ATTN_TRACE
#define FB_TRACE(role, st, k)                                                                         \
  do {                                                                                                \
    if (tr && (st) < 20) p.trace[((role) * 20 + (st)) * 8 + (k)] = clock64();                         \
  } while (0)
#else
#define FB_TRACE(role, st, k) do { } while (0)
#endif

// One work item = one 64-key tile of one (sequence, head); the kernel is persistent: CTA c works through the items
// c, c + grid, c + 2 grid, ... with every per-step barrier and buffer index running on a global step counter, so the
// loads of the next item are in flight while the current one finishes and nothing is set up or torn down per item
// (the items are short -- 4.5 steps on average at L = 1024 -- and the per-CTA prologue / epilogue of the one-item
// version cost a third of the kernel).
struct FbItem {
  int kt, h, bl, b, j0, qi0, nsteps;
};

template <int DH, bool SAVED>
__global__ void __launch_bounds__(FB_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                   const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmdS, FbParams p) {
  constexpr int HC = DH / 2;          // accumulator columns handled by each of the two threads of a row
  using ST = FbStage<SAVED>;
  extern __shared__ __align__(1024) uint8_t fb_smem[];
  uint8_t* sV = fb_smem + FB_OFF_V;       // two buffers (item parity), 8 KB each
  uint8_t* sPdS = fb_smem + FB_OFF_P;     // buffer (g & 1): P at +0, dS at +16384
  uint8_t* sStage = fb_smem + FB_OFF_STAGE;
  constexpr int NS = ST::NS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(fb_smem + FB_OFF_STAGE + NS * ST::BYTES);
  uint64_t* v_full = bars + 0;    // [2] V of an item (item parity)
  uint64_t* ld_full = bars + 2;   // [NS] per step: (K, the band of E | the saved P tile), Q and dO
  uint64_t* m2_done = bars + 5;   // [NS] dV, dK accumulated: the step's stage, P and dS are free
  uint64_t* m1_done = bars + 8;   // S, R ready
  uint64_t* dp_done = bars + 9;   // dP ready
  uint64_t* a1_done = bars + 10;  // S, R, dP of the step are in registers (256 arrivals)
  uint64_t* a_done = bars + 11;   // [2] P, dS of step g written (barrier g & 1, 256 arrivals)
  uint64_t* ds_free = bars + 13;  // [2] the TMA store of the step's dS tile has read shared memory (barrier g & 1)
  uint64_t* dkv_read = bars + 15; // dK / dV of an item are out of tensor memory (256 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nq = (p.L + FB_BM - 1) / FB_BM;
  const int nkt = (p.L + FB_BN - 1) / FB_BN;
  const int n_items = nkt * p.H * p.nb;
  // item i -> (kt, h, bl): consecutive items are the key tiles of one head (their CTAs run at the same time and share
  // Q / dO through the L2); the key tile is skewed by the head index so that the round-robin over CTAs hands every
  // CTA the same mix of long (early keys) and short (late keys) items
  auto get_item = [&](int i) -> FbItem {
    FbItem it;
    const int grp = i / nkt;
    it.kt = (i % nkt + grp) % nkt;
    it.h = grp % p.H;
    it.bl = grp / p.H;
    it.b = p.b0 + it.bl;
    it.j0 = it.kt * FB_BN;
    it.qi0 = p.noncausal ? 0 : it.j0 / FB_BM;
    it.nsteps = nq - it.qi0;
    return it;
  };
  auto tile_index = [&](const FbItem& it, int st, bool saved_layout) -> int64_t {
    const int qi = it.qi0 + st;
    if (saved_layout)
      return (static_cast<int64_t>(it.b) * p.H + it.h) * p.saved_tiles_per_head +
             (p.noncausal ? static_cast<int64_t>(qi) * nkt : static_cast<int64_t>(qi) * (qi + 1)) + it.kt;
    return (static_cast<int64_t>(it.bl) * p.H + it.h) * p.tiles_per_head +
           (p.noncausal ? static_cast<int64_t>(qi) * nkt : static_cast<int64_t>(qi) * (qi + 1)) + it.kt;
  };
  const bool tr = p.trace != nullptr && blockIdx.x == 0 && lane == 0;
  (void)tr;

  if (tid == 0) {
    if ((smem_u32(fb_smem) & 1023u) != 0) __trap();
    for (int k = 0; k < 2; ++k) {
      mbar_init(&v_full[k], 1);
      mbar_init(&a_done[k], FB_COMPUTE_THREADS);
      mbar_init(&ds_free[k], 1);
    }
    for (int k = 0; k < NS; ++k) {
      mbar_init(&ld_full[k], 1);
      mbar_init(&m2_done[k], 1);
    }
    mbar_init(m1_done, 1);
    mbar_init(dp_done, 1);
    mbar_init(a1_done, FB_COMPUTE_THREADS);
    mbar_init(dkv_read, FB_COMPUTE_THREADS);
    fence_mbar_init();
  }
  if (warp == FB_CONTROL_WARP) tmem_alloc(tmem_slot, FB_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == FB_CONTROL_WARP) {
    // ======================= MMA issuer =======================
    // The whole warp runs this loop converged and one elected lane issues (from a divergent `if (lane == 0)` every
    // tcgen05.mma is wrapped by the compiler in a serialising loop, ~90 cycles per instruction).
    constexpr uint32_t idesc_s = make_idesc_bf16(128, FB_BN, 0, 0);              // dP : K-major x K-major
    constexpr uint32_t idesc_sr = make_idesc_bf16(128, FB_BN + FB_EROWS, 0, 0);  // [S | R]
    constexpr uint32_t idesc_tt = make_idesc_bf16(128, 128, 1, 1);               // [dK | dV] : A^T (MN-major) x B (MN-major)
    auto issue_mma1 = [&](int g, int item_n) {
      const uint32_t base = smem_u32(sStage + (g % NS) * ST::BYTES);
      const uint32_t k_addr = base + ST::K, q_addr = base + ST::Q, do_addr = base + ST::DO;
      const uint32_t v_addr = smem_u32(sV + (item_n & 1) * 8192);
      if (elect_one()) {
        if (!SAVED) {
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)  // [S | R] = Q [K ; Eband]^T
            umma_bf16(tmem_base + FB_COL_S, make_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                      make_smem_desc_sw128(k_addr + k * 32, 16, 1024), idesc_sr, k > 0);
          umma_commit(m1_done);
        }
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)  // dP = dO V^T
          umma_bf16(tmem_base + FB_COL_DP, make_smem_desc_sw128(do_addr + k * 32, 16, 1024),
                    make_smem_desc_sw128(v_addr + k * 32, 16, 1024), idesc_s, k > 0);
        umma_commit(dp_done);
      }
      __syncwarp();
    };
    int g = 0, n = 0;
    // MMA 1 of a step is issued one step ahead of its MMA 2: `pend_*` is the step whose MMA 1 comes next
    int i_next = blockIdx.x;
    FbItem nx = i_next < n_items ? get_item(i_next) : FbItem{};
    int nx_st = 0, nx_n = 0, nx_g = 0;
    auto issue_next_mma1 = [&]() {   // MMA 1 of (item nx_n, step nx_st), global step nx_g; then advance
      if (i_next >= n_items) return;
      if (nx_st == 0) mbar_wait(&v_full[nx_n & 1], (nx_n >> 1) & 1);
      mbar_wait(&ld_full[nx_g % NS], (nx_g / NS) & 1);
      if (nx_g > 0) mbar_wait(a1_done, (nx_g - 1) & 1);   // S, R, dP of the previous step are out of tensor memory
      tc_fence_after();
      issue_mma1(nx_g, nx_n);
      ++nx_g;
      if (++nx_st == nx.nsteps) {
        nx_st = 0;
        ++nx_n;
        i_next += gridDim.x;
        if (i_next < n_items) nx = get_item(i_next);
      }
    };
    issue_next_mma1();
    for (int i = blockIdx.x; i < n_items; i += gridDim.x, ++n) {
      const FbItem it = get_item(i);
      for (int st = 0; st < it.nsteps; ++st, ++g) {
        issue_next_mma1();   // MMA 1 of step g+1 (possibly the first step of the next item)
        FB_TRACE(1, g, 0);
        mbar_wait(&a_done[g & 1], (g >> 1) & 1);
        if (st == 0 && n > 0) mbar_wait(dkv_read, (n - 1) & 1);   // the previous item's dK / dV have been read out
        tc_fence_after();
        FB_TRACE(1, g, 1);
        const uint32_t p_addr = smem_u32(sPdS + (g & 1) * FB_PDS_BYTES);
        const uint32_t q_addr = smem_u32(sStage + (g % NS) * ST::BYTES) + ST::Q;
        const uint32_t acc0 = st > 0 ? 1u : 0u;
        if (elect_one()) {
          // [. | dV ; dK | .] += [P | dS]^T [Q | dO]: rows 0..63 of the accumulator are P^T (keys), rows 64..127
          // dS^T; columns 0..63 multiply Q, columns 64..127 dO.  dK = lanes 64..127 x columns 0..63, dV = lanes
          // 0..63 x columns 64..127; the other two blocks are never read.
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(tmem_base + FB_COL_DK, make_smem_desc_sw128(p_addr + k * 2048, 16384, 1024),
                      make_smem_desc_sw128(q_addr + k * 2048, 16384, 1024), idesc_tt, (k > 0) ? 1u : acc0);
          umma_commit(&m2_done[g % NS]);
        }
        __syncwarp();
        FB_TRACE(1, g, 2);
      }
    }
  } else if (warp == FB_LOAD_WARP) {
    // ======================= TMA loads and the dS stores =======================
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      tma_prefetch_desc(&tmdO);
      tma_prefetch_desc(&tmE);
      tma_prefetch_desc(&tmdS);
    }
    __syncwarp();
    // the loads run NS - 1 steps ahead of the stores: `ld_*` walks the (item, step) sequence for the loads
    int ld_i = blockIdx.x, ld_n = 0, ld_st = 0, ld_g = 0;
    FbItem ld = ld_i < n_items ? get_item(ld_i) : FbItem{};
    int last_g_of_item[2] = {-1, -1};   // global index of the last step of items n-1, n-2 (by parity): V buffer reuse
    auto load_next = [&]() {
      if (ld_i >= n_items) return;
      // stage (ld_g % NS) was last read by the MMAs of step ld_g - NS
      if (ld_g >= NS) mbar_wait(&m2_done[ld_g % NS], ((ld_g - NS) / NS) & 1);
      if (ld_st == 0) {
        // V buffer (ld_n & 1) was last read by dP of the last step of item ld_n - 2
        const int gl = last_g_of_item[ld_n & 1];
        if (ld_n >= 2 && gl > ld_g - NS) mbar_wait(&m2_done[gl % NS], (gl / NS) & 1);
        last_g_of_item[ld_n & 1] = ld_g + ld.nsteps - 1;
      }
      if (elect_one()) {
        if (ld_st == 0) {
          mbar_arrive_expect_tx(&v_full[ld_n & 1], 8192);
          tma_load_4d(&tmV, &v_full[ld_n & 1], sV + (ld_n & 1) * 8192, 0, ld.h, ld.j0, ld.b);
        }
        uint8_t* base = sStage + (ld_g % NS) * ST::BYTES;
        uint64_t* bar = &ld_full[ld_g % NS];
        const int i0 = (ld.qi0 + ld_st) * FB_BM;
        mbar_arrive_expect_tx(bar, ST::TX);
        if (SAVED) {
          tma_load_2d(&tmK, bar, base + ST::P, 0, static_cast<int>(tile_index(ld, ld_st, true) * FB_BM));  // (tmK: saved-P map)
        } else {
          tma_load_4d(&tmK, bar, base + ST::K, 0, ld.h, ld.j0, ld.b);
          tma_load_2d(&tmE, bar, base + ST::E, 0, p.max_seq - FB_BM - (i0 - ld.j0));
        }
        tma_load_4d(&tmQ, bar, base + ST::Q, 0, ld.h, i0, ld.b);
        tma_load_4d(&tmdO, bar, base + ST::DO, 0, ld.h, i0, ld.b);
      }
      __syncwarp();
      ++ld_g;
      if (++ld_st == ld.nsteps) {
        ld_st = 0;
        ++ld_n;
        ld_i += gridDim.x;
        if (ld_i < n_items) ld = get_item(ld_i);
      }
    };
    for (int k = 0; k < NS - 1; ++k) load_next();
    int g = 0;
    for (int i = blockIdx.x; i < n_items; i += gridDim.x) {
      const FbItem it = get_item(i);
      for (int st = 0; st < it.nsteps; ++st, ++g) {
        load_next();    // the loads of step g + NS - 1
        // the step's dS tile -> scratch (the query-side kernel multiplies it by K, E and Q)
        mbar_wait(&a_done[g & 1], (g >> 1) & 1);
        if (elect_one()) {
          tma_store_2d(&tmdS, sPdS + (g & 1) * FB_PDS_BYTES + 16384, 0, static_cast<int>(tile_index(it, st, false) * FB_BM));
          bulk_commit();
          bulk_wait_read_all();
          mbar_arrive(&ds_free[g & 1]);
        }
        __syncwarp();
      }
    }
    if (elect_one()) bulk_wait_all();
    __syncwarp();
  } else {
    // ================================ compute warps ================================
    const int half = warp >> 2, quarter = warp & 3;
    const int a = quarter * 32 + lane;   // query row inside the tile == TMEM lane
    const int shift = 31 - lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const float cs = p.scale_log2;

    // the per-row scalars of a step (log-sum-exp, D, and the exponent offset of the saved tile) come from global
    // memory: they are fetched one step ahead, their latency (~1000 cycles each) used to sit on every step
    float nxt_l = -INFINITY, nxt_D = 0.f, nxt_m = 0.f;
    auto fetch_row_scalars = [&](const FbItem& it, int st) {
      const int i = (it.qi0 + st) * FB_BM + a;
      nxt_l = -INFINITY;
      nxt_D = 0.f;
      nxt_m = 0.f;
      if (i < p.L) {
        const int64_t stat = (static_cast<int64_t>(it.b) * p.H + it.h) * p.L + i;
        nxt_l = p.lse[stat];
        nxt_D = p.dsum[stat];
      }
      if (SAVED) nxt_m = p.m_tiles[tile_index(it, st, true) * FB_BM + a];
    };
    // dK / dV of an item: lanes 0..63 (quarters 0, 1) hold dV rows, lanes 64..127 (quarters 2, 3) dK rows, of key
    // j0 + (a & 63); read out by the threads during the first step of the NEXT item (or after the last one)
    auto write_dkv = [&](const FbItem& it) {
      const bool is_dv = quarter < 2;
      const int j = it.j0 + (a & 63);
      const bool key_ok = j < p.L;
      bf16* row = is_dv ? p.dv + static_cast<int64_t>(it.b) * p.v_sb + static_cast<int64_t>(j) * p.v_sj + it.h * p.v_sh + half * HC
                        : p.dk + static_cast<int64_t>(it.b) * p.k_sb + static_cast<int64_t>(j) * p.k_sj + it.h * p.k_sh + half * HC;
      const uint32_t col = (is_dv ? FB_COL_DV : FB_COL_DK) + half * HC;
#pragma unroll
      for (int c0 = 0; c0 < HC; c0 += 8) {
        uint32_t v8[8];
        tmem_ld8(t_lane + col + c0, v8);
        tc_wait_ld();
        if (key_ok) {
          uint4 u;
          __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(__uint_as_float(v8[2 * e]), __uint_as_float(v8[2 * e + 1]));
          *reinterpret_cast<uint4*>(row + c0) = u;
        }
      }
      tc_fence_before();
      mbar_arrive(dkv_read);
    };

    int g = 0, n = 0;
    FbItem prev{};
    int prev_last_g = -1;
    if (static_cast<int>(blockIdx.x) < n_items) fetch_row_scalars(get_item(blockIdx.x), 0);
    for (int i = blockIdx.x; i < n_items; i += gridDim.x, ++n) {
      const FbItem it = get_item(i);
      const uint8_t* kp = p.keypad ? p.keypad + static_cast<int64_t>(it.b) * p.keypad_ld : nullptr;
      uint32_t kpm = 0;  // key-pad bits of this thread's 32 keys
      if (kp || p.noncausal) {  // (keys past the sequence only matter without the causal predicate)
        const int j = it.j0 + 32 * half + lane;
        kpm = __ballot_sync(0xffffffffu, j >= p.L || (kp && kp[j] != 0));
      }
      for (int st = 0; st < it.nsteps; ++st, ++g) {
        const uint32_t ph = g & 1;
        const int i0 = (it.qi0 + st) * FB_BM;
        const int irow = i0 + a;
        const bool row_ok = irow < p.L;
        const float l_nat = nxt_l, Di = nxt_D, m_saved = nxt_m;
        if (st + 1 < it.nsteps) fetch_row_scalars(it, st + 1);
        else if (i + static_cast<int>(gridDim.x) < n_items) fetch_row_scalars(get_item(i + gridDim.x), 0);
        const int lim = p.noncausal ? 31 : irow - it.j0 - 32 * half;  // this thread's columns bb <= lim are causal-visible
        uint32_t vm = lim >= 31 ? 0xffffffffu : (lim < 0 ? 0u : ((2u << lim) - 1u));
        vm &= ~kpm;

        const float lse2 = (!row_ok || l_nat == -INFINITY) ? INFINITY : l_nat * 1.4426950408889634f;
        if (warp == 0) FB_TRACE(0, g, 0);

        uint32_t pw[16], dw[16];  // bf16x2 words of P[a, 32h..] and dS[a, 32h..]
        {
          float pe[32];
          if (SAVED) {
            // P = p_saved * exp2(m_saved - lse): the forward pass formed p_saved = exp2(x c - m_saved), masks included
            const float cfac = fast_exp2(m_saved - lse2);
            mbar_wait(&ld_full[g % NS], (g / NS) & 1);
            const uint8_t* prow_in = sStage + (g % NS) * ST::BYTES + ST::P + a * 128;
#pragma unroll
            for (int nn = 0; nn < 4; ++nn) {
              const uint4 u = *reinterpret_cast<const uint4*>(prow_in + (((4 * half + nn) ^ (a & 7)) << 4));
              const uint32_t wv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                pe[8 * nn + 2 * e] = __uint_as_float(wv[e] << 16) * cfac;
                pe[8 * nn + 2 * e + 1] = __uint_as_float(wv[e] & 0xFFFF0000u) * cfac;
              }
            }
            (void)vm;
            (void)cs;
            (void)shift;
          } else {
            mbar_wait(m1_done, ph);
            tc_fence_after();
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
          }
#pragma unroll
          for (int bb = 0; bb < 32; bb += 2) {
            __nv_bfloat162 ph2 = __floats2bfloat162_rn(pe[bb], pe[bb + 1]);
            pw[bb / 2] = *reinterpret_cast<uint32_t*>(&ph2);
          }
          if (warp == 0) FB_TRACE(0, g, 1);
          mbar_wait(dp_done, ph);  // dP was issued behind [S | R]: it lands while P is being formed
          tc_fence_after();
          if (warp == 0) FB_TRACE(0, g, 2);
          {
            uint32_t dpv[32];
            tmem_ld32(t_lane + FB_COL_DP + 32 * half, dpv);
            tc_wait_ld();
            tc_fence_before();
            mbar_arrive(a1_done);  // S, R, dP are in registers: MMA 1 of the next step may overwrite them
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
        if (warp == 0) FB_TRACE(0, g, 3);
        // the previous item's dK / dV leave tensor memory before this item's first MMA 2 overwrites them
        if (st == 0 && n > 0) {
          mbar_wait(&m2_done[prev_last_g % NS], (prev_last_g / NS) & 1);
          tc_fence_after();
          write_dkv(prev);
        }
        // MMA 2 and the TMA store of step g-2 must have read this buffer of P / dS
        if (g > 1) {
          mbar_wait(&m2_done[(g - 2) % NS], ((g - 2) / NS) & 1);
          mbar_wait(&ds_free[g & 1], ((g - 2) >> 1) & 1);
        }
        // P and dS rows: UMMA SWIZZLE_128B rows of 128 B (chunk kc of row a at position kc ^ (a & 7))
        {
          uint8_t* prow = sPdS + (g & 1) * FB_PDS_BYTES + a * 128;
          uint8_t* drow = prow + 16384;
#pragma unroll
          for (int nn = 0; nn < 4; ++nn) {
            const int pos = (((4 * half + nn) ^ (a & 7))) << 4;
            *reinterpret_cast<uint4*>(prow + pos) = make_uint4(pw[4 * nn], pw[4 * nn + 1], pw[4 * nn + 2], pw[4 * nn + 3]);
            *reinterpret_cast<uint4*>(drow + pos) = make_uint4(dw[4 * nn], dw[4 * nn + 1], dw[4 * nn + 2], dw[4 * nn + 3]);
          }
        }
        if (warp == 0) FB_TRACE(0, g, 4);
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&a_done[g & 1]);
        if (warp == 0) FB_TRACE(0, g, 5);
      }
      prev = it;
      prev_last_g = g - 1;
    }
    if (n > 0) {   // the last item
      mbar_wait(&m2_done[prev_last_g % NS], (prev_last_g / NS) & 1);
      tc_fence_after();
      write_dkv(prev);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == FB_CONTROL_WARP) tmem_dealloc(tmem_base, FB_TMEM_COLS);
}

// dsum[b, h, i] = sum_c dO[b, i, h, c] * O[b, i, h, c]   (the "D" term of the softmax backward)
// Eight lanes per (row, head): lane c multiplies one 16-byte chunk of O with the matching chunk of dO (consecutive
// lanes read consecutive 16 bytes: whole 128-byte lines per head), three shuffles add the partial dot products.
template <int DH>
__global__ void attn_bwd_prep_kernel(const bf16* __restrict__ out, const bf16* __restrict__ dout, int64_t o_sb,
                                     int64_t o_si, int B, int H, int L, float* __restrict__ dsum) {
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int c = static_cast<int>(t & 7);
  const int64_t idx = t >> 3;
  const int64_t total = static_cast<int64_t>(B) * L * H;
  float acc = 0.f;
  int h = 0, i = 0, b = 0;
  if (idx < total) {
    h = static_cast<int>(idx % H);
    const int64_t bi = idx / H;
    i = static_cast<int>(bi % L);
    b = static_cast<int>(bi / L);
    if (c * 8 < DH) {
      const int64_t off = b * o_sb + i * o_si + h * DH + c * 8;
      const uint4 uo = __ldg(reinterpret_cast<const uint4*>(out + off));
      const uint4 ug = __ldg(reinterpret_cast<const uint4*>(dout + off));
      const __nv_bfloat162* ho = reinterpret_cast<const __nv_bfloat162*>(&uo);
      const __nv_bfloat162* hg = reinterpret_cast<const __nv_bfloat162*>(&ug);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 fo = __bfloat1622float2(ho[e]), fg = __bfloat1622float2(hg[e]);
        acc = fmaf(fo.x, fg.x, acc);
        acc = fmaf(fo.y, fg.y, acc);
      }
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (c == 0 && idx < total) dsum[(static_cast<int64_t>(b) * H + h) * L + i] = acc;
}

// dE[e, :] += sum over the private copies of dE_ws[., e, :]
__global__ void attn_bwd_finish_kernel(int dh, const float* __restrict__ dE_ws, float* __restrict__ dE, int max_seq) {
  const int q4 = dh / 4;
  const int64_t n_de = static_cast<int64_t>(max_seq) * q4;
  for (int64_t u = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; u < n_de;
       u += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(u % q4) * 4;
    const int64_t e = u / q4;
    const float* src = dE_ws + e * dh + c;
    float4 v[FB_DE_COPIES];
#pragma unroll
    for (int cp = 0; cp < FB_DE_COPIES; ++cp)
      v[cp] = *reinterpret_cast<const float4*>(src + static_cast<int64_t>(cp) * max_seq * dh);
    float4 o = *reinterpret_cast<float4*>(dE + e * dh + c);
#pragma unroll
    for (int cp = 0; cp < FB_DE_COPIES; ++cp) { o.x += v[cp].x; o.y += v[cp].y; o.z += v[cp].z; o.w += v[cp].w; }
    *reinterpret_cast<float4*>(dE + e * dh + c) = o;
  }
}

// tk: the K map (recompute) or the saved-P tile map (SAVED)
template <int DH, bool SAVED>
static int launch_bwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& tdo,
                      const CUtensorMap& te, const CUtensorMap& tds, const FbParams& p, dim3 grid, cudaStream_t st) {
  auto kern = attn_bwd_tc_kernel<DH, SAVED>;
  constexpr int smem = FbStage<SAVED>::SMEM;
  static bool configured = false;
  if (!configured) {
    ME_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  cudaEvent_t pe = prof_begin(3.0 * attn_unit_flops(p.nb, p.H, p.L, DH), st, 2);   // dP, dV, dK
  kern<<<grid, FB_THREADS, smem, st>>>(tq, tk, tv, tdo, te, tds, p);
  prof_end(pe, st);
  ME_LAUNCH_CHECK();
  return 0;
}

int launch_attn_bwd_q_tc(const me_attn_bwd_args* ba, float* dE_ws, const CUtensorMap& tds, int tiles_per_head, int b0,
                         int nb, int32_t* sched_dev);
int64_t attn_bwd_q_sched_words(int L, int nb);

// dS scratch: tiles of 128 query rows x 64 keys, per head nq * nkt + nq of them (enough for the causal and the
// non-causal indexing)
static inline int64_t ds_tiles_per_head(int L) {
  const int64_t nq = (L + FB_BM - 1) / FB_BM, nkt = (L + FB_BN - 1) / FB_BN;
  return nq * nkt + nq;
}
// The two kernels can run over slices of the batch small enough for a slice's dS tiles to stay in the L2 between them
// (ME_DS_SLICE_MB; the scratch is sized for one slice).  Measured on the B200 at cfg2 the wave tails of the smaller
// launches cost more than the L2 hits save (0.69 ms unsliced, 0.75 / 0.88 ms at 200 / 64 MB), so the default is one
// slice and the hand-over goes through HBM.
static int64_t ds_slice_bytes() {
  static int64_t v = 0;
  if (v == 0) {
    const char* e = getenv("ME_DS_SLICE_MB");   // tuning knob; 0 / unset = default
    const long mb = e ? atol(e) : 0;
    v = (mb > 0 ? mb : (1l << 20)) << 20;   // default: one slice (slicing costs more in wave tails than it saves, profiles/)
  }
  return v;
}
static inline int ds_slice_batch(int B, int H, int L) {
  const int64_t nq = (L + FB_BM - 1) / FB_BM;
  const int64_t touched = static_cast<int64_t>(H) * nq * (nq + 1) * FB_BM * 64 * 2;   // bytes per sequence (causal)
  int64_t nb = ds_slice_bytes() / (touched > 0 ? touched : 1);
  if (nb < 1) nb = 1;
  return static_cast<int>(nb < B ? nb : B);
}
static inline int64_t de_ws_floats(int dh, int max_seq) { return static_cast<int64_t>(FB_DE_COPIES) * max_seq * dh; }

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
  ME_CHECK(a->q_si % 8 == 0 && a->q_sb % 8 == 0 && (reinterpret_cast<uintptr_t>(ba->dq) & 15) == 0,
           "me_attention_backward: dq rows must be 16-byte aligned");
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
  // D = rowsum(dO * O); zero the private dE accumulators
  {
    const int64_t total = static_cast<int64_t>(B) * L * H * 8;   // eight lanes per (row, head)
    const int blocks = static_cast<int>((total + 255) / 256);
    const bf16* o = static_cast<const bf16*>(a->out);
    const bf16* g = static_cast<const bf16*>(ba->dout);
    if (dh == 64) attn_bwd_prep_kernel<64><<<blocks, 256, 0, st>>>(o, g, a->o_sb, a->o_si, B, H, L, ba->dsum);
    else if (dh == 48) attn_bwd_prep_kernel<48><<<blocks, 256, 0, st>>>(o, g, a->o_sb, a->o_si, B, H, L, ba->dsum);
    else attn_bwd_prep_kernel<32><<<blocks, 256, 0, st>>>(o, g, a->o_sb, a->o_si, B, H, L, ba->dsum);
    ME_LAUNCH_CHECK();
    ME_CUDA(cudaMemsetAsync(ba->dq_acc, 0, sizeof(float) * static_cast<size_t>(FB_DE_COPIES) * a->max_seq * dh, st));
  }
  // the dS scratch sits behind the private dE accumulators in the caller's workspace
  const int tph = static_cast<int>(ds_tiles_per_head(L));
  const int slice = ds_slice_batch(B, H, L);
  const int64_t ds_floats = static_cast<int64_t>(slice) * H * tph * FB_BM * 64 / 2;
  int32_t* sched_dev = reinterpret_cast<int32_t*>(ba->dq_acc + de_ws_floats(dh, a->max_seq) + ds_floats);
  CUtensorMap tds;
  {
    const uint64_t rows = static_cast<uint64_t>(slice) * H * tph * FB_BM;
    ME_CHECK(rows < (1ull << 31), "me_attention_backward: dS scratch too large");
    const uint64_t dims[2] = {64, rows};
    const uint64_t strides[1] = {64};
    const uint32_t box[2] = {64, FB_BM};
    if (make_tmap_nd_bf16(&tds, ba->dq_acc + de_ws_floats(dh, a->max_seq), 2, dims, strides, box)) return 1;
  }
  // probability tiles saved by the forward pass (optional)
  const bool saved = a->p_tiles != nullptr && a->m_tiles != nullptr;
  CUtensorMap tps = tk;
  FbParams p;
  p.m_tiles = nullptr;
  p.saved_tiles_per_head = static_cast<int>(me_attention_saved_tiles(L, a->flags));
  if (saved) {
    const uint64_t rows = static_cast<uint64_t>(B) * H * p.saved_tiles_per_head * FB_BM;
    ME_CHECK(rows < (1ull << 31), "me_attention_backward: saved-tile tensor too large");
    const uint64_t dims[2] = {64, rows};
    const uint64_t strides[1] = {64};
    const uint32_t box[2] = {64, FB_BM};
    if (make_tmap_nd_bf16(&tps, a->p_tiles, 2, dims, strides, box)) return 1;
    p.m_tiles = a->m_tiles;
  }
  p.tiles_per_head = tph;
  p.trace = g_attn_trace;
  p.B = B; p.H = H; p.L = L; p.max_seq = a->max_seq;
  p.k_sb = a->k_sb; p.k_sh = a->k_sh; p.k_sj = a->k_sj;
  p.v_sb = a->v_sb; p.v_sh = a->v_sh; p.v_sj = a->v_sj;
  p.keypad_ld = a->keypad_ld; p.keypad = a->keypad;
  p.lse = a->lse; p.dsum = ba->dsum;
  p.dk = static_cast<bf16*>(ba->dk); p.dv = static_cast<bf16*>(ba->dv);
  p.scale = 1.f / sqrtf(static_cast<float>(dh));
  p.scale_log2 = 1.4426950408889634f * p.scale;
  p.noncausal = (a->flags & ME_ATTN_NONCAUSAL) ? 1 : 0;
  for (int b0 = 0; b0 < B; b0 += slice) {
    const int nb = B - b0 < slice ? B - b0 : slice;
    p.b0 = b0;
    p.nb = nb;
    const int n_items = ((L + FB_BN - 1) / FB_BN) * H * nb;
    dim3 grid(n_items < sm_count() ? n_items : sm_count(), 1, 1);
    int rc;
    if (saved) {
      if (dh == 64) rc = launch_bwd<64, true>(tq, tps, tv, tdo, te, tds, p, grid, st);
      else if (dh == 48) rc = launch_bwd<48, true>(tq, tps, tv, tdo, te, tds, p, grid, st);
      else rc = launch_bwd<32, true>(tq, tps, tv, tdo, te, tds, p, grid, st);
    } else {
      if (dh == 64) rc = launch_bwd<64, false>(tq, tk, tv, tdo, te, tds, p, grid, st);
      else if (dh == 48) rc = launch_bwd<48, false>(tq, tk, tv, tdo, te, tds, p, grid, st);
      else rc = launch_bwd<32, false>(tq, tk, tv, tdo, te, tds, p, grid, st);
    }
    if (rc) return rc;
    if (launch_attn_bwd_q_tc(ba, ba->dq_acc, tds, tph, b0, nb, sched_dev)) return 1;
  }
  {
    const int blocks = static_cast<int>((static_cast<int64_t>(a->max_seq) * (dh / 4) + 255) / 256);
    attn_bwd_finish_kernel<<<blocks, 256, 0, st>>>(dh, ba->dq_acc, ba->dE, a->max_seq);
    ME_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace me

extern "C" int me_debug_trace_set(long long* device_buf) {
  me::g_attn_trace = device_buf;   // read by tuning builds (-DME_ATTN_TRACE) of the query-side backward kernel
  return 0;
}

// Scratch of the tensor-core backward, in floats: the private dE accumulators, then the bf16 dS tiles that travel
// from the key-side to the query-side kernel (128 x 64 x 2 bytes each).  (The layer code also parks the partial rows
// of its bias-gradient column sums here -- 160 rows of at most 4 H dh columns -- which always fits.)
extern "C" int64_t me_attention_backward_workspace_floats(int B, int H, int L, int dh, int max_seq) {
  const int64_t de = me::de_ws_floats(dh, max_seq);
  const int nbs = me::ds_slice_batch(B, H, L);
  const int64_t ds = static_cast<int64_t>(nbs) * H * me::ds_tiles_per_head(L) * me::FB_BM * 64 / 2 +
                     me::attn_bwd_q_sched_words(L, nbs);   // + the query-side kernel's unit schedule (int32 words)
  const int64_t colsum = static_cast<int64_t>(160) * 4 * H * dh;
  return de + ds > colsum ? de + ds : colsum;
}
