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
//
// Persistent: a work unit is (query tile, sequence) over ALL heads, units are handed to the CTAs by a host-side
// longest-first schedule, and every barrier / buffer index runs on a global step counter, so the loads of the next head
// are in flight while the current one finishes.  The first five pairs of dE rows of a unit stay in tensor memory and
// accumulate over the unit's heads (E is shared by the heads, music_multi.py:185) -- they are reduce-added once per
// unit instead of once per head, which cuts the fp32 reduce-add volume, the next limit of this kernel, four- to
// twelve-fold; later pairs go through one temporary accumulator per head.
#include <string.h>

#include <algorithm>
#include <vector>

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
constexpr int QB_OFF_Q = 0;                      // two Q tiles (head parity)
constexpr int QB_OFF_STAGE = QB_OFF_Q + 2 * 16384;
constexpr int QB_OFF_T = QB_OFF_STAGE + QB_STAGES * QB_STAGE_BYTES;   // 128 x 64 bf16 panels
constexpr int QB_OFF_BAR = QB_OFF_T + QB_TSLOTS * 16384;
constexpr int QB_SMEM = QB_OFF_BAR + 256;
static_assert(QB_OFF_STAGE % 1024 == 0 && QB_OFF_T % 1024 == 0 && QB_OFF_BAR % 1024 == 0, "tile alignment");
static_assert(QB_SMEM <= 227 * 1024, "shared memory budget");
constexpr uint32_t QB_TMEM_COLS = 512;
constexpr int QB_NRES = 5;                       // pairs of dE rows resident in tensor memory over a unit's heads
// dQ accumulators of two consecutive heads (head parity), the temporary dE pair, the resident dE pairs
constexpr uint32_t QB_COL_DQ = 0, QB_COL_TMP = 128, QB_COL_RES = 192;

long long* g_attn_trace = nullptr;   // me_debug_trace_set (no trace points in this kernel at present)

struct QbParams {
  int B, H, L, max_seq;
  int64_t q_sb, q_sh, q_si;
  float* dE_ws;  // fp32 [FB_DE_COPIES, max_seq, dh]
  int b0;        // first sequence of this launch's slice of the batch
  bf16* dq;
  int noncausal;
  int tiles_per_head;  // dS scratch: tile (qi, kt) of head (b, h) starts at row ((b H + h) tiles_per_head + index) 128
  const int32_t* sched;      // [gridDim.x, sched_pitch]: work units (qi | bl << 8) of every CTA, longest first
  const int32_t* sched_cnt;  // [gridDim.x]
  int sched_pitch;
};

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
  // per-step barriers come in threes (index g % 3): a waiter can then never be a whole phase behind the barrier
  // it polls, however far the loads run ahead
  uint64_t* q_full = bars + 0;     // [2] Q of a head (head parity)
  uint64_t* ld_full = bars + 2;    // [3]: dS tile, K tile and the chunk of E of step g
  uint64_t* a_done = bars + 5;     // [3] the T chunks of step g are in shared memory (128 arrivals)
  uint64_t* mma_done = bars + 8;   // [3] every MMA of step g has retired
  uint64_t* de_read = bars + 11;   // the temporary dE pair is out of tensor memory (128 arrivals)
  uint64_t* res_read = bars + 12;  // the resident dE pairs of a unit are out of tensor memory (128 arrivals)
  uint64_t* dq_read = bars + 13;   // [2] dQ of a head is out of tensor memory (head parity, 128 arrivals)
  // completion barriers of their own for the drain warps (they may fall several steps behind; the issuer's waits on
  // de_read / res_read keep each of these at most one phase ahead of its reader)
  uint64_t* tmp_done = bars + 15;  // the MMAs of a step that used the temporary dE pair have retired
  uint64_t* unit_done = bars + 16; // the MMAs of the last step of a unit have retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_units = p.sched_cnt[blockIdx.x];
  const int32_t* my_units = p.sched + static_cast<int64_t>(blockIdx.x) * p.sched_pitch;
  const int nkt = (p.L + QB_BN - 1) / QB_BN;
  float* const dE_mine = p.dE_ws + static_cast<int64_t>(blockIdx.x % FB_DE_COPIES) * p.max_seq * DH;

  struct Unit {
    int qi, bl, b, i0, nt, t_end, e_base;
  };
  auto get_unit = [&](int u) -> Unit {
    Unit un;
    const int32_t w = my_units[u];
    un.qi = w & 0xFF;
    un.bl = w >> 8;
    un.b = p.b0 + un.bl;
    un.i0 = un.qi * QB_BM;
    const int kmax = p.noncausal ? p.L : min(un.i0 + QB_BM, p.L);
    un.nt = (kmax + QB_BN - 1) / QB_BN;        // real steps (key tiles)
    // chunks of T that can hold a non-zero value AND meet a row of E below max_seq: c <= min(nt + 1, 2 qi + 1);
    // steps nt .. t_end-1 only drain them.  t_end is even so that the last pair is complete.
    const int c_last = max(un.nt - 1, min(un.nt + 1, 2 * un.qi + 1));
    un.t_end = (c_last + 2) & ~1;
    un.e_base = p.max_seq - QB_BM - un.i0;
    return un;
  };

  if (tid == 0) {
    if ((smem_u32(qb_smem) & 1023u) != 0) __trap();
    for (int k = 0; k < 2; ++k) {
      mbar_init(&q_full[k], 1);
      mbar_init(&dq_read[k], QB_ROW_THREADS);
    }
    for (int k = 0; k < 3; ++k) {
      mbar_init(&ld_full[k], 1);
      mbar_init(&a_done[k], QB_ROW_THREADS);
      mbar_init(&mma_done[k], 1);
    }
    mbar_init(de_read, QB_ROW_THREADS);
    mbar_init(res_read, QB_ROW_THREADS);
    mbar_init(tmp_done, 1);
    mbar_init(unit_done, 1);
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
    int g = 0, hn = 0, tmp_uses = 0;
    for (int u = 0; u < n_units; ++u) {
      const Unit un = get_unit(u);
      for (int h = 0; h < p.H; ++h, ++hn) {
        const uint32_t q_addr = smem_u32(sQ + (hn & 1) * 16384);
        const uint32_t dq_col = tmem_base + QB_COL_DQ + 64 * (hn & 1);
        mbar_wait(&q_full[hn & 1], (hn >> 1) & 1);
        if (hn >= 2) mbar_wait(&dq_read[hn & 1], ((hn - 2) >> 1) & 1);   // dQ of head hn-2 has left this accumulator
        uint32_t dq_acc = 0;
        for (int t = 0; t < un.t_end; ++t, ++g) {
          const int s3 = g % 3;
          const uint32_t ph3 = (g / 3) & 1;
          mbar_wait(&ld_full[s3], ph3);   // dS, K, E chunk of the step (dS was also read by the threads)
          mbar_wait(&a_done[s3], ph3);
          const int f = (t - 1) >> 1;     // pair finished by an odd step
          const bool resident = f < QB_NRES;
          if (t & 1) {
            if (resident) {
              if (h == 0 && u > 0 && f == 0) mbar_wait(res_read, (u - 1) & 1);  // the previous unit's pairs are out
            } else {
              if (tmp_uses > 0) mbar_wait(de_read, (tmp_uses - 1) & 1);
              ++tmp_uses;
            }
          }
          tc_fence_after();
          const uint32_t st_addr = smem_u32(sStage + s3 * QB_STAGE_BYTES);
          const uint32_t k_addr = st_addr + QB_STAGE_K, ds_addr = st_addr + QB_STAGE_DS, e_addr = st_addr + QB_STAGE_E;
          const uint32_t t_addr = smem_u32(sT + (g % QB_TSLOTS) * 16384);
          if (elect_one()) {
            const bool real = t < un.nt;
            const uint32_t pair_addr = smem_u32(sT + ((g - 1) % QB_TSLOTS) * 16384);
            const uint32_t de_col = tmem_base + (resident ? QB_COL_RES + 64 * f : QB_COL_TMP);
            const uint32_t de_acc0 = (resident && h > 0) ? 1u : 0u;
            // dQ += dS K_t (real steps); dQ += T_t E_t (chunk t is complete); odd steps: the pair of dE rows
            // [T_{t-1} | T_t]^T Q -- issued interleaved
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (real) {
                umma_bf16(dq_col, make_smem_desc_sw128(ds_addr + k * 32, 16, 1024),
                          make_smem_desc_sw128(k_addr + k * 2048, 8192, 1024), idesc_nt, (k > 0) ? 1u : dq_acc);
                if (k == 0) dq_acc = 1;
              }
              umma_bf16(dq_col, make_smem_desc_sw128(t_addr + k * 32, 16, 1024),
                        make_smem_desc_sw128(e_addr + k * 2048, 8192, 1024), idesc_nt, (k > 0) ? 1u : dq_acc);
              if (k == 0) dq_acc = 1;
              if (t & 1) {
#pragma unroll
                for (int kk = 2 * k; kk < 2 * k + 2; ++kk)
                  umma_bf16(de_col, make_smem_desc_sw128(pair_addr + kk * 2048, 16384, 1024),
                            make_smem_desc_sw128(q_addr + kk * 2048, 8192, 1024), idesc_tt, (kk > 0) ? 1u : de_acc0);
              }
            }
            umma_commit(&mma_done[s3]);
            if ((t & 1) && !resident) umma_commit(tmp_done);
            if (h == p.H - 1 && t == un.t_end - 1) umma_commit(unit_done);
          }
          dq_acc = 1;
          __syncwarp();
        }
      }
    }
  } else if (warp == QB_LOAD_WARP) {
    // ======================= TMA loads =======================
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmdS);
      tma_prefetch_desc(&tmE64);
    }
    __syncwarp();
    int g = 0, hn = 0;
    int last_g_of_head[2] = {-1, -1};
    for (int u = 0; u < n_units; ++u) {
      const Unit un = get_unit(u);
      for (int h = 0; h < p.H; ++h, ++hn) {
        const int64_t ds_row0 = ((static_cast<int64_t>(un.bl) * p.H + h) * p.tiles_per_head +
                                 (p.noncausal ? un.qi * nkt : un.qi * (un.qi + 1))) * QB_BM;
        for (int t = 0; t < un.t_end; ++t, ++g) {
          // stage g % 3 was last read by the MMAs of step g - 3
          if (g >= 3) mbar_wait(&mma_done[g % 3], ((g - 3) / 3) & 1);
          if (t == 0) {
            // Q buffer (hn & 1) was last read by the dE MMAs of head hn - 2
            const int gl = last_g_of_head[hn & 1];
            if (hn >= 2 && gl > g - 3) mbar_wait(&mma_done[gl % 3], (gl / 3) & 1);
            last_g_of_head[hn & 1] = g + un.t_end - 1;
          }
          if (elect_one()) {
            if (t == 0) {
              mbar_arrive_expect_tx(&q_full[hn & 1], 16384);
              tma_load_4d(&tmQ, &q_full[hn & 1], sQ + (hn & 1) * 16384, 0, h, un.i0, un.b);
            }
            const int s3 = g % 3;
            uint8_t* st = sStage + s3 * QB_STAGE_BYTES;
            const bool real = t < un.nt;
            mbar_arrive_expect_tx(&ld_full[s3], 8192 + (real ? 16384 + 8192 : 0));
            if (real) {
              tma_load_2d(&tmdS, &ld_full[s3], st + QB_STAGE_DS, 0, static_cast<int>(ds_row0 + static_cast<int64_t>(t) * QB_BM));
              tma_load_4d(&tmK, &ld_full[s3], st + QB_STAGE_K, 0, h, t * QB_BN, un.b);
            }
            tma_load_2d(&tmE64, &ld_full[s3], st + QB_STAGE_E, 0, un.e_base + 64 * t);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    // ================================ dE drain (warps 4-7) ================================
    const int a = tid - QB_ROW_THREADS;          // row of a pair == TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp - 4) * 32) << 16);
    auto drain_pair = [&](uint32_t col, uint64_t* done_bar, bool arrive, int row) {
      uint32_t v[DH];
      tmem_ld_cols<DH>(t_lane + col, v);
      tc_wait_ld();
      if (arrive) {
        tc_fence_before();
        mbar_arrive(done_bar);
      }
      if (row < p.max_seq) {
        float* dst = dE_mine + static_cast<int64_t>(row) * DH;
#pragma unroll
        for (int c = 0; c < DH; c += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c), "f"(__uint_as_float(v[c])),
                       "f"(__uint_as_float(v[c + 1])), "f"(__uint_as_float(v[c + 2])), "f"(__uint_as_float(v[c + 3]))
                       : "memory");
      }
    };
    int tmp_idx = 0;
    for (int u = 0; u < n_units; ++u) {
      const Unit un = get_unit(u);
      const int npairs = un.t_end / 2;
      for (int h = 0; h < p.H; ++h) {
        for (int f = QB_NRES; f < npairs; ++f, ++tmp_idx) {   // pairs that went through the temporary accumulator
          mbar_wait(tmp_done, tmp_idx & 1);
          tc_fence_after();
          drain_pair(QB_COL_TMP, de_read, true, un.e_base + 128 * f + a);
        }
      }
      // the unit's resident pairs: complete after its last step
      mbar_wait(unit_done, u & 1);
      tc_fence_after();
      const int nres = npairs < QB_NRES ? npairs : QB_NRES;
      for (int f = 0; f < nres; ++f) drain_pair(QB_COL_RES + 64 * f, res_read, f == nres - 1, un.e_base + 128 * f + a);
    }
  } else {
    // ================================ T assembly (warps 0-3) ================================
    const int a = tid;                           // query row inside the tile == TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    // the row's window of step t covers the band columns [64 t + base_g, 64 t + base_g + 64): `cl` values carried over
    // from the previous tile followed by the first 64 - cl values of this one
    const int cl = (127 - a) & 7;
    const int base_q = (127 - a) >> 3;           // base_g / 8: first 16-byte chunk of the window (step 0)

    // dQ of a finished head: complete in tensor memory, written once
    auto write_dq = [&](const Unit& un, int h, int hn) {
      const int i = un.i0 + a;
      const bool row_ok = i < p.L;
      bf16* dqrow = p.dq + static_cast<int64_t>(un.b) * p.q_sb + static_cast<int64_t>(i) * p.q_si + h * p.q_sh;
#pragma unroll
      for (int c0 = 0; c0 < DH; c0 += 8) {
        uint32_t vq[8];
        tmem_ld8(t_lane + QB_COL_DQ + 64 * (hn & 1) + c0, vq);   // (warp-collective: outside the row predicate)
        tc_wait_ld();
        uint4 uu;
        __nv_bfloat162* hq = reinterpret_cast<__nv_bfloat162*>(&uu);
#pragma unroll
        for (int e = 0; e < 4; ++e) hq[e] = __floats2bfloat162_rn(__uint_as_float(vq[2 * e]), __uint_as_float(vq[2 * e + 1]));
        if (row_ok) *reinterpret_cast<uint4*>(dqrow + c0) = uu;
      }
      tc_fence_before();
      mbar_arrive(&dq_read[hn & 1]);
    };

    int g = 0, hn = 0;
    Unit prev_un{};
    int prev_h = -1, prev_last_g = -1;
    for (int u = 0; u < n_units; ++u) {
      const Unit un = get_unit(u);
      for (int h = 0; h < p.H; ++h, ++hn) {
        const int g_head = g;                        // panels of this head: global index g_head + panel
        auto chunk_ptr = [&](int q) -> uint8_t* {    // q: 16-byte chunk index along the band (8 columns each)
          return sT + ((g_head + (q >> 3)) % QB_TSLOTS) * 16384 + a * 128 + (((q & 7) ^ (a & 7)) << 4);
        };
        uint32_t carry[4] = {0u, 0u, 0u, 0u};        // the previous tile's last eight values of this row
        bool lead_zeroed = false;
        for (int t = 0; t < un.t_end; ++t, ++g) {
          const bool real = t < un.nt;
          if (real) {
            mbar_wait(&ld_full[g % 3], (g / 3) & 1);
            if (!lead_zeroed) {
              // band columns below the row's first window never get a value: zero them once per head (panels 0, 1:
              // their slots were last read by MMAs at least four steps back -- the ld_full wait above implies the
              // MMAs of step g-3 -- every later panel is written whole before it is read)
              for (int q = 0; q < base_q; ++q) *reinterpret_cast<uint4*>(chunk_ptr(q)) = make_uint4(0, 0, 0, 0);
              lead_zeroed = true;
            }
            uint32_t w[36];
#pragma unroll
            for (int c = 0; c < 4; ++c) w[c] = carry[c];
            const uint8_t* drow = sStage + (g % 3) * QB_STAGE_BYTES + QB_STAGE_DS + a * 128;
#pragma unroll
            for (int n = 0; n < 8; ++n) {
              const uint4 uu = *reinterpret_cast<const uint4*>(drow + ((n ^ (a & 7)) << 4));
              w[4 + 4 * n] = uu.x; w[5 + 4 * n] = uu.y; w[6 + 4 * n] = uu.z; w[7 + 4 * n] = uu.w;
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
#pragma unroll
            for (int n = 0; n < 8; ++n)
              *reinterpret_cast<uint4*>(chunk_ptr(q0 + n)) = make_uint4(w[4 + 4 * n], w[5 + 4 * n], w[6 + 4 * n], w[7 + 4 * n]);
            if (t == un.nt - 1) {
              // end of the row: the carried tail (cl values, then zeros) and zeros up to the end of the last panel the
              // MMAs will read -- those panels still hold chunks of six steps ago
              uint32_t x[8] = {carry[0], carry[1], carry[2], carry[3], 0u, 0u, 0u, 0u};
              const int sh = 8 - cl;                 // move elements 8-cl..7 to 0..cl-1
              const uint32_t l4 = sh & 4, l2 = sh & 2, l8 = sh & 8;
#pragma unroll
              for (int k = 0; k < 4; ++k) x[k] = sel_b32(x[k + 2], x[k], l4);
#pragma unroll
              for (int k = 0; k < 4; ++k) x[k] = sel_b32(x[k + 1], x[k], l2);
              const uint32_t hl = (sh & 1) ? 16u : 0u;
#pragma unroll
              for (int k = 0; k < 4; ++k) x[k] = __funnelshift_r(x[k], x[k + 1], hl);
#pragma unroll
              for (int k = 0; k < 4; ++k) x[k] = l8 ? 0u : x[k];
              *reinterpret_cast<uint4*>(chunk_ptr(q0 + 8)) = make_uint4(x[0], x[1], x[2], x[3]);
              for (int q = q0 + 9; q < 8 * un.t_end; ++q) *reinterpret_cast<uint4*>(chunk_ptr(q)) = make_uint4(0, 0, 0, 0);
            }
            fence_proxy_async_smem();
          }
          // (a real step got here through ld_full, which the loader arms only after the MMAs of step g-3: the same
          // bound keeps a drain step from arriving on a_done[g % 3] before the issuer has consumed its previous phase)
          if (!real && g >= 3) mbar_wait(&mma_done[g % 3], ((g - 3) / 3) & 1);
          mbar_arrive(&a_done[g % 3]);
          // dQ of the previous head leaves tensor memory once its last MMAs have retired -- one step into this head,
          // long before the issuer needs that accumulator again (head hn + 1)
          if (t == 0 && prev_h >= 0) {
            mbar_wait(&mma_done[prev_last_g % 3], (prev_last_g / 3) & 1);
            tc_fence_after();
            write_dq(prev_un, prev_h, hn - 1);
            prev_h = -1;
          }
        }
        prev_un = un;
        prev_h = h;
        prev_last_g = g - 1;
      }
    }
    if (prev_h >= 0) {
      mbar_wait(&mma_done[prev_last_g % 3], (prev_last_g / 3) & 1);
      tc_fence_after();
      write_dq(prev_un, prev_h, hn - 1);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == QB_MMA_WARP) tmem_dealloc(tmem_base, QB_TMEM_COLS);
}

// Longest-first assignment of the work units (query tile, sequence) to `n_cta` CTAs: unit (qi, bl) costs
// H * t_end(qi) steps.  Writes the packed units per CTA into `sched` ([n_cta, pitch]) and the counts into `cnt`.
static void build_schedule(int nq, int nb, int L, int noncausal, int n_cta, int pitch, std::vector<int32_t>& sched,
                           std::vector<int32_t>& cnt) {
  struct U { int cost, qi, bl; };
  std::vector<U> units;
  for (int qi = 0; qi < nq; ++qi) {
    const int i0 = qi * QB_BM;
    const int kmax = noncausal ? L : std::min(i0 + QB_BM, L);
    const int nt = (kmax + QB_BN - 1) / QB_BN;
    const int c_last = std::max(nt - 1, std::min(nt + 1, 2 * qi + 1));
    const int t_end = (c_last + 2) & ~1;
    for (int bl = 0; bl < nb; ++bl) units.push_back({t_end, qi, bl});
  }
  std::stable_sort(units.begin(), units.end(), [](const U& x, const U& y) { return x.cost > y.cost; });
  sched.assign(static_cast<size_t>(n_cta) * pitch, 0);
  cnt.assign(n_cta, 0);
  std::vector<long long> load(n_cta, 0);
  for (const U& u : units) {
    int best = 0;
    for (int c = 1; c < n_cta; ++c)
      if (load[c] < load[best]) best = c;
    sched[static_cast<size_t>(best) * pitch + cnt[best]++] = u.qi | (u.bl << 8);
    load[best] += u.cost;
  }
}

template <int DH>
static int launch_bwd_q(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tds, const CUtensorMap& te64,
                        const QbParams& p, dim3 grid, int nb, cudaStream_t st) {
  auto kern = attn_bwd_q_tc_kernel<DH>;
  static bool configured = false;
  if (!configured) {
    ME_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, QB_SMEM));
    configured = true;
  }
  cudaEvent_t pe = prof_begin(3.0 * attn_unit_flops(nb, p.H, p.L, DH), st, 3);   // dS K, T E, T^T Q
  kern<<<grid, QB_THREADS, QB_SMEM, st>>>(tq, tk, tds, te64, p);
  prof_end(pe, st);
  ME_LAUNCH_CHECK();
  return 0;
}

// dq (bf16, q strides) and the dE partial sums (into the private copies of dE_ws) for one attention call, from the
// dS tiles the key-side kernel left in `ds_scratch` ([tiles, 128, 64] bf16, described by `tds`)
// number of int32 words of the schedule region for a slice of nb sequences of length L (n_cta <= 256)
int64_t attn_bwd_q_sched_words(int L, int nb) {
  const int64_t units = static_cast<int64_t>((L + QB_BM - 1) / QB_BM) * nb;
  return 256 * units + 256;
}

// dq (bf16, q strides) and the dE partial sums (into the private copies of dE_ws) for one attention call, from the
// dS tiles the key-side kernel left in the scratch described by `tds`.  `sched_dev`: attn_bwd_q_sched_words() int32
// words of device scratch for the unit schedule.
int launch_attn_bwd_q_tc(const me_attn_bwd_args* ba, float* dE_ws, const CUtensorMap& tds, int tiles_per_head, int b0,
                         int nb, int32_t* sched_dev) {
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
  const int noncausal = (a->flags & ME_ATTN_NONCAUSAL) ? 1 : 0;
  const int nq = (L + QB_BM - 1) / QB_BM;
  const int units = nq * nb;
  int n_cta = sm_count() < 256 ? sm_count() : 256;
  if (units < n_cta) n_cta = units;
  // the unit schedule: computed on the host (cached per shape), staged in pinned memory, copied behind the stream
  {
    static std::vector<int32_t> sched, cnt;
    static int key[5] = {-1, -1, -1, -1, -1};
    static int32_t* pinned = nullptr;
    static size_t pinned_words = 0;
    const size_t words = static_cast<size_t>(n_cta) * units + n_cta;
    const int k[5] = {nq, nb, L, noncausal, n_cta};
    bool same = true;
    for (int i = 0; i < 5; ++i) same = same && key[i] == k[i];
    if (!same || pinned == nullptr || pinned_words < words) {
      build_schedule(nq, nb, L, noncausal, n_cta, units, sched, cnt);
      if (pinned_words < words) {
        if (pinned) cudaFreeHost(pinned);
        ME_CUDA(cudaMallocHost(&pinned, words * sizeof(int32_t)));
        pinned_words = words;
      } else {
        ME_CUDA(cudaStreamSynchronize(st));   // (a copy of the previous schedule may still be in flight)
      }
      memcpy(pinned, sched.data(), sched.size() * sizeof(int32_t));
      memcpy(pinned + sched.size(), cnt.data(), cnt.size() * sizeof(int32_t));
      for (int i = 0; i < 5; ++i) key[i] = k[i];
    }
    ME_CUDA(cudaMemcpyAsync(sched_dev, pinned, words * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  }
  QbParams p;
  p.B = B; p.H = H; p.L = L; p.max_seq = a->max_seq;
  p.q_sb = a->q_sb; p.q_sh = a->q_sh; p.q_si = a->q_si;
  p.dE_ws = dE_ws;
  p.dq = static_cast<bf16*>(ba->dq);
  p.noncausal = noncausal;
  p.tiles_per_head = tiles_per_head;
  p.b0 = b0;
  p.sched = sched_dev;
  p.sched_cnt = sched_dev + static_cast<size_t>(n_cta) * units;
  p.sched_pitch = units;
  dim3 grid(n_cta, 1, 1);
  if (dh == 64) return launch_bwd_q<64>(tq, tk, tds, te64, p, grid, nb, st);
  if (dh == 48) return launch_bwd_q<48>(tq, tk, tds, te64, p, grid, nb, st);
  return launch_bwd_q<32>(tq, tk, tds, te64, p, grid, nb, st);
}

}  // namespace me
