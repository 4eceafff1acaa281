// Output head fused with the training loss: logits = x W^T + b never leave the chip.
//   replaces  models/music_multi.py:106 (self.fc) + train.py:288-290 (CrossEntropyLoss(ignore_index = pad), mean over
//   the non-pad targets) + utils.py:15-80 (top-1 / top-5 counts), and produces the gradient w.r.t. the logits that
//   the two backward GEMMs of the head consume.
//
// One CTA owns 128 rows (positions) at a time and sweeps the vocabulary twice, 256 columns per tcgen05 tile
// (the head is 0.8 % of the step's FLOPs, so recomputing it is cheaper than a [M, V] round trip through HBM:
// 66 MB written and read back at cfg2):
//   pass 0   tile -> + bias -> rounded to bf16 (the logits the reference holds under autocast) -> running max /
//            sum of exponentials per row, and the target's logit
//   pass 1   tile again -> p = exp(x - lse) -> gradient (p - onehot) / count as bf16, written once, row pitch ld;
//            rank of the target among the row's logits (top-1 / top-5), loss = lse - x[target]
// Same pipeline as the GEMM kernel: warp 0 TMA producer (4-stage ring), warp 1 tcgen05.mma issuer with the
// accumulator double-buffered in TMEM, 8 epilogue warps (two threads per row, each half of a tile's columns).
#include "gemm_common.cuh"

namespace me {

constexpr int HC_BM = 128, HC_BN = 256, HC_BK = 64, HC_STAGES = 4;
constexpr int HC_EPI_THREADS = 256;
constexpr int HC_THREADS = 64 + HC_EPI_THREADS;
constexpr int HC_A_BYTES = HC_BM * HC_BK * 2, HC_B_BYTES = HC_BN * HC_BK * 2, HC_STAGE_BYTES = HC_A_BYTES + HC_B_BYTES;
constexpr int HC_SMEM = HC_STAGES * HC_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 2 * HC_BN * 4 /*bias*/ +
                        3 * 2 * HC_BM * 4 /*row exchange*/;

struct HeadCeParams {
  int M, V, K, ld;        // ld: row pitch of the gradient (>= V, multiple of 8)
  int num_m_tiles, num_n_tiles, num_kb;
  const float* bias;
  const int64_t* targets;
  int64_t ignore_index;
  bf16* grad;             // [M, ld] or NULL (evaluation: loss and counts only)
  float* stats;           // { sum of losses, count (already there), top-1 hits, top-5 hits }
};

__global__ void __launch_bounds__(HC_THREADS, 1)
head_ce_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, HeadCeParams p) {
  extern __shared__ uint8_t hc_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(hc_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + HC_STAGES * HC_A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + HC_STAGES * HC_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + HC_STAGES;
  uint64_t* tfull_bar = empty_bar + HC_STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4);   // [2][HC_BN]
  float* xch = bias_s + 2 * HC_BN;                           // [3][2][HC_BM]: max, sum, target logit of each half

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < HC_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], HC_EPI_THREADS);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * HC_BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_m = 2 * p.num_n_tiles;   // two passes

  if (warp == 0) {
    // ============================== TMA producer ==============================
    int s = 0;
    uint32_t phase = 0;
    for (int mt = blockIdx.x; mt < p.num_m_tiles; mt += gridDim.x) {
      for (int tn = 0; tn < tiles_per_m; ++tn) {
        const int n0 = (tn % p.num_n_tiles) * HC_BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[s], phase ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&full_bar[s], HC_STAGE_BYTES);
            tma_load_2d(&tmA, &full_bar[s], smA + s * HC_A_BYTES, kb * HC_BK, mt * HC_BM);
            tma_load_2d(&tmB, &full_bar[s], smB + s * HC_B_BYTES, kb * HC_BK, n0);
          }
          __syncwarp();
          if (++s == HC_STAGES) { s = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    constexpr uint32_t idesc = make_idesc_bf16(HC_BM, HC_BN, 0, 0);
    int s = 0, it = 0;
    uint32_t phase = 0;
    for (int mt = blockIdx.x; mt < p.num_m_tiles; mt += gridDim.x) {
      for (int tn = 0; tn < tiles_per_m; ++tn, ++it) {
        const int acc = it & 1;
        mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * HC_BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[s], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smA + s * HC_A_BYTES), b_addr = smem_u32(smB + s * HC_B_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < HC_BK / 16; ++k)
              umma_bf16(d_tmem, make_smem_desc_sw128(a_addr + k * 32, 16, 1024), make_smem_desc_sw128(b_addr + k * 32, 16, 1024),
                        idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty_bar[s]);
            if (kb == p.num_kb - 1) umma_commit(&tfull_bar[acc]);
          }
          __syncwarp();
          if (++s == HC_STAGES) { s = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ============================== epilogue ==============================
    const int quarter = warp & 3;        // TMEM lane quarter this warp may access
    const int chalf = (warp - 2) >> 2;   // which half of a tile's columns
    const int et = threadIdx.x - 64;
    const int row_in_tile = quarter * 32 + lane;
    const float inv_count = p.stats[1] > 0.f ? 1.f / p.stats[1] : 0.f;
    float loss_acc = 0.f, top1 = 0.f, top5 = 0.f;
    int it = 0;
    for (int mt = blockIdx.x; mt < p.num_m_tiles; mt += gridDim.x) {
      const int m = mt * HC_BM + row_in_tile;
      const bool row_in = m < p.M;
      const int64_t t64 = row_in ? p.targets[m] : p.ignore_index;
      const bool counted = row_in && t64 != p.ignore_index && t64 >= 0 && t64 < p.V;
      const int tt = counted ? static_cast<int>(t64) : -1;
      float mx = -INFINITY, se = 0.f, xt = 0.f, lse = 0.f;
      int bigger = 0;
      for (int tn = 0; tn < tiles_per_m; ++tn, ++it) {
        const int pass = tn / p.num_n_tiles;
        const int n0 = (tn % p.num_n_tiles) * HC_BN;
        const int acc = it & 1;
        float* bs = bias_s + acc * HC_BN;
        for (int c = et; c < HC_BN; c += HC_EPI_THREADS) bs[c] = (n0 + c < p.V) ? __ldg(p.bias + n0 + c) : 0.f;
        if (pass == 1 && tn == p.num_n_tiles) {
          // between the passes: the two threads of a row combine their halves -> lse and the target's logit
          xch[0 * 2 * HC_BM + chalf * HC_BM + row_in_tile] = mx;
          xch[1 * 2 * HC_BM + chalf * HC_BM + row_in_tile] = se;
          xch[2 * 2 * HC_BM + chalf * HC_BM + row_in_tile] = xt;
        }
        named_bar_sync(1, HC_EPI_THREADS);
        if (pass == 1 && tn == p.num_n_tiles) {
          const float mo = xch[0 * 2 * HC_BM + (chalf ^ 1) * HC_BM + row_in_tile];
          const float so = xch[1 * 2 * HC_BM + (chalf ^ 1) * HC_BM + row_in_tile];
          xt += xch[2 * 2 * HC_BM + (chalf ^ 1) * HC_BM + row_in_tile];   // only one half saw the target
          const float mm = fmaxf(mx, mo);
          const float st = se * __expf(mx - mm) + so * __expf(mo - mm);
          lse = mm + __logf(st);
        }
        mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * HC_BN;
#pragma unroll 1
        for (int c0 = chalf * (HC_BN / 2); c0 < (chalf + 1) * (HC_BN / 2); c0 += 32) {
          const int nb = n0 + c0;
          if (nb >= p.ld) break;   // warp-uniform: nothing to compute or store past the row pitch
          uint32_t r[32];
          tmem_ld32(t_row + c0, r);
          tc_wait_ld();
          float x[32];
#pragma unroll
          for (int j = 0; j < 32; ++j)   // the bf16 logit the unfused path would have stored
            x[j] = __bfloat162float(__float2bfloat16_rn(__uint_as_float(r[j]) + bs[c0 + j]));
          if (pass == 0) {
            float cm = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) cm = (nb + j < p.V) ? fmaxf(cm, x[j]) : cm;
            if (cm > mx) {
              se *= __expf(mx - cm);
              mx = cm;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) se += (nb + j < p.V) ? __expf(x[j] - mx) : 0.f;
            if (tt >= nb && tt < nb + 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j) xt = (nb + j == tt) ? x[j] : xt;
            }
          } else {
            if (counted) {
#pragma unroll
              for (int j = 0; j < 32; ++j) bigger += (nb + j < p.V) && (x[j] > xt);
            }
            if (p.grad != nullptr && row_in) {
              bf16* gp = p.grad + static_cast<int64_t>(m) * p.ld + nb;
#pragma unroll
              for (int j8 = 0; j8 < 4; ++j8) {
                if (nb + 8 * j8 >= p.ld) break;
                uint4 u;
                __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float g2[2];
#pragma unroll
                  for (int q = 0; q < 2; ++q) {
                    const int j = 8 * j8 + 2 * e + q;
                    const float pj = (counted && nb + j < p.V) ? __expf(x[j] - lse) : 0.f;
                    g2[q] = (pj - ((counted && nb + j == tt) ? 1.f : 0.f)) * inv_count;
                  }
                  h2[e] = __floats2bfloat162_rn(g2[0], g2[1]);
                }
                *reinterpret_cast<uint4*>(gp + 8 * j8) = u;
              }
            }
          }
        }
        __syncwarp();
        tc_fence_before();
        mbar_arrive(&tempty_bar[acc]);
      }
      // rank of the target: the two halves add their counts
      named_bar_sync(1, HC_EPI_THREADS);
      xch[chalf * HC_BM + row_in_tile] = static_cast<float>(bigger);
      named_bar_sync(1, HC_EPI_THREADS);
      if (chalf == 0 && counted) {
        const int total = bigger + static_cast<int>(xch[HC_BM + row_in_tile]);
        loss_acc += lse - xt;
        top1 += total < 1 ? 1.f : 0.f;
        top5 += total < 5 ? 1.f : 0.f;
      }
    }
    loss_acc = warp_sum(loss_acc);
    top1 = warp_sum(top1);
    top5 = warp_sum(top5);
    if (lane == 0 && chalf == 0) {
      if (loss_acc != 0.f) atomicAdd(&p.stats[0], loss_acc);
      if (top1 != 0.f) atomicAdd(&p.stats[2], top1);
      if (top5 != 0.f) atomicAdd(&p.stats[3], top5);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 2 * HC_BN);
  }
}

__global__ void head_ce_count_kernel(const int64_t* __restrict__ targets, int M, int V, int64_t ignore_index,
                                     float* __restrict__ stats) {
  __shared__ int part[32];
  int c = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
    const int64_t t = targets[i];
    c += (t != ignore_index && t >= 0 && t < V);
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) s += part[w];
    atomicAdd(&stats[1], static_cast<float>(s));
  }
}

}  // namespace me

using namespace me;

extern "C" int me_head_cross_entropy(const void* x, const void* W, const float* bias, int M, int V, int K, int ldx,
                                     int ldw, const int64_t* targets, int64_t ignore_index, void* grad_logits,
                                     int ld_grad, float* stats, void* stream) {
  ME_CHECK(me_device_is_sm100(), "me_head_cross_entropy: needs an sm_100 device");
  ME_CHECK(x && W && bias && targets && stats, "me_head_cross_entropy: NULL pointer");
  ME_CHECK(M > 0 && V > 0 && K > 0 && K % 8 == 0 && ldx % 8 == 0 && ldw % 8 == 0, "me_head_cross_entropy: bad sizes");
  ME_CHECK(grad_logits == nullptr || (ld_grad >= V && ld_grad % 8 == 0 && (reinterpret_cast<uintptr_t>(grad_logits) & 15) == 0),
           "me_head_cross_entropy: gradient rows must be 16-byte aligned with pitch >= V");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap ta, tb;
  if (make_tmap_2d_bf16(&ta, x, K, M, ldx, HC_BK, HC_BM)) return 1;
  if (make_tmap_2d_bf16(&tb, W, K, V, ldw, HC_BK, HC_BN)) return 1;
  HeadCeParams p;
  p.M = M; p.V = V; p.K = K; p.ld = grad_logits ? ld_grad : (V + 7) / 8 * 8;
  p.num_m_tiles = (M + HC_BM - 1) / HC_BM;
  p.num_n_tiles = (V + HC_BN - 1) / HC_BN;
  p.num_kb = (K + HC_BK - 1) / HC_BK;
  p.bias = bias; p.targets = targets; p.ignore_index = ignore_index;
  p.grad = static_cast<bf16*>(grad_logits);
  p.stats = stats;
  ME_CUDA(cudaMemsetAsync(stats, 0, 4 * sizeof(float), st));
  head_ce_count_kernel<<<min((M + 255) / 256, 64), 256, 0, st>>>(targets, M, V, ignore_index, stats);
  ME_LAUNCH_CHECK();
  static bool configured = false;
  if (!configured) {
    ME_CUDA(cudaFuncSetAttribute(head_ce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HC_SMEM));
    configured = true;
  }
  const int grid = p.num_m_tiles < sm_count() ? p.num_m_tiles : sm_count();
  cudaEvent_t pe = prof_begin(2.0 * M * V * K, st, 0);   // algorithmic: the second sweep is recomputation
  head_ce_kernel<<<grid, HC_THREADS, HC_SMEM, st>>>(ta, tb, p);
  prof_end(pe, st);
  ME_LAUNCH_CHECK();
  return 0;
}
