// Optimiser step of the reference's training loop (train.py:319-325): GradScaler.unscale_ + clip_grad_norm_ +
// Adam over every parameter tensor, as multi-tensor kernels.  HBM-bound: the norm pass reads the gradients once
// (4 B/parameter), the update pass reads p, g, m, v and writes p, m, v (28 B/parameter); nothing else moves.
// Tensor pointers travel in kernel-parameter space (64 tensors per launch), so there is no descriptor upload and
// no host synchronisation; the clip coefficient, the skip decision and the bias corrections are produced on the
// device by a one-block kernel between the two passes.
#include <limits.h>

#include "common.cuh"
#include "../../include/midi_emotion_b200.h"

namespace me {

constexpr int OPT_BATCH = 64;      // tensors per launch
constexpr int OPT_CHUNK = 8192;    // elements per thread block
constexpr int OPT_THREADS = 256;
constexpr int OPT_VPT = OPT_CHUNK / 4 / OPT_THREADS;  // 16-byte vectors per thread and array (8)

struct OptBatch {
  float* p[OPT_BATCH];
  const float* g[OPT_BATCH];
  float* m[OPT_BATCH];
  float* v[OPT_BATCH];
  long long n[OPT_BATCH];
  int chunk0[OPT_BATCH + 1];  // first block of each tensor inside this launch
  int count;
};
static_assert(sizeof(OptBatch) <= 4000, "OptBatch must fit the 4 KB kernel-parameter space");

struct AdamScalars {
  float one_minus_beta1, beta2, one_minus_beta2, eps, weight_decay, inv_scale;
};

// tensor that owns block `blk`: the last t with chunk0[t] <= blk
__device__ __forceinline__ int opt_locate(const OptBatch& b, int blk) {
  int lo = 0, hi = b.count;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (b.chunk0[mid] <= blk) lo = mid;
    else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(OPT_THREADS)
grad_sqnorm_kernel(const __grid_constant__ OptBatch b, float inv_scale, float* __restrict__ partials) {
  __shared__ float part[OPT_THREADS / 32];
  const int t = opt_locate(b, blockIdx.x);
  const long long off = static_cast<long long>(blockIdx.x - b.chunk0[t]) * OPT_CHUNK;
  const long long rem = b.n[t] - off;
  const int cnt = rem < OPT_CHUNK ? static_cast<int>(rem) : OPT_CHUNK;
  const float* g = b.g[t] + off;
  float acc = 0.f;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0 && cnt >= 4) {
    const int nv = cnt >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4 x[OPT_VPT];
#pragma unroll
    for (int k = 0; k < OPT_VPT; ++k) {   // all loads first (clamped index), then the arithmetic
      const int i = threadIdx.x + k * OPT_THREADS;
      x[k] = __ldg(g4 + (i < nv ? i : 0));
    }
#pragma unroll
    for (int k = 0; k < OPT_VPT; ++k) {
      const int i = threadIdx.x + k * OPT_THREADS;
      if (i < nv) {
        const float a = x[k].x * inv_scale, c = x[k].y * inv_scale, d = x[k].z * inv_scale, e = x[k].w * inv_scale;
        acc += a * a + c * c + d * d + e * e;
      }
    }
    for (int i = (nv << 2) + threadIdx.x; i < cnt; i += OPT_THREADS) {
      const float a = g[i] * inv_scale;
      acc += a * a;
    }
  } else {
    for (int i = threadIdx.x; i < cnt; i += OPT_THREADS) {
      const float a = g[i] * inv_scale;
      acc += a * a;
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < OPT_THREADS / 32; ++w) s += part[w];
    partials[blockIdx.x] = s;
  }
}

// One block: fixed-order (deterministic) fp64 sum of the partials, then every scalar the update pass needs.
__global__ void __launch_bounds__(1024)
adam_prepare_kernel(const float* __restrict__ partials, long long n_partials, float max_norm, double lr, double beta1,
                    double beta2, float* __restrict__ step_dev, float* __restrict__ stats) {
  __shared__ double red[32];
  double acc = 0.0;
  for (long long i = threadIdx.x; i < n_partials; i += 1024) acc += static_cast<double>(partials[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 32; ++w) s += red[w];
    const float total = n_partials > 0 ? static_cast<float>(sqrt(s)) : 0.f;
    const bool finite = isfinite(total);
    float coef = 1.f;
    if (max_norm > 0.f && n_partials > 0) coef = fminf(max_norm / (total + 1e-6f), 1.f);  // clip_grad_norm_
    float step = *step_dev;
    if (finite) {
      step += 1.f;
      *step_dev = step;
    }
    const double bc1 = 1.0 - pow(beta1, static_cast<double>(step));
    const double bc2 = 1.0 - pow(beta2, static_cast<double>(step));
    stats[0] = total;
    stats[1] = finite ? coef : 0.f;
    stats[2] = finite ? 0.f : 1.f;
    stats[3] = step;
    stats[4] = static_cast<float>(lr / bc1);
    stats[5] = static_cast<float>(sqrt(bc2));
    stats[6] = 0.f;
    stats[7] = 0.f;
  }
}

__device__ __forceinline__ void adam_element(float& p, float g, float& m, float& v, const AdamScalars& s, float coef,
                                             float step_size, float bc2_sqrt) {
  g = (g * s.inv_scale) * coef;                       // unscale_ then clip, the reference's two roundings
  if (s.weight_decay != 0.f) g = g + s.weight_decay * p;
  m = m + s.one_minus_beta1 * (g - m);                // exp_avg.lerp_(grad, 1 - beta1)
  v = s.beta2 * v + s.one_minus_beta2 * g * g;        // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) / bc2_sqrt + s.eps;
  p = p - step_size * (m / denom);                    // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(OPT_THREADS)
adam_update_kernel(const __grid_constant__ OptBatch b, const AdamScalars s, const float* __restrict__ stats) {
  if (stats[2] != 0.f) return;   // non-finite gradient norm: the step is skipped (GradScaler.step)
  const float coef = stats[1], step_size = stats[4], bc2_sqrt = stats[5];
  const int t = opt_locate(b, blockIdx.x);
  const long long off = static_cast<long long>(blockIdx.x - b.chunk0[t]) * OPT_CHUNK;
  const long long rem = b.n[t] - off;
  const int cnt = rem < OPT_CHUNK ? static_cast<int>(rem) : OPT_CHUNK;
  float* p = b.p[t] + off;
  const float* g = b.g[t] + off;
  float* m = b.m[t] + off;
  float* v = b.v[t] + off;
  const uintptr_t bits = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                         reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v);
  int done = 0;
  if ((bits & 15) == 0) {
    const int nv = cnt >> 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    constexpr int ROUND = 4;   // vectors per array in flight per thread: 16 independent 16-byte loads
#pragma unroll 1
    for (int base = 0; base < nv; base += ROUND * OPT_THREADS) {
      float4 P[ROUND], G[ROUND], M[ROUND], V[ROUND];
#pragma unroll
      for (int k = 0; k < ROUND; ++k) {
        const int i = base + threadIdx.x + k * OPT_THREADS;
        const int ic = i < nv ? i : nv - 1;
        G[k] = __ldg(g4 + ic);
        P[k] = p4[ic];
        M[k] = m4[ic];
        V[k] = v4[ic];
      }
#pragma unroll
      for (int k = 0; k < ROUND; ++k) {
        const int i = base + threadIdx.x + k * OPT_THREADS;
        if (i < nv) {
          adam_element(P[k].x, G[k].x, M[k].x, V[k].x, s, coef, step_size, bc2_sqrt);
          adam_element(P[k].y, G[k].y, M[k].y, V[k].y, s, coef, step_size, bc2_sqrt);
          adam_element(P[k].z, G[k].z, M[k].z, V[k].z, s, coef, step_size, bc2_sqrt);
          adam_element(P[k].w, G[k].w, M[k].w, V[k].w, s, coef, step_size, bc2_sqrt);
          p4[i] = P[k];
          m4[i] = M[k];
          v4[i] = V[k];
        }
      }
    }
    done = nv << 2;
  }
  for (int i = done + threadIdx.x; i < cnt; i += OPT_THREADS) {
    float pp = p[i], mm = m[i], vv = v[i];
    adam_element(pp, g[i], mm, vv, s, coef, step_size, bc2_sqrt);
    p[i] = pp;
    m[i] = mm;
    v[i] = vv;
  }
}

static int check_tensors(const me_adam_tensor* t, int n, bool need_state, const char* who) {
  ME_CHECK(t != nullptr && n > 0, "%s: empty tensor list", who);
  for (int i = 0; i < n; ++i) {
    ME_CHECK(t[i].numel >= 0, "%s: tensor %d has a negative size", who, i);
    if (t[i].numel == 0) continue;
    ME_CHECK(t[i].grad != nullptr && (reinterpret_cast<uintptr_t>(t[i].grad) & 3) == 0, "%s: tensor %d: bad gradient pointer",
             who, i);
    if (need_state)
      ME_CHECK(t[i].param && t[i].exp_avg && t[i].exp_avg_sq &&
                   ((reinterpret_cast<uintptr_t>(t[i].param) | reinterpret_cast<uintptr_t>(t[i].exp_avg) |
                     reinterpret_cast<uintptr_t>(t[i].exp_avg_sq)) & 3) == 0,
               "%s: tensor %d: bad parameter / state pointer", who, i);
  }
  return 0;
}

// Cuts the tensor list into launches of at most OPT_BATCH tensors; launch(batch, blocks, first_block_overall).
template <typename F>
static int for_each_batch(const me_adam_tensor* t, int n, F&& launch) {
  int i = 0;
  long long first = 0;
  while (i < n) {
    OptBatch b;
    b.count = 0;
    long long blocks = 0;
    while (i < n && b.count < OPT_BATCH) {
      const long long chunks = (t[i].numel + OPT_CHUNK - 1) / OPT_CHUNK;
      if (chunks == 0) { ++i; continue; }
      if (blocks + chunks > INT_MAX / 2) {
        if (b.count == 0) { set_error("optimizer: tensor %d is too large for one launch", i); return 1; }
        break;
      }
      b.p[b.count] = t[i].param;
      b.g[b.count] = t[i].grad;
      b.m[b.count] = t[i].exp_avg;
      b.v[b.count] = t[i].exp_avg_sq;
      b.n[b.count] = t[i].numel;
      b.chunk0[b.count] = static_cast<int>(blocks);
      blocks += chunks;
      ++b.count;
      ++i;
    }
    if (b.count == 0) break;
    for (int k = b.count; k <= OPT_BATCH; ++k) b.chunk0[k] = static_cast<int>(blocks);
    for (int k = b.count; k < OPT_BATCH; ++k) { b.p[k] = nullptr; b.g[k] = nullptr; b.m[k] = nullptr; b.v[k] = nullptr; b.n[k] = 0; }
    if (launch(b, static_cast<int>(blocks), first)) return 1;
    first += blocks;
  }
  return 0;
}

}  // namespace me

using namespace me;

extern "C" int64_t me_grad_sqnorm_chunks(const me_adam_tensor* tensors, int n) {
  if (tensors == nullptr || n <= 0) return -1;
  int64_t total = 0;
  for (int i = 0; i < n; ++i) {
    if (tensors[i].numel < 0) return -1;
    total += (tensors[i].numel + OPT_CHUNK - 1) / OPT_CHUNK;
  }
  return total;
}

extern "C" int me_grad_sqnorm_partials(const me_adam_tensor* tensors, int n, double grad_scale, float* partials,
                                       int64_t partials_capacity, void* stream) {
  if (check_tensors(tensors, n, false, "me_grad_sqnorm_partials")) return 1;
  ME_CHECK(grad_scale > 0.0, "me_grad_sqnorm_partials: grad_scale must be positive");
  const int64_t need = me_grad_sqnorm_chunks(tensors, n);
  ME_CHECK(partials != nullptr && partials_capacity >= need, "me_grad_sqnorm_partials: %lld partial sums needed, room for %lld",
           static_cast<long long>(need), static_cast<long long>(partials_capacity));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float inv_scale = static_cast<float>(1.0 / grad_scale);
  return for_each_batch(tensors, n, [&](const OptBatch& b, int blocks, long long first) -> int {
    grad_sqnorm_kernel<<<blocks, OPT_THREADS, 0, st>>>(b, inv_scale, partials + first);
    ME_LAUNCH_CHECK();
    return 0;
  });
}

extern "C" int me_adam_prepare(const float* partials, int64_t n_partials, double max_grad_norm, double lr, double beta1,
                               double beta2, float* step_dev, float* stats, void* stream) {
  ME_CHECK(n_partials >= 0 && (n_partials == 0 || partials != nullptr), "me_adam_prepare: bad partial sums");
  ME_CHECK(step_dev != nullptr && stats != nullptr, "me_adam_prepare: step / stats pointers are required");
  ME_CHECK(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0, "me_adam_prepare: betas must lie in [0, 1)");
  adam_prepare_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      partials, static_cast<long long>(n_partials), static_cast<float>(max_grad_norm > 0.0 ? max_grad_norm : 0.0), lr, beta1,
      beta2, step_dev, stats);
  ME_LAUNCH_CHECK();
  return 0;
}

extern "C" int me_adam_update(const me_adam_tensor* tensors, int n, double beta1, double beta2, double eps,
                              double weight_decay, double grad_scale, const float* stats, void* stream) {
  if (check_tensors(tensors, n, true, "me_adam_update")) return 1;
  ME_CHECK(stats != nullptr, "me_adam_update: stats (written by me_adam_prepare) is required");
  ME_CHECK(grad_scale > 0.0, "me_adam_update: grad_scale must be positive");
  AdamScalars s;
  s.one_minus_beta1 = static_cast<float>(1.0 - beta1);
  s.beta2 = static_cast<float>(beta2);
  s.one_minus_beta2 = static_cast<float>(1.0 - beta2);
  s.eps = static_cast<float>(eps);
  s.weight_decay = static_cast<float>(weight_decay);
  s.inv_scale = static_cast<float>(1.0 / grad_scale);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return for_each_batch(tensors, n, [&](const OptBatch& b, int blocks, long long) -> int {
    adam_update_kernel<<<blocks, OPT_THREADS, 0, st>>>(b, s, stats);
    ME_LAUNCH_CHECK();
    return 0;
  });
}
