// Micro-benchmark: per-SM throughput of the ways a tile of fp32 partial sums can leave an SM.
//   mode 0  cp.reduce.async.bulk add.f32, every CTA its own global region (dQ-like)
//   mode 1  cp.reduce.async.bulk add.f32, all CTAs into one shared 512 KB region (dE-like)
//   mode 2  cp.async.bulk plain store, private region (upper bound: no read-modify-write)
//   mode 3  cp.reduce.async.bulk add.noftz.bf16, private region
//   mode 4  red.global.add.v4.f32 issued by 256 threads, private region
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/reduce_bw scripts/micro/reduce_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../midi_emotion_b200/csrc/common.cuh"
using namespace me;

constexpr int CH = 32768;            // bytes per bulk operation
constexpr int REGION = 1 << 20;      // private bytes per CTA

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_reduce_add_bf16(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.noftz.bf16 [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) k_red(int iters, char* g, long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t sm[];
  for (int i = threadIdx.x; i < 2 * CH / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 1.0f;
  fence_proxy_async_smem();
  __syncthreads();
  const long long t0 = clock64();
  if (MODE == 4) {
    float4* dst = reinterpret_cast<float4*>(g + static_cast<size_t>(blockIdx.x) * REGION);
    for (int it = 0; it < iters; ++it) {
      const int off = (it % (REGION / CH)) * (CH / 16);
#pragma unroll
      for (int k = 0; k < CH / 16 / 256; ++k) {
        float4* p = dst + off + k * 256 + threadIdx.x;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p), "f"(1.0f) : "memory");
      }
    }
    __syncthreads();
  } else if (threadIdx.x == 0) {
    for (int it = 0; it < iters; ++it) {
      char* dst = (MODE == 1) ? g + static_cast<size_t>((blockIdx.x * 5 + it) % 16) * CH
                              : g + static_cast<size_t>(blockIdx.x) * REGION + static_cast<size_t>(it % (REGION / CH)) * CH;
      const void* src = sm + (it & 1) * CH;
      if (MODE == 2) bulk_store(dst, src, CH);
      else if (MODE == 3) bulk_reduce_add_bf16(dst, src, CH);
      else bulk_reduce_add_f32(reinterpret_cast<float*>(dst), src, CH);
      bulk_commit();
      bulk_wait_read_1();
    }
    bulk_wait_all();
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int grid) {
  char* g;
  long long* d;
  cudaMalloc(&g, static_cast<size_t>(grid) * REGION);
  cudaMemset(g, 0, static_cast<size_t>(grid) * REGION);
  cudaMalloc(&d, grid * 8);
  cudaFuncSetAttribute(k_red<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * CH);
  const int iters = 400;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_red<MODE><<<grid, 256, 2 * CH>>>(iters, g, d);
  cudaEventRecord(e0);
  k_red<MODE><<<grid, 256, 2 * CH>>>(iters, g, d);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[296];
  cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  const double bytes = double(iters) * CH;
  printf("%-44s grid %3d %s  %.1f B/cycle/SM   %.2f TB/s total\n", name, grid, cudaGetErrorString(e), bytes / double(mx),
         bytes * grid / (ms * 1e-3) / 1e12);
  cudaFree(g);
  cudaFree(d);
}

int main() {
  run<0>("bulk reduce add.f32, private regions", 148);
  run<0>("bulk reduce add.f32, private regions", 16);
  run<1>("bulk reduce add.f32, one shared 512 KB region", 148);
  run<2>("bulk store, private regions", 148);
  run<3>("bulk reduce add.bf16, private regions", 148);
  run<4>("red.global.add.v4.f32 by threads, private", 148);
  return 0;
}
