// Micro-benchmark: tcgen05.ld (TMEM -> registers) throughput per SM for 1/4/8 warps, shapes x32/x64.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/tmem_bw scripts/micro/tmem_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../midi_emotion_b200/csrc/common.cuh"
using namespace me;

template <int NW, int X>
__global__ void __launch_bounds__(NW * 32, 1) k_ld(int iters, long long* out_cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (X == 128) {  // two x64 loads in flight before the wait
      uint32_t r[64], q[64];
      tmem_ld64(base + (it & 1) * 128, r);
      tmem_ld64(base + (it & 1) * 128 + 64, q);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 64; i += 16) acc ^= r[i] ^ q[i];
    } else if (X == 64) {
      uint32_t r[64];
      tmem_ld64(base + (it & 3) * 64, r);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 64; i += 16) acc ^= r[i];
    } else {
      uint32_t r[32];
      tmem_ld32(base + (it & 7) * 32, r);
      tc_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; i += 16) acc ^= r[i];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

template <int NW, int X>
void run(const char* name) {
  long long* d;
  uint32_t* sink;
  cudaMalloc(&d, 148 * 8);
  cudaMalloc(&sink, 148 * 1024 * 4);
  const int iters = 2000;
  k_ld<NW, X><<<148, NW * 32>>>(iters, d, sink);
  k_ld<NW, X><<<148, NW * 32>>>(iters, d, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const double bytes = double(iters) * NW * 32 * X * 4;
  printf("%-28s %s cycles %lld  -> %.1f B/cycle/SM\n", name, cudaGetErrorString(e), h[0], bytes / double(h[0]));
  cudaFree(d);
  cudaFree(sink);
}

int main() {
  run<1, 64>("1 warp  x64");
  run<4, 64>("4 warps x64 (1 per quarter)");
  run<8, 64>("8 warps x64 (2 per quarter)");
  run<1, 32>("1 warp  x32");
  run<4, 32>("4 warps x32");
  run<8, 32>("8 warps x32");
  run<4, 128>("4 warps 2 x x64 in flight");
  run<8, 128>("8 warps 2 x x64 in flight");
  run<16, 64>("16 warps x64 (4 per quarter)");
  return 0;
}
