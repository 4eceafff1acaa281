// Micro-benchmark: execution time of short tcgen05.mma sequences (M = 128, K = 16 per instruction, bf16) as the
// attention kernels issue them: N = 64 / 128 / 256, K-major or MN-major operands, one or several accumulators.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/mma_rate scripts/micro/mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../midi_emotion_b200/csrc/common.cuh"
using namespace me;

// mode: amn, bmn = operand majors; N; nacc = accumulators used round-robin; ninstr per batch
__global__ void __launch_bounds__(128, 1) k_mma(int N, int amn, int bmn, int nacc, int ninstr, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  const uint32_t a_addr = smem_u32(sm), b_addr = smem_u32(sm + 65536);
  const uint32_t idesc = make_idesc_bf16(128, N, amn, bmn);
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    uint32_t ph = 0;
    for (int it = 0; it < iters + 1; ++it) {
      if (it == 1) t0 = clock64();
      if (elect_one()) {
        for (int k = 0; k < ninstr; ++k) {
          const uint64_t ad = amn ? make_smem_desc_sw128(a_addr + (k & 7) * 2048, 16384, 1024)
                                  : make_smem_desc_sw128(a_addr + (k & 3) * 32, 16, 1024);
          const uint64_t bd = bmn ? make_smem_desc_sw128(b_addr + (k & 7) * 2048, 8192, 1024)
                                  : make_smem_desc_sw128(b_addr + (k & 3) * 32, 16, 1024);
          umma_bf16(tb + (k % nacc) * N, ad, bd, idesc, 1u);
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, ph);
      ph ^= 1;
      tc_fence_after();
    }
    t1 = clock64();
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

void run(const char* name, int N, int amn, int bmn, int nacc, int ninstr) {
  long long* d;
  cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const int iters = 200;
  k_mma<<<148, 128, 160 * 1024>>>(N, amn, bmn, nacc, ninstr, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const double per_batch = double(h[0]) / iters;
  printf("%-58s %s  %7.1f cyc/batch  %6.1f cyc/instr  (math floor %d)\n", name, cudaGetErrorString(e), per_batch,
         per_batch / ninstr, N / 2);
  cudaFree(d);
}

int main() {
  run("N=256 K-major x K-major, 4 instr, 1 acc", 256, 0, 0, 1, 4);
  run("N=256 K-major x K-major, 16 instr, 1 acc", 256, 0, 0, 1, 16);
  run("N=128 K-major x K-major, 16 instr, 1 acc", 128, 0, 0, 1, 16);
  run("N=64  K-major x K-major, 4 instr, 1 acc", 64, 0, 0, 1, 4);
  run("N=64  K-major x K-major, 16 instr, 1 acc", 64, 0, 0, 1, 16);
  run("N=64  K-major x K-major, 16 instr, 2 acc", 64, 0, 0, 2, 16);
  run("N=64  K-major x K-major, 16 instr, 4 acc", 64, 0, 0, 4, 16);
  run("N=64  K-major x MN-major (dQ), 4 instr, 1 acc", 64, 0, 1, 1, 4);
  run("N=64  K-major x MN-major (dQ), 16 instr, 1 acc", 64, 0, 1, 1, 16);
  run("N=64  K-major x MN-major (dQ), 16 instr, 2 acc", 64, 0, 1, 2, 16);
  run("N=64  MN-major x MN-major (dK/dV/dE), 8 instr, 1 acc", 64, 1, 1, 1, 8);
  run("N=64  MN-major x MN-major (dK/dV/dE), 16 instr, 1 acc", 64, 1, 1, 1, 16);
  run("N=64  MN-major x MN-major (dK/dV/dE), 16 instr, 2 acc", 64, 1, 1, 2, 16);
  run("N=64  MN-major x MN-major (dK/dV/dE), 16 instr, 4 acc", 64, 1, 1, 4, 16);
  run("N=128 MN-major x MN-major, 16 instr, 1 acc", 128, 1, 1, 1, 16);
  run("N=256 MN-major x MN-major, 16 instr, 1 acc", 256, 1, 1, 1, 16);
  return 0;
}
