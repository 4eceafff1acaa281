// Exact-order fp32 SIMT GEMM: D[M,N] = A . B^T (+ epilogue).  This is the fp32 parity path
// (no TF32, fp32 FMA accumulation in ascending k) and the on-device reference the tcgen05 GEMM
// is unit-tested against.  Operand majorness is handled with element strides.
#include "common.cuh"
#include "../../include/midi_emotion_b200.h"

namespace me {

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16, SG_TM = 4, SG_TN = 4;

template <typename TA, typename TB>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const TA* __restrict__ A, const TB* __restrict__ B, float* __restrict__ D, int M, int N, int K,
                 int64_t a_sm, int64_t a_sk, int64_t b_sn, int64_t b_sk, int ldd, int flags,
                 const float* __restrict__ bias, const float* __restrict__ addend,
                 const float* __restrict__ relu_mask, int ldmask) {
  __shared__ float As[SG_BK][SG_BM + 1];
  __shared__ float Bs[SG_BK][SG_BN + 1];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  float acc[SG_TM][SG_TN];
#pragma unroll
  for (int i = 0; i < SG_TM; ++i)
#pragma unroll
    for (int j = 0; j < SG_TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += SG_BK) {
    // cooperative loads: 64x16 elements each, 256 threads -> 4 per thread
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = threadIdx.x + i * 256;
      int r, kk;
      if (a_sk == 1) { r = e / SG_BK; kk = e % SG_BK; } else { r = e % SG_BM; kk = e / SG_BM; }
      const int m = m0 + r, k = k0 + kk;
      As[kk][r] = (m < M && k < K) ? to_f32<TA>(A[m * a_sm + k * a_sk]) : 0.f;
      if (b_sk == 1) { r = e / SG_BK; kk = e % SG_BK; } else { r = e % SG_BN; kk = e / SG_BN; }
      const int n = n0 + r;
      Bs[kk][r] = (n < N && k0 + kk < K) ? to_f32<TB>(B[n * b_sn + (k0 + kk) * b_sk]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      float a[SG_TM], b[SG_TN];
#pragma unroll
      for (int i = 0; i < SG_TM; ++i) a[i] = As[kk][ty * SG_TM + i];
#pragma unroll
      for (int j = 0; j < SG_TN; ++j) b[j] = Bs[kk][tx * SG_TN + j];
#pragma unroll
      for (int i = 0; i < SG_TM; ++i)
#pragma unroll
        for (int j = 0; j < SG_TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < SG_TM; ++i) {
    const int m = m0 + ty * SG_TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < SG_TN; ++j) {
      const int n = n0 + tx * SG_TN + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (flags & ME_EPI_BIAS) v += bias[n];
      if (flags & ME_EPI_ADD_F32) v += addend[static_cast<int64_t>(m) * ldd + n];
      if (flags & ME_EPI_RELU) v = fmaxf(v, 0.f);
      if (flags & ME_EPI_RELU_MASK) v = relu_mask[static_cast<int64_t>(m) * ldmask + n] > 0.f ? v : 0.f;
      D[static_cast<int64_t>(m) * ldd + n] = v;
    }
  }
}

int launch_gemm_f32(const float* A, const float* B, float* D, int M, int N, int K, int lda, int ldb, int ldd,
                    int a_mn, int b_mn, int flags, const float* bias, const float* addend,
                    const float* relu_mask, int ldmask, cudaStream_t st) {
  dim3 grid((N + SG_BN - 1) / SG_BN, (M + SG_BM - 1) / SG_BM);
  const int64_t a_sm = a_mn ? 1 : lda, a_sk = a_mn ? lda : 1;
  const int64_t b_sn = b_mn ? 1 : ldb, b_sk = b_mn ? ldb : 1;
  gemm_simt_kernel<float, float><<<grid, 256, 0, st>>>(A, B, D, M, N, K, a_sm, a_sk, b_sn, b_sk, ldd, flags, bias,
                                                       addend, relu_mask, ldmask);
  ME_LAUNCH_CHECK();
  return 0;
}

// bf16 operands, fp32 accumulate/output: reference implementation for the tcgen05 unit tests.
int launch_gemm_bf16_ref(const bf16* A, const bf16* B, float* D, int M, int N, int K, int lda, int ldb, int ldd,
                         int a_mn, int b_mn, cudaStream_t st) {
  dim3 grid((N + SG_BN - 1) / SG_BN, (M + SG_BM - 1) / SG_BM);
  const int64_t a_sm = a_mn ? 1 : lda, a_sk = a_mn ? lda : 1;
  const int64_t b_sn = b_mn ? 1 : ldb, b_sk = b_mn ? ldb : 1;
  gemm_simt_kernel<bf16, bf16><<<grid, 256, 0, st>>>(A, B, D, M, N, K, a_sm, a_sk, b_sn, b_sk, ldd, 0, nullptr,
                                                     nullptr, nullptr, 0);
  ME_LAUNCH_CHECK();
  return 0;
}

}  // namespace me

extern "C" int me_gemm_f32(const float* A, const float* B, float* D, int M, int N, int K, int lda, int ldb,
                           int ldd, int a_mn, int b_mn, int epi_flags, const float* bias, const float* addend,
                           const float* relu_mask, int ldmask, void* stream) {
  using namespace me;
  ME_CHECK(M > 0 && N > 0 && K > 0, "me_gemm_f32: bad dims %d %d %d", M, N, K);
  ME_CHECK(!(epi_flags & ME_EPI_BIAS) || bias, "me_gemm_f32: bias flag without pointer");
  ME_CHECK(!(epi_flags & ME_EPI_ADD_F32) || addend, "me_gemm_f32: addend flag without pointer");
  ME_CHECK(!(epi_flags & ME_EPI_RELU_MASK) || relu_mask, "me_gemm_f32: mask flag without pointer");
  return launch_gemm_f32(A, B, D, M, N, K, lda, ldb, ldd, a_mn, b_mn, epi_flags, bias, addend, relu_mask, ldmask,
                         static_cast<cudaStream_t>(stream));
}

// test hook: fp32 = bf16 x bf16 on CUDA cores
extern "C" int me_gemm_bf16_reference(const void* A, const void* B, float* D, int M, int N, int K, int lda,
                                      int ldb, int ldd, int a_mn, int b_mn, void* stream) {
  using namespace me;
  return launch_gemm_bf16_ref(static_cast<const bf16*>(A), static_cast<const bf16*>(B), D, M, N, K, lda, ldb, ldd,
                              a_mn, b_mn, static_cast<cudaStream_t>(stream));
}
