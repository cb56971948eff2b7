// fp32 FMA GEMM for the 1e-5 parity path: C[M,N] = op(A) op(B) + bias + beta*C.
// Serves every nn.Linear call site and the x*W_ih^T projections hoisted out of
// nn.LSTM / nn.GRU (see include/slnlp_b200.h).  64x64x16 tiles, 256 threads,
// 4x4 register micro-tiles, register-staged double buffering.  Deterministic
// (no split-K atomics): the summation order depends only on the shape.
#include "common.cuh"

namespace slnlp {

constexpr int GM = 64, GN = 64, GK = 16;

// A(m,k) = A[m*a_rs + k*a_cs]; B(k,n) = B[k*b_rs + n*b_cs]
template <bool A_KCONTIG, bool B_NCONTIG>
__global__ void __launch_bounds__(256) gemm_f32_kernel(int M, int N, int K,
                                                       const float* __restrict__ A, int64_t a_rs, int64_t a_cs,
                                                       const float* __restrict__ B, int64_t b_rs, int64_t b_cs,
                                                       float* __restrict__ C, int ldc,
                                                       const float* __restrict__ bias, float beta) {
  __shared__ __align__(16) float As[2][GK][GM + 4];
  __shared__ __align__(16) float Bs[2][GK][GN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads; thread -> rows ty*4.., cols tx*4..

  // global->register staging: 64x16 = 1024 elements per operand, 4 per thread.
  // the thread->element map keeps the contiguous global dimension fastest.
  int a_m[4], a_k[4], b_k[4], b_n[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int e = tid + i * 256;
    if (A_KCONTIG) { a_k[i] = e & 15; a_m[i] = e >> 4; } else { a_m[i] = e & 63; a_k[i] = e >> 6; }
    if (B_NCONTIG) { b_n[i] = e & 63; b_k[i] = e >> 6; } else { b_k[i] = e & 15; b_n[i] = e >> 4; }
  }
  float ra[4], rb[4];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + a_m[i], k = k0 + a_k[i];
      ra[i] = (m < M && k < K) ? __ldg(A + (int64_t)m * a_rs + (int64_t)k * a_cs) : 0.f;
      const int n = n0 + b_n[i], kb = k0 + b_k[i];
      rb[i] = (n < N && kb < K) ? __ldg(B + (int64_t)kb * b_rs + (int64_t)n * b_cs) : 0.f;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[buf][a_k[i]][a_m[i]] = ra[i];
      Bs[buf][b_k[i]][b_n[i]] = rb[i];
    }
  };

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = (K + GK - 1) / GK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tile((kt + 1) * GK);
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      float* c = C + (int64_t)m * ldc + n;
      if (beta != 0.f) v += beta * *c;
      *c = v;
    }
  }
}

}  // namespace slnlp

using namespace slnlp;

extern "C" int slnlp_gemm_f32(int transA, int transB, int M, int N, int K, const float* A, int lda,
                              const float* B, int ldb, float* C, int ldc, const float* bias, float beta,
                              slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(A && B && C, "gemm_f32: null pointer");
  SLNLP_CHECK_ARG(M >= 0 && N >= 0 && K >= 0 && ldc >= N, "gemm_f32: bad shape M=%d N=%d K=%d ldc=%d", M, N, K, ldc);
  SLNLP_CHECK_ARG(lda >= (transA ? M : K) && ldb >= (transB ? K : N), "gemm_f32: bad lda/ldb");
  if (M == 0 || N == 0) return 0;
  dim3 grid(ceil_div(N, GN), ceil_div(M, GM));
  SLNLP_CHECK_ARG(grid.y <= 65535, "gemm_f32: M too large");
  const int64_t a_rs = transA ? 1 : lda, a_cs = transA ? lda : 1;
  const int64_t b_rs = transB ? 1 : ldb, b_cs = transB ? ldb : 1;
  cudaStream_t s = as_stream(stream);
  if (!transA && !transB)
    gemm_f32_kernel<true, true><<<grid, 256, 0, s>>>(M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, bias, beta);
  else if (!transA && transB)
    gemm_f32_kernel<true, false><<<grid, 256, 0, s>>>(M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, bias, beta);
  else if (transA && !transB)
    gemm_f32_kernel<false, true><<<grid, 256, 0, s>>>(M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, bias, beta);
  else
    gemm_f32_kernel<false, false><<<grid, 256, 0, s>>>(M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, bias, beta);
  SLNLP_LAUNCH_OK("gemm_f32");
  return 0;
}
