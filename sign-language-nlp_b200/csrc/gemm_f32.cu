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
  pdl_wait();
  pdl_launch_dependents();
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

  // read-modify-write epilogue: issue every load of the old C before the first store so the
  // loads overlap instead of serialising behind possibly-aliasing stores
  if (beta != 0.f) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        if (m < M && n < N) acc[i][j] += beta * C[(int64_t)m * ldc + n];
      }
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
      C[(int64_t)m * ldc + n] = v;
    }
  }
}


// Vectorised variant: 128x64x16 tile, 8x4 register micro-tile, float4 global loads along
// each operand's contiguous dimension.  Requires 16-byte aligned operands and the
// contiguous extent (K for k-contiguous, M / N otherwise) and leading dims % 4 == 0.
constexpr int VM = 128, VN = 64, VK = 16;
template <bool A_KCONTIG, bool B_NCONTIG>
__global__ void __launch_bounds__(256) gemm_f32_vec_kernel(int M, int N, int Kfull, int kchunk,
                                                           const float* __restrict__ A, int64_t lda,
                                                           const float* __restrict__ B, int64_t ldb,
                                                           float* __restrict__ C, int ldc,
                                                           const float* __restrict__ bias, float beta,
                                                           float* __restrict__ partial) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ __align__(16) float As[2][VK][VM + 4];
  __shared__ __align__(16) float Bs[2][VK][VN + 4];
  // split-K: this CTA reduces k in [kbeg, K); partial sums go to partial[z][M][N]
  const int kbeg = blockIdx.z * kchunk;
  const int K = min(Kfull, kbeg + kchunk);
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * VM, n0 = blockIdx.x * VN;
  const int tx = tid & 15, ty = tid >> 4;  // rows ty*8.., cols tx*4..
  float4 ra[2], rb;
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int e = tid + i * 256;  // 512 float4 in the A tile
      if (A_KCONTIG) {              // A[m][k]: 4 float4 per row
        const int m = m0 + (e >> 2), k = k0 + (e & 3) * 4;
        ra[i] = (m < M && k < K) ? __ldg(reinterpret_cast<const float4*>(A + (int64_t)m * lda + k))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {                      // A[k][m]: 32 float4 per k row
        const int k = k0 + (e >> 5), m = m0 + (e & 31) * 4;
        ra[i] = (m < M && k < K) ? __ldg(reinterpret_cast<const float4*>(A + (int64_t)k * lda + m))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    if (B_NCONTIG) {                // B[k][n]: 16 float4 per k row
      const int k = k0 + (tid >> 4), n = n0 + (tid & 15) * 4;
      rb = (n < N && k < K) ? __ldg(reinterpret_cast<const float4*>(B + (int64_t)k * ldb + n))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {                        // B[n][k]: 4 float4 per n row
      const int n = n0 + (tid >> 2), k = k0 + (tid & 3) * 4;
      rb = (n < N && k < K) ? __ldg(reinterpret_cast<const float4*>(B + (int64_t)n * ldb + k))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int e = tid + i * 256;
      if (A_KCONTIG) {
        const int m = e >> 2, k = (e & 3) * 4;
        As[buf][k][m] = ra[i].x; As[buf][k + 1][m] = ra[i].y; As[buf][k + 2][m] = ra[i].z; As[buf][k + 3][m] = ra[i].w;
      } else {
        *reinterpret_cast<float4*>(&As[buf][e >> 5][(e & 31) * 4]) = ra[i];
      }
    }
    if (B_NCONTIG) {
      *reinterpret_cast<float4*>(&Bs[buf][tid >> 4][(tid & 15) * 4]) = rb;
    } else {
      const int n = tid >> 2, k = (tid & 3) * 4;
      Bs[buf][k][n] = rb.x; Bs[buf][k + 1][n] = rb.y; Bs[buf][k + 2][n] = rb.z; Bs[buf][k + 3][n] = rb.w;
    }
  };
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int nk = (K - kbeg + VK - 1) / VK;
  load_tile(kbeg);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tile(kbeg + (kt + 1) * VK);
#pragma unroll
    for (int k = 0; k < VK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }
  if (partial) {
    float* P = partial + (int64_t)blockIdx.z * M * N;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = m0 + ty * 8 + i;
      if (m >= M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        if (n < N) P[(int64_t)m * N + n] = acc[i][j];
      }
    }
    return;
  }
  if (beta != 0.f) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = m0 + ty * 8 + i;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        if (m < M && n < N) acc[i][j] += beta * C[(int64_t)m * ldc + n];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      C[(int64_t)m * ldc + n] = v;
    }
  }
}

// C = sum_z partial[z] + bias + beta*C, fixed summation order (deterministic split-K)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ partial, int splits,
                                                            int M, int N, float* __restrict__ C, int ldc,
                                                            const float* __restrict__ bias, float beta) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t total = (int64_t)M * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    float v = 0.f;
    for (int z = 0; z < splits; ++z) v += partial[(int64_t)z * total + i];
    const int m = (int)(i / N), n = (int)(i % N);
    if (bias) v += bias[n];
    float* c = C + (int64_t)m * ldc + n;
    if (beta != 0.f) v += beta * *c;
    *c = v;
  }
}

// host-side launcher shared with gemm_tma.cu
void launch_splitk_reduce(const float* partial, int splits, int M, int N, float* C, int ldc, const float* bias,
                          float beta, cudaStream_t s) {
  const int64_t total = (int64_t)M * N;
  int rg = (int)((total + 255) / 256);
  if (rg > sm_count() * 4) rg = sm_count() * 4;
  launch_pdl(splitk_reduce_kernel, dim3(rg), dim3(256), 0, s, partial, splits, M, N, C, ldc, bias, beta);
  note_launches(1);
}

}  // namespace slnlp

using namespace slnlp;

extern "C" int64_t slnlp_gemm_workspace_floats(void) { return (int64_t)2 * 160 * 128 * 64; }

extern "C" int slnlp_gemm_f32(int transA, int transB, int M, int N, int K, const float* A, int lda,
                              const float* B, int ldb, float* C, int ldc, const float* bias, float beta,
                              float* workspace, int64_t workspace_floats, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(A && B && C, "gemm_f32: null pointer");
  SLNLP_CHECK_ARG(M >= 0 && N >= 0 && K >= 0 && ldc >= N, "gemm_f32: bad shape M=%d N=%d K=%d ldc=%d", M, N, K, ldc);
  SLNLP_CHECK_ARG(lda >= (transA ? M : K) && ldb >= (transB ? K : N), "gemm_f32: bad lda/ldb");
  if (M == 0 || N == 0) return 0;
  dim3 grid(ceil_div(N, GN), ceil_div(M, GM));
  SLNLP_CHECK_ARG(grid.y <= 65535, "gemm_f32: M too large");
  const int64_t a_rs = transA ? 1 : lda, a_cs = transA ? lda : 1;
  const int64_t b_rs = transB ? 1 : ldb, b_cs = transB ? ldb : 1;
  cudaStream_t s = as_stream(stream);
  // vector path: contiguous extents and leading dims multiples of 4, 16-byte aligned bases
  const bool a_ok = ((uintptr_t)A % 16 == 0) && lda % 4 == 0 && ((transA ? M : K) % 4 == 0);
  const bool b_ok = ((uintptr_t)B % 16 == 0) && ldb % 4 == 0 && ((transB ? K : N) % 4 == 0);
  if (a_ok && b_ok && M >= 64) {
    dim3 vgrid(ceil_div(N, VN), ceil_div(M, VM));
    // split-K when the output has too few tiles to fill the GPU and K is long (the dW GEMMs):
    // tiles*splits ~ 2 CTAs per SM, each split >= 128 deep, partials in the caller's workspace
    const int tiles = vgrid.x * vgrid.y;
    int splits = 1;
    if (workspace && tiles < sm_count() && K >= 512) {
      splits = (2 * sm_count()) / tiles;
      if (splits > K / 128) splits = K / 128;
      while (splits > 1 && (int64_t)splits * M * N > workspace_floats) --splits;
      if (splits < 1) splits = 1;
    }
    int kchunk = K;
    float* partial = nullptr;
    if (splits > 1) {
      kchunk = ((K + splits - 1) / splits + VK - 1) / VK * VK;
      splits = (K + kchunk - 1) / kchunk;
      vgrid.z = splits;
      partial = workspace;
    }
    if (!transA && !transB) launch_pdl(gemm_f32_vec_kernel<true, true>, dim3(vgrid), dim3(256), 0, s, M, N, K, kchunk, A, lda, B, ldb, C, ldc, bias, beta, partial);
    else if (!transA && transB) launch_pdl(gemm_f32_vec_kernel<true, false>, dim3(vgrid), dim3(256), 0, s, M, N, K, kchunk, A, lda, B, ldb, C, ldc, bias, beta, partial);
    else if (transA && !transB) launch_pdl(gemm_f32_vec_kernel<false, true>, dim3(vgrid), dim3(256), 0, s, M, N, K, kchunk, A, lda, B, ldb, C, ldc, bias, beta, partial);
    else launch_pdl(gemm_f32_vec_kernel<false, false>, dim3(vgrid), dim3(256), 0, s, M, N, K, kchunk, A, lda, B, ldb, C, ldc, bias, beta, partial);
    if (partial) launch_splitk_reduce(partial, splits, M, N, C, ldc, bias, beta, s);
    SLNLP_LAUNCH_OK("gemm_f32");
    return 0;
  }
  if (!transA && !transB)
    launch_pdl(gemm_f32_kernel<true, true>, dim3(grid), dim3(256), 0, s, M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, bias, beta);
  else if (!transA && transB)
    launch_pdl(gemm_f32_kernel<true, false>, dim3(grid), dim3(256), 0, s, M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, bias, beta);
  else if (transA && !transB)
    launch_pdl(gemm_f32_kernel<false, true>, dim3(grid), dim3(256), 0, s, M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, bias, beta);
  else
    launch_pdl(gemm_f32_kernel<false, false>, dim3(grid), dim3(256), 0, s, M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, bias, beta);
  SLNLP_LAUNCH_OK("gemm_f32");
  return 0;
}
