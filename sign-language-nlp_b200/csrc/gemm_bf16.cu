// K3/K6 and the hoisted dW GEMMs on the 5th-gen tensor cores (the 2e-2 path):
// C[M,N] = op(A) op(B) + bias + beta*C with A, B read as fp32 from HBM, rounded to bf16
// on the way into shared memory (canonical K-major no-swizzle UMMA layout), multiplied
// by tcgen05.mma kind::f16 with fp32 accumulators in TMEM, and written back as fp32.
//
// 128 x TN x 64 tiles, 128 threads.  All threads stage operands (global -> registers one
// k-tile ahead -> bf16 -> shared, double buffered); thread 0 issues the MMAs; an mbarrier
// armed by tcgen05.commit hands each shared buffer back when its MMAs have read it, so
// the loads of tile i+1 overlap the MMAs of tile i.  Split-K (deterministic, partials in
// the caller's workspace) covers the dW GEMMs whose M*N is small and K = B*T is long.
#include "tc05.cuh"

namespace slnlp {

constexpr int TM = 128, TK = 64, GT = 128;  // tile rows, k-tile, threads

template <int TN, bool A_KCONTIG, bool B_NCONTIG>
__global__ void __launch_bounds__(GT) gemm_bf16_kernel(int M, int N, int Kfull, int kchunk,
                                                        const float* __restrict__ A, int64_t lda,
                                                        const float* __restrict__ B, int64_t ldb,
                                                        float* __restrict__ C, int ldc,
                                                        const float* __restrict__ bias, float beta,
                                                        float* __restrict__ partial) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int A_BYTES = TM * TK * 2, B_BYTES = TN * TK * 2;
  uint8_t* sA = smem_raw;                  // 2 stages
  uint8_t* sB = smem_raw + 2 * A_BYTES;    // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 2 * B_BYTES);  // free[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5, warp_u = warp_uniform();
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int kbeg = blockIdx.z * kchunk;
  const int K = min(Kfull, kbeg + kchunk);
  const int nk = (K - kbeg + TK - 1) / TK;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
  }
  if (warp == 0) tmem_alloc(tmem_slot, TN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // register staging of one k-tile: A = 8 chunks of 8 k per thread (row = tid),
  // B = TN*8/128 chunks per thread (row = tid % TN, k8 = tid / TN + (128/TN) * i)
  constexpr int NB = TN * 8 / GT;
  constexpr int BSTEP = GT / TN;
  float ra[8][8], rb[NB][8];
  const int am = m0 + tid;
  const int bn = n0 + (tid % TN);
  auto load_regs = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = k0 + i * 8;
      if (A_KCONTIG) {
        if (am < M && k < K) {
          const float4 v0 = __ldg(reinterpret_cast<const float4*>(A + (int64_t)am * lda + k));
          const float4 v1 = __ldg(reinterpret_cast<const float4*>(A + (int64_t)am * lda + k) + 1);
          ra[i][0] = v0.x; ra[i][1] = v0.y; ra[i][2] = v0.z; ra[i][3] = v0.w;
          ra[i][4] = v1.x; ra[i][5] = v1.y; ra[i][6] = v1.z; ra[i][7] = v1.w;
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u) ra[i][u] = 0.f;
        }
      } else {
#pragma unroll
        for (int u = 0; u < 8; ++u) ra[i][u] = (am < M && k + u < K) ? __ldg(A + (int64_t)(k + u) * lda + am) : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int k = k0 + (tid / TN + BSTEP * i) * 8;
      if (!B_NCONTIG) {
        if (bn < N && k < K) {
          const float4 v0 = __ldg(reinterpret_cast<const float4*>(B + (int64_t)bn * ldb + k));
          const float4 v1 = __ldg(reinterpret_cast<const float4*>(B + (int64_t)bn * ldb + k) + 1);
          rb[i][0] = v0.x; rb[i][1] = v0.y; rb[i][2] = v0.z; rb[i][3] = v0.w;
          rb[i][4] = v1.x; rb[i][5] = v1.y; rb[i][6] = v1.z; rb[i][7] = v1.w;
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u) rb[i][u] = 0.f;
        }
      } else {
#pragma unroll
        for (int u = 0; u < 8; ++u) rb[i][u] = (bn < N && k + u < K) ? __ldg(B + (int64_t)(k + u) * ldb + bn) : 0.f;
      }
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      *reinterpret_cast<uint4*>(sA + buf * A_BYTES + canon_off(tid, i * 8, TM)) = pack8_bf16(ra[i]);
#pragma unroll
    for (int i = 0; i < NB; ++i)
      *reinterpret_cast<uint4*>(sB + buf * B_BYTES + canon_off(tid % TN, (tid / TN + BSTEP * i) * 8, TN)) =
          pack8_bf16(rb[i]);
  };

  constexpr uint32_t idesc = make_idesc(TM, TN);
  const uint64_t descA0 = make_desc(smem_u32(sA), TM * 16, 128), descB0 = make_desc(smem_u32(sB), TN * 16, 128);

  if (nk > 0) load_regs(kbeg);
  for (int i = 0; i < nk; ++i) {
    const int buf = i & 1;
    if (i >= 2) mbar_wait(&bars[buf], (uint32_t)((i >> 1) - 1) & 1u);  // MMAs of tile i-2 have read this buffer
    store_smem(buf);
    if (i + 1 < nk) load_regs(kbeg + (i + 1) * TK);  // next tile's loads fly while this tile's MMAs run
    fence_async_smem();
    __syncthreads();
    if (warp_u == 0 && elect_one()) {
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < TK / 16; ++kk)
        umma_bf16(tmem, descA0 + (uint64_t)((buf * A_BYTES + kk * 2 * (TM * 16)) >> 4),
                  descB0 + (uint64_t)((buf * B_BYTES + kk * 2 * (TN * 16)) >> 4), idesc, (i > 0 || kk > 0) ? 1u : 0u);
      umma_commit(&bars[buf]);
    }
  }
  float* P = partial ? partial + (int64_t)blockIdx.z * M * N : nullptr;
  if (nk > 0) {
    const int last = nk - 1;
    mbar_wait(&bars[last & 1], (uint32_t)(last >> 1) & 1u);  // the last commit covers every MMA issued
    tc_fence_after();
  }
  // Epilogue: TMEM -> registers -> shared (the operand buffers are free now) -> coalesced rows.
  // A thread owns a TMEM lane (= a row of C), so storing straight from the tcgen05.ld registers
  // would touch 32 different rows per instruction; staging lets a warp write 128 contiguous
  // bytes of ONE row per instruction instead.
  constexpr int SLD = TN + 1;
  float* stage = reinterpret_cast<float*>(smem_raw) + warp * 32 * SLD;
  static_assert(4 * 32 * SLD * 4 <= 2 * A_BYTES + 2 * B_BYTES, "staging tile must fit the operand buffers");
  const int lane = tid & 31;
#pragma unroll 1
  for (int c0 = 0; c0 < TN; c0 += 16) {
    float v[16];
    if (nk > 0) {
      tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    } else {
#pragma unroll
      for (int q = 0; q < 16; ++q) v[q] = 0.f;
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) stage[lane * SLD + c0 + q] = v[q];
  }
  __syncwarp();
  float bv[TN / 32];
#pragma unroll
  for (int j = 0; j < TN / 32; ++j) {
    const int n = n0 + lane + 32 * j;
    bv[j] = (bias && !P && n < N) ? bias[n] : 0.f;
  }
  const int mrow0 = m0 + warp * 32;
  // old C for 8 rows is loaded before any of their stores: interleaved load/store pairs
  // serialise on the load latency (3x slower for beta != 0)
  const bool rmw = !P && beta != 0.f;
#pragma unroll 1
  for (int r0 = 0; r0 < 32; r0 += 8) {
    float cold[8][TN / 32];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int j = 0; j < TN / 32; ++j) {
        const int m = mrow0 + r0 + r, n = n0 + lane + 32 * j;
        cold[r][j] = (rmw && m < M && n < N) ? C[(int64_t)m * ldc + n] : 0.f;
      }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int m = mrow0 + r0 + r;
      if (m >= M) break;
#pragma unroll
      for (int j = 0; j < TN / 32; ++j) {
        const int n = n0 + lane + 32 * j;
        if (n >= N) continue;
        const float acc = stage[(r0 + r) * SLD + lane + 32 * j];
        if (P) P[(int64_t)m * N + n] = acc;
        else C[(int64_t)m * ldc + n] = acc + bv[j] + beta * cold[r][j];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TN);
}

// defined in gemm_f32.cu
void launch_splitk_reduce(const float* partial, int splits, int M, int N, float* C, int ldc, const float* bias,
                          float beta, cudaStream_t s);

}  // namespace slnlp

using namespace slnlp;

template <int TN>
static void launch_bf16(int transA, int transB, dim3 grid, size_t sm, cudaStream_t s, int M, int N, int K, int kchunk,
                        const float* A, int lda, const float* B, int ldb, float* C, int ldc, const float* bias,
                        float beta, float* partial) {
#define SLNLP_GO(AK, BN)                                                                                          \
  do {                                                                                                            \
    cudaFuncSetAttribute(gemm_bf16_kernel<TN, AK, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);      \
    launch_pdl(gemm_bf16_kernel<TN, AK, BN>, dim3(grid), dim3(GT), sm, s, M, N, K, kchunk, A, lda, B, ldb, C, ldc, bias, beta, partial); \
  } while (0)
  if (!transA && !transB) SLNLP_GO(true, true);
  else if (!transA && transB) SLNLP_GO(true, false);
  else if (transA && !transB) SLNLP_GO(false, true);
  else SLNLP_GO(false, false);
#undef SLNLP_GO
}

extern "C" int slnlp_gemm_bf16(int transA, int transB, int M, int N, int K, const float* A, int lda,
                               const float* B, int ldb, float* C, int ldc, const float* bias, float beta,
                               float* workspace, int64_t workspace_floats, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(A && B && C, "gemm_bf16: null pointer");
  SLNLP_CHECK_ARG(M >= 0 && N >= 0 && K >= 0 && ldc >= N, "gemm_bf16: bad shape M=%d N=%d K=%d ldc=%d", M, N, K, ldc);
  SLNLP_CHECK_ARG(lda >= (transA ? M : K) && ldb >= (transB ? K : N), "gemm_bf16: bad lda/ldb");
  if (M == 0 || N == 0) return 0;
  // float4 staging needs 16-byte aligned k-contiguous operands with K % 8 == 0; anything else
  // (and tiny problems) goes to the fp32 kernel - still CUDA, never a CPU path.
  const bool a_ok = transA || (((uintptr_t)A % 16 == 0) && lda % 4 == 0 && K % 8 == 0);
  const bool b_ok = !transB || (((uintptr_t)B % 16 == 0) && ldb % 4 == 0 && K % 8 == 0);
  if (!a_ok || !b_ok || M < 64 || N < 32 || K < 32)
    return slnlp_gemm_f32(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, beta, workspace, workspace_floats, stream);
  cudaStream_t s = as_stream(stream);
  constexpr int TN = 64;
  dim3 grid(ceil_div(N, TN), ceil_div(M, TM));
  SLNLP_CHECK_ARG(grid.y <= 65535, "gemm_bf16: M too large");
  const int tiles = grid.x * grid.y;
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
    kchunk = ((K + splits - 1) / splits + TK - 1) / TK * TK;
    splits = (K + kchunk - 1) / kchunk;
    grid.z = splits;
    partial = workspace;
  }
  const size_t sm = 2 * (TM * TK * 2) + 2 * (TN * TK * 2) + 64;
  launch_bf16<TN>(transA, transB, grid, sm, s, M, N, K, kchunk, A, lda, B, ldb, C, ldc, bias, beta, partial);
  if (partial) launch_splitk_reduce(partial, splits, M, N, C, ldc, bias, beta, s);
  SLNLP_LAUNCH_OK("gemm_bf16");
  return 0;
}
