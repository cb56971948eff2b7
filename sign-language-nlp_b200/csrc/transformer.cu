// K13: Transformer attention + LayerNorm kernels (nn.Transformer as used by
// model/transformer.py:40-45,82-87: post-norm, ReLU, LayerNorm eps 1e-5).
//
//   slnlp_mha_fwd / slnlp_mha_bwd   scaled-dot-product attention of nn.MultiheadAttention,
//       FlashAttention-style: scores, mask (causal and/or key padding), softmax, attention
//       dropout and P.V per (sequence, head, 64-query tile) with an online softmax over
//       64-key tiles; the S x S matrix never reaches HBM.  Backward recomputes P from the
//       saved row log-sum-exp: one pass per query tile for dQ, one per key tile for dK/dV,
//       so there are no atomics and the result is deterministic.
//   slnlp_add_layernorm_fwd / slnlp_layernorm_bwd   y = LN(x + res): the residual add is
//       fused into the normalisation, one warp per row, warp-shuffle reductions.
//
// fp32 FMA arithmetic: at the reference's shapes (S = tens of frames, head dim <= 256) the
// attention is a few MFLOP per (sequence, head) and bound by reading Q/K/V once; the
// projections around it are the GEMMs (gemm_tma.cu / gemm_f32.cu).
#include "common.cuh"

namespace slnlp {

struct MhaArgs {
  const float *q, *k, *v;
  int ldq, ldk, ldv;
  float* o;            // fwd: out; bwd: forward output (for D = rowsum(dO * O))
  const float* dout;   // bwd
  int ldo;
  float* lse;          // [B, nhead, Sq] row log-sum-exp (fwd: written, bwd: read)
  float* dvec;         // [B, nhead, Sq] D (bwd: written by the dQ pass, read by the dK/dV pass)
  float *dq, *dk, *dv; // bwd, same leading dimensions as q / k / v
  int B, Sq, Sk, nhead, dh, causal;
  const int64_t* key_tokens;  // [B, Sk] or NULL: key j of sequence b is masked when == pad_idx
  int64_t pad_idx;
  float scale, p_drop;
  const uint64_t* rng;
  uint32_t site;
};

// thread layout of every tile product: 256 threads = 16 x 16, thread (ty, tx) owns rows
// ty*RPT + r and columns tx + 16*c (strided columns keep shared-memory reads conflict-free)
template <int RPT>
struct Tile {
  static constexpr int N = 16 * RPT;   // tile edge (queries and keys)
};

__device__ __forceinline__ float half_warp_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// rows [r0, r0+N) of a [rows, ld] matrix (head columns [0, dh)) -> smem tile [N][dh+1], zero-filled
template <int N>
__device__ __forceinline__ void load_tile(float* s, const float* g, int ld, int r0, int rows, int dh, float mul) {
  const int st = dh + 1;
  const int total = N * (dh >> 2);
  // four 16-byte global loads in flight per thread before the first shared store (one L2 round trip
  // per batch instead of one per element)
  for (int e0 = threadIdx.x; e0 < total; e0 += 4 * blockDim.x) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * blockDim.x;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < total) {
        const int r = e / (dh >> 2), c4 = (e % (dh >> 2)) << 2;
        if (r0 + r < rows) v[u] = *reinterpret_cast<const float4*>(g + (int64_t)(r0 + r) * ld + c4);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * blockDim.x;
      if (e < total) {
        const int r = e / (dh >> 2), c4 = (e % (dh >> 2)) << 2;
        float* d = s + r * st + c4;
        d[0] = v[u].x * mul; d[1] = v[u].y * mul; d[2] = v[u].z * mul; d[3] = v[u].w * mul;
      }
    }
  }
}

// s[r][c] = <A[ty*RPT+r, :], Bm[tx+16c, :]>
template <int RPT>
__device__ __forceinline__ void tile_dot(const float* sA, const float* sB, int dh, int ty, int tx, float (&s)[RPT][RPT]) {
  const int st = dh + 1;
#pragma unroll
  for (int r = 0; r < RPT; ++r)
#pragma unroll
    for (int c = 0; c < RPT; ++c) s[r][c] = 0.f;
  const float* a = sA + ty * RPT * st;
  const float* b = sB + tx * st;
#pragma unroll 4
  for (int d = 0; d < dh; ++d) {
    float av[RPT], bv[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) av[r] = a[r * st + d];
#pragma unroll
    for (int c = 0; c < RPT; ++c) bv[c] = b[c * 16 * st + d];
#pragma unroll
    for (int r = 0; r < RPT; ++r)
#pragma unroll
      for (int c = 0; c < RPT; ++c) s[r][c] = fmaf(av[r], bv[c], s[r][c]);
  }
}

__device__ __forceinline__ bool masked(const MhaArgs& p, const int64_t* ktok, int i, int j) {
  if (j >= p.Sk) return true;
  if (p.causal && j > i) return true;
  return ktok != nullptr && ktok[j] == p.pad_idx;
}

// dropout keep-factor (0 or 1/(1-p)) of attention weight (b, h, i, j)
__device__ __forceinline__ float drop_factor(const MhaArgs& p, uint64_t seed, uint64_t step, int bh, int i, int j) {
  const uint64_t e = ((uint64_t)bh * p.Sq + i) * p.Sk + j;
  float u[4];
  philox_uniform4(seed, step, p.site, e >> 2, u);
  return u[e & 3] < 1.f - p.p_drop ? 1.f / (1.f - p.p_drop) : 0.f;
}

// grid (ceil(Sq/N), nhead, B), block 256.  smem: Q, K, V tiles [N][dh+1] + P [N][N+1].
template <int RPT, int NDC>
__global__ void __launch_bounds__(256, (NDC <= 4 ? 3 : 1)) mha_fwd_kernel(MhaArgs p) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int N = Tile<RPT>::N;
  extern __shared__ float sm[];
  const int dh = p.dh, st = dh + 1;
  float* sQ = sm;
  float* sK = sQ + N * st;
  float* sV = sK + N * st;
  float* sP = sV + N * st;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.x * N, h = blockIdx.y, b = blockIdx.z;
  const int bh = b * p.nhead + h;
  const float* q = p.q + (int64_t)b * p.Sq * p.ldq + h * dh;
  const float* k = p.k + (int64_t)b * p.Sk * p.ldk + h * dh;
  const float* v = p.v + (int64_t)b * p.Sk * p.ldv + h * dh;
  const int64_t* ktok = p.key_tokens ? p.key_tokens + (int64_t)b * p.Sk : nullptr;
  const bool drop = p.p_drop > 0.f;
  const uint64_t seed = drop ? p.rng[0] : 0, step = drop ? p.rng[1] : 0;

  load_tile<N>(sQ, q, p.ldq, i0, p.Sq, dh, p.scale);
  float m[RPT], l[RPT], acc[RPT][NDC];
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    m[r] = -INFINITY;
    l[r] = 0.f;
#pragma unroll
    for (int n = 0; n < NDC; ++n) acc[r][n] = 0.f;
  }
  for (int k0 = 0; k0 < p.Sk; k0 += N) {
    if (p.causal && k0 > i0 + N - 1) break;
    __syncthreads();
    load_tile<N>(sK, k, p.ldk, k0, p.Sk, dh, 1.f);
    load_tile<N>(sV, v, p.ldv, k0, p.Sk, dh, 1.f);
    __syncthreads();
    float s[RPT][RPT];
    tile_dot<RPT>(sQ, sK, dh, ty, tx, s);
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int i = i0 + ty * RPT + r;
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < RPT; ++c) {
        if (masked(p, ktok, i, k0 + tx + 16 * c)) s[r][c] = -INFINITY;
        mx = fmaxf(mx, s[r][c]);
      }
      mx = half_warp_max(mx);
      const float mn = fmaxf(m[r], mx);
      const float corr = mn == -INFINITY ? 1.f : expf(m[r] - mn);
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < RPT; ++c) {
        float e = mn == -INFINITY ? 0.f : expf(s[r][c] - mn);
        sum += e;
        if (drop && e != 0.f) e *= drop_factor(p, seed, step, bh, i, k0 + tx + 16 * c);
        sP[(ty * RPT + r) * (N + 1) + tx + 16 * c] = e;
      }
      sum = half_warp_sum(sum);
      l[r] = l[r] * corr + sum;
      m[r] = mn;
#pragma unroll
      for (int n = 0; n < NDC; ++n) acc[r][n] *= corr;
    }
    __syncthreads();
    // acc[r][n] += sum_j P[row r][j] V[j][tx + 16 n]
#pragma unroll 2
    for (int j = 0; j < N; ++j) {
      float pv[RPT], vv[NDC];
#pragma unroll
      for (int r = 0; r < RPT; ++r) pv[r] = sP[(ty * RPT + r) * (N + 1) + j];
#pragma unroll
      for (int n = 0; n < NDC; ++n) vv[n] = tx + 16 * n < dh ? sV[j * st + tx + 16 * n] : 0.f;
#pragma unroll
      for (int r = 0; r < RPT; ++r)
#pragma unroll
        for (int n = 0; n < NDC; ++n) acc[r][n] = fmaf(pv[r], vv[n], acc[r][n]);
    }
  }
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    const int i = i0 + ty * RPT + r;
    if (i >= p.Sq) continue;
    const float inv = 1.f / l[r];   // every key masked: 0/0 = NaN, as torch's softmax of all -inf
    float* o = p.o + ((int64_t)b * p.Sq + i) * p.ldo + h * dh;
#pragma unroll
    for (int n = 0; n < NDC; ++n)
      if (tx + 16 * n < dh) o[tx + 16 * n] = acc[r][n] * inv;
    if (tx == 0) p.lse[(int64_t)bh * p.Sq + i] = m[r] + logf(l[r]);
  }
}

// One (query tile, key tile) step of the backward pass: fills sP (dropped probabilities)
// and sDS (d scores) for the tile; both are [N][N+1], indexed [query][key].
template <int RPT>
__device__ __forceinline__ void bwd_tile(const MhaArgs& p, const float* sQ, const float* sK, const float* sV,
                                         const float* sDO, float* sP, float* sDS, const float* lse, const float* dvec,
                                         const int64_t* ktok, int bh, int i0, int k0, uint64_t seed, uint64_t step) {
  constexpr int N = Tile<RPT>::N;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const bool drop = p.p_drop > 0.f;
  float s[RPT][RPT], dp[RPT][RPT];
  tile_dot<RPT>(sQ, sK, p.dh, ty, tx, s);     // Q already carries the 1/sqrt(dh) scale
  tile_dot<RPT>(sDO, sV, p.dh, ty, tx, dp);
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    const int i = i0 + ty * RPT + r;
    const bool row_ok = i < p.Sq;
    const float L = row_ok ? lse[i] : 0.f, D = row_ok ? dvec[i] : 0.f;
#pragma unroll
    for (int c = 0; c < RPT; ++c) {
      const int j = k0 + tx + 16 * c;
      float pr = 0.f, dsc = 0.f;
      if (row_ok && !masked(p, ktok, i, j)) {
        pr = expf(s[r][c] - L);
        float f = 1.f;
        if (drop) f = drop_factor(p, seed, step, bh, i, j);
        dsc = pr * (dp[r][c] * f - D);
        pr *= f;
      }
      sP[(ty * RPT + r) * (N + 1) + tx + 16 * c] = pr;
      sDS[(ty * RPT + r) * (N + 1) + tx + 16 * c] = dsc;
    }
  }
}

// DKV = false: grid (ceil(Sq/N), nhead, B), writes D and dQ.  DKV = true: grid (ceil(Sk/N), nhead, B),
// writes dK and dV.  smem: four [N][dh+1] tiles + two [N][N+1] tiles.
template <int RPT, int NDC, bool DKV>
__global__ void __launch_bounds__(256) mha_bwd_kernel(MhaArgs p) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int N = Tile<RPT>::N;
  extern __shared__ float sm[];
  const int dh = p.dh, st = dh + 1;
  float* sQ = sm;
  float* sK = sQ + N * st;
  float* sV = sK + N * st;
  float* sDO = sV + N * st;
  float* sP = sDO + N * st;
  float* sDS = sP + N * (N + 1);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int h = blockIdx.y, b = blockIdx.z, bh = b * p.nhead + h;
  const float* q = p.q + (int64_t)b * p.Sq * p.ldq + h * dh;
  const float* k = p.k + (int64_t)b * p.Sk * p.ldk + h * dh;
  const float* v = p.v + (int64_t)b * p.Sk * p.ldv + h * dh;
  const float* dO = p.dout + (int64_t)b * p.Sq * p.ldo + h * dh;
  const float* lse = p.lse + (int64_t)bh * p.Sq;
  float* dvec = p.dvec + (int64_t)bh * p.Sq;
  const int64_t* ktok = p.key_tokens ? p.key_tokens + (int64_t)b * p.Sk : nullptr;
  const bool drop = p.p_drop > 0.f;
  const uint64_t seed = drop ? p.rng[0] : 0, step = drop ? p.rng[1] : 0;
  float acc[RPT][NDC], acc2[RPT][NDC];
#pragma unroll
  for (int r = 0; r < RPT; ++r)
#pragma unroll
    for (int n = 0; n < NDC; ++n) acc[r][n] = acc2[r][n] = 0.f;

  if (!DKV) {
    const int i0 = blockIdx.x * N;
    load_tile<N>(sQ, q, p.ldq, i0, p.Sq, dh, p.scale);
    load_tile<N>(sDO, dO, p.ldo, i0, p.Sq, dh, 1.f);
    // D_i = <dO_i, O_i>, one half-warp per row
    const float* O = p.o + (int64_t)b * p.Sq * p.ldo + h * dh;
    for (int r = ty; r < N; r += 16) {
      float s = 0.f;
      if (i0 + r < p.Sq)
        for (int d = tx; d < dh; d += 16) s = fmaf(dO[(int64_t)(i0 + r) * p.ldo + d], O[(int64_t)(i0 + r) * p.ldo + d], s);
      s = half_warp_sum(s);
      if (tx == 0 && i0 + r < p.Sq) dvec[i0 + r] = s;
    }
    __syncthreads();   // dvec rows of this tile are re-read below (same CTA wrote them)
    for (int k0 = 0; k0 < p.Sk; k0 += N) {
      if (p.causal && k0 > i0 + N - 1) break;
      __syncthreads();
      load_tile<N>(sK, k, p.ldk, k0, p.Sk, dh, 1.f);
      load_tile<N>(sV, v, p.ldv, k0, p.Sk, dh, 1.f);
      __syncthreads();
      bwd_tile<RPT>(p, sQ, sK, sV, sDO, sP, sDS, lse, dvec, ktok, bh, i0, k0, seed, step);
      __syncthreads();
      // dQ[i][d] += sum_j dS[i][j] K[j][d]
#pragma unroll 2
      for (int j = 0; j < N; ++j) {
        float a[RPT], kv[NDC];
#pragma unroll
        for (int r = 0; r < RPT; ++r) a[r] = sDS[(ty * RPT + r) * (N + 1) + j];
#pragma unroll
        for (int n = 0; n < NDC; ++n) kv[n] = tx + 16 * n < dh ? sK[j * st + tx + 16 * n] : 0.f;
#pragma unroll
        for (int r = 0; r < RPT; ++r)
#pragma unroll
          for (int n = 0; n < NDC; ++n) acc[r][n] = fmaf(a[r], kv[n], acc[r][n]);
      }
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int i = i0 + ty * RPT + r;
      if (i >= p.Sq) continue;
      float* g = p.dq + ((int64_t)b * p.Sq + i) * p.ldq + h * dh;
#pragma unroll
      for (int n = 0; n < NDC; ++n)
        if (tx + 16 * n < dh) g[tx + 16 * n] = acc[r][n] * p.scale;
    }
  } else {
    const int k0 = blockIdx.x * N;
    load_tile<N>(sK, k, p.ldk, k0, p.Sk, dh, 1.f);
    load_tile<N>(sV, v, p.ldv, k0, p.Sk, dh, 1.f);
    for (int i0 = 0; i0 < p.Sq; i0 += N) {
      if (p.causal && i0 + N - 1 < k0) continue;
      __syncthreads();
      load_tile<N>(sQ, q, p.ldq, i0, p.Sq, dh, p.scale);
      load_tile<N>(sDO, dO, p.ldo, i0, p.Sq, dh, 1.f);
      __syncthreads();
      bwd_tile<RPT>(p, sQ, sK, sV, sDO, sP, sDS, lse, dvec, ktok, bh, i0, k0, seed, step);
      __syncthreads();
      // dV[j][d] += sum_i P[i][j] dO[i][d];  dK[j][d] += sum_i dS[i][j] Q[i][d]  (Q carries the scale)
#pragma unroll 2
      for (int i = 0; i < N; ++i) {
        float pj[RPT], dj[RPT], ov[NDC], qv[NDC];
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
          pj[r] = sP[i * (N + 1) + ty * RPT + r];
          dj[r] = sDS[i * (N + 1) + ty * RPT + r];
        }
#pragma unroll
        for (int n = 0; n < NDC; ++n) {
          const bool ok = tx + 16 * n < dh;
          ov[n] = ok ? sDO[i * st + tx + 16 * n] : 0.f;
          qv[n] = ok ? sQ[i * st + tx + 16 * n] : 0.f;
        }
#pragma unroll
        for (int r = 0; r < RPT; ++r)
#pragma unroll
          for (int n = 0; n < NDC; ++n) {
            acc[r][n] = fmaf(pj[r], ov[n], acc[r][n]);
            acc2[r][n] = fmaf(dj[r], qv[n], acc2[r][n]);
          }
      }
    }
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int j = k0 + ty * RPT + r;
      if (j >= p.Sk) continue;
      float* gv = p.dv + ((int64_t)b * p.Sk + j) * p.ldv + h * dh;
      float* gk = p.dk + ((int64_t)b * p.Sk + j) * p.ldk + h * dh;
#pragma unroll
      for (int n = 0; n < NDC; ++n)
        if (tx + 16 * n < dh) {
          gv[tx + 16 * n] = acc[r][n];
          gk[tx + 16 * n] = acc2[r][n];
        }
    }
  }
}

// ------------------------------------------------------------------ tensor-core attention
// The same flash-style passes with every tile product on the tensor cores (warp-level tf32 MMA,
// m16n8k8, fp32 accumulate; operands rounded to tf32 as they leave shared memory).  64-query x
// 64-key tiles, 4 warps, warp w owns query (or, in the dK/dV pass, key) rows 16w..16w+15, so the
// softmax needs only quad shuffles and P / dS go through warp-private shared-memory rows.
// A 64 x 64 x 64 tile is far below what tcgen05 needs to pay for its TMEM round trip and a
// per-(sequence, head) M = 128 tile would be half padding, hence the register-fragment form here.
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// fragment coordinates: g = lane / 4, t = lane % 4.
// C[16, 8 NT] += A[m0.., :] B^T with A stored [m][k] and B stored [n][k]
template <int NT, int KS>
__device__ __forceinline__ void mma_nk(float (&c)[NT][4], const float* sA, int sta, int m0, const float* sB, int stb,
                                       int g, int t) {
#pragma unroll
  for (int kk = 0; kk < KS; ++kk) {
    uint32_t a[4];
    a[0] = to_tf32(sA[(m0 + g) * sta + kk * 8 + t]);
    a[1] = to_tf32(sA[(m0 + g + 8) * sta + kk * 8 + t]);
    a[2] = to_tf32(sA[(m0 + g) * sta + kk * 8 + t + 4]);
    a[3] = to_tf32(sA[(m0 + g + 8) * sta + kk * 8 + t + 4]);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
      mma_tf32(c[nt], a, to_tf32(sB[(nt * 8 + g) * stb + kk * 8 + t]), to_tf32(sB[(nt * 8 + g) * stb + kk * 8 + t + 4]));
  }
}
// C[16, 8 NT] += A[m0.., :] B with A stored [m][k] and B stored [k][n]
template <int NT, int KS>
__device__ __forceinline__ void mma_kn(float (&c)[NT][4], const float* sA, int sta, int m0, const float* sB, int stb,
                                       int g, int t) {
#pragma unroll
  for (int kk = 0; kk < KS; ++kk) {
    uint32_t a[4];
    a[0] = to_tf32(sA[(m0 + g) * sta + kk * 8 + t]);
    a[1] = to_tf32(sA[(m0 + g + 8) * sta + kk * 8 + t]);
    a[2] = to_tf32(sA[(m0 + g) * sta + kk * 8 + t + 4]);
    a[3] = to_tf32(sA[(m0 + g + 8) * sta + kk * 8 + t + 4]);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
      mma_tf32(c[nt], a, to_tf32(sB[(kk * 8 + t) * stb + nt * 8 + g]), to_tf32(sB[(kk * 8 + t + 4) * stb + nt * 8 + g]));
  }
}
// C[16, 8 NT] += A^T[m0.., :] B with A stored [k][m] and B stored [k][n]
template <int NT, int KS>
__device__ __forceinline__ void mma_tkn(float (&c)[NT][4], const float* sA, int sta, int m0, const float* sB, int stb,
                                        int g, int t) {
#pragma unroll
  for (int kk = 0; kk < KS; ++kk) {
    uint32_t a[4];
    a[0] = to_tf32(sA[(kk * 8 + t) * sta + m0 + g]);
    a[1] = to_tf32(sA[(kk * 8 + t) * sta + m0 + g + 8]);
    a[2] = to_tf32(sA[(kk * 8 + t + 4) * sta + m0 + g]);
    a[3] = to_tf32(sA[(kk * 8 + t + 4) * sta + m0 + g + 8]);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
      mma_tf32(c[nt], a, to_tf32(sB[(kk * 8 + t) * stb + nt * 8 + g]), to_tf32(sB[(kk * 8 + t + 4) * stb + nt * 8 + g]));
  }
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// keep-factors of the two adjacent attention weights (b, h, i, j) and (b, h, i, j + 1): the same
// Philox stream as drop_factor(), one block of four uniforms serving both whenever they share it
__device__ __forceinline__ void drop_factor2(const MhaArgs& p, uint64_t seed, uint64_t step, int bh, int i, int j,
                                             float& f0, float& f1) {
  const uint64_t e = ((uint64_t)bh * p.Sq + i) * p.Sk + j;
  const int ph = (int)(e & 3);
  const float keep = 1.f - p.p_drop, inv = 1.f / keep;
  float u[4];
  philox_uniform4(seed, step, p.site, e >> 2, u);
  const float u0 = ph == 0 ? u[0] : ph == 1 ? u[1] : ph == 2 ? u[2] : u[3];
  float u1 = ph == 0 ? u[1] : ph == 1 ? u[2] : u[3];
  if (ph == 3) {
    philox_uniform4(seed, step, p.site, (e >> 2) + 1, u);
    u1 = u[0];
  }
  f0 = u0 < keep ? inv : 0.f;
  f1 = u1 < keep ? inv : 0.f;
}
// bit (2 n + e) set: key k0 + 8 n + 2 t + e is past the end or a padding token
__device__ __forceinline__ uint32_t key_mask_bits(const MhaArgs& p, const int64_t* ktok, int k0, int t) {
  int64_t tok[16];
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = k0 + n * 8 + 2 * t + e;
      tok[2 * n + e] = (ktok != nullptr && j < p.Sk) ? ktok[j] : p.pad_idx + 1;
    }
  uint32_t bits = 0;
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = k0 + n * 8 + 2 * t + e;
      if (j >= p.Sk || tok[2 * n + e] == p.pad_idx) bits |= 1u << (2 * n + e);
    }
  return bits;
}

// NT tiles at once: U float4 loads per tile and thread are in flight before the first shared
// store, so a CTA pays one L2 / HBM round trip per pass for all its operands instead of one per tile
struct TileSrc {
  float* s;
  const float* g;
  int ld, r0, rows;
  float mul;
};
template <int NT, int U>
__device__ __forceinline__ void load_tiles_v(const TileSrc (&ts)[NT], int dh, int st) {
  const int c4n = dh >> 2, total = 64 * c4n;
  for (int e0 = threadIdx.x; e0 < total; e0 += U * blockDim.x) {
    float4 v[NT][U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = e0 + u * blockDim.x;
      const int r = e / c4n, c4 = (e - r * c4n) << 2;
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        v[n][u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < total && ts[n].r0 + r < ts[n].rows)
          v[n][u] = *reinterpret_cast<const float4*>(ts[n].g + (int64_t)(ts[n].r0 + r) * ts[n].ld + c4);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = e0 + u * blockDim.x;
      const int r = e / c4n, c4 = (e - r * c4n) << 2;
      if (e < total) {
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          const float m = ts[n].mul;
          *reinterpret_cast<float4*>(ts[n].s + r * st + c4) = make_float4(v[n][u].x * m, v[n][u].y * m, v[n][u].z * m, v[n][u].w * m);
        }
      }
    }
  }
}

constexpr int TC_N = 64;        // tile edge
constexpr int TC_SP = TC_N + 4; // P / dS row stride

// grid (ceil(Sq/64), nhead, B), block 128.  smem: Q, K, V [64][DH+4], P [64][68]
template <int DH>
__global__ void __launch_bounds__(128) mha_tc_fwd_kernel(MhaArgs p) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int N = TC_N, ST = DH + 4, SP = TC_SP, NTD = DH / 8;
  extern __shared__ float sm[];
  float* sQ = sm;
  float* sK = sQ + N * ST;
  float* sV = sK + N * ST;
  float* sP = sV + N * ST;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int i0 = blockIdx.x * N, h = blockIdx.y, b = blockIdx.z;
  const int bh = b * p.nhead + h;
  const float* q = p.q + (int64_t)b * p.Sq * p.ldq + h * DH;
  const float* k = p.k + (int64_t)b * p.Sk * p.ldk + h * DH;
  const float* v = p.v + (int64_t)b * p.Sk * p.ldv + h * DH;
  const int64_t* ktok = p.key_tokens ? p.key_tokens + (int64_t)b * p.Sk : nullptr;
  const bool drop = p.p_drop > 0.f;
  const uint64_t seed = drop ? p.rng[0] : 0, step = drop ? p.rng[1] : 0;

  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  float acc[NTD][4];
#pragma unroll
  for (int n = 0; n < NTD; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
  for (int k0 = 0; k0 < p.Sk; k0 += N) {
    if (p.causal && k0 > i0 + N - 1) break;
    const uint32_t kbits = key_mask_bits(p, ktok, k0, t);
    __syncthreads();
    if (k0 == 0) {
      const TileSrc ts[3] = {{sQ, q, p.ldq, i0, p.Sq, p.scale}, {sK, k, p.ldk, k0, p.Sk, 1.f}, {sV, v, p.ldv, k0, p.Sk, 1.f}};
      load_tiles_v<3, 4>(ts, DH, ST);
    } else {
      const TileSrc ts[2] = {{sK, k, p.ldk, k0, p.Sk, 1.f}, {sV, v, p.ldv, k0, p.Sk, 1.f}};
      load_tiles_v<2, 4>(ts, DH, ST);
    }
    __syncthreads();
    float s[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
    mma_nk<8, DH / 8>(s, sQ, ST, 16 * w, sK, ST, g, t);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = 16 * w + g + 8 * r, i = i0 + row;
      float mx = -INFINITY;
#pragma unroll
      for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = k0 + n * 8 + 2 * t + e;
          if (((kbits >> (2 * n + e)) & 1u) || (p.causal && j > i)) s[n][2 * r + e] = -INFINITY;
          mx = fmaxf(mx, s[n][2 * r + e]);
        }
      mx = quad_max(mx);
      const float mn = fmaxf(m[r], mx);
      const float corr = mn == -INFINITY ? 1.f : expf(m[r] - mn);
      float sum = 0.f;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        float ev[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          ev[e] = mn == -INFINITY ? 0.f : expf(s[n][2 * r + e] - mn);
          sum += ev[e];
        }
        if (drop && (ev[0] != 0.f || ev[1] != 0.f)) {
          float f0, f1;
          drop_factor2(p, seed, step, bh, i, k0 + n * 8 + 2 * t, f0, f1);
          ev[0] *= f0;
          ev[1] *= f1;
        }
        *reinterpret_cast<float2*>(sP + row * SP + n * 8 + 2 * t) = make_float2(ev[0], ev[1]);
      }
      sum = quad_sum(sum);
      l[r] = l[r] * corr + sum;
      m[r] = mn;
#pragma unroll
      for (int n = 0; n < NTD; ++n) {
        acc[n][2 * r] *= corr;
        acc[n][2 * r + 1] *= corr;
      }
    }
    __syncwarp();
    mma_kn<NTD, 8>(acc, sP, SP, 16 * w, sV, ST, g, t);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = i0 + 16 * w + g + 8 * r;
    if (i >= p.Sq) continue;
    const float inv = 1.f / l[r];   // every key masked: 0/0 = NaN, as torch's softmax of all -inf
    float* o = p.o + ((int64_t)b * p.Sq + i) * p.ldo + h * DH;
#pragma unroll
    for (int n = 0; n < NTD; ++n)
      *reinterpret_cast<float2*>(o + n * 8 + 2 * t) = make_float2(acc[n][2 * r] * inv, acc[n][2 * r + 1] * inv);
    if (t == 0) p.lse[(int64_t)bh * p.Sq + i] = m[r] + logf(l[r]);
  }
}

// P (dropped) and dS of one (query tile, key tile) pair into the warp's 16 query rows of sP / sDS
template <int DH>
__device__ __forceinline__ void tc_bwd_tile(const MhaArgs& p, const float* sQ, const float* sK, const float* sV,
                                            const float* sDO, float* sP, float* sDS, const float* lse, const float* dvec,
                                            const int64_t* ktok, int bh, int i0, int k0, uint64_t seed, uint64_t step,
                                            int w, int g, int t) {
  constexpr int ST = DH + 4, SP = TC_SP;
  const bool drop = p.p_drop > 0.f;
  // row statistics and key mask are fetched before the tile products so that their latency hides
  float Lr[2], Dr[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = i0 + 16 * w + g + 8 * r;
    Lr[r] = i < p.Sq ? lse[i] : 0.f;
    Dr[r] = i < p.Sq ? dvec[i] : 0.f;
  }
  const uint32_t kbits = key_mask_bits(p, ktok, k0, t);
  float s[8][4], dp[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
    dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
  }
  mma_nk<8, DH / 8>(s, sQ, ST, 16 * w, sK, ST, g, t);      // Q carries the 1/sqrt(dh) scale
  mma_nk<8, DH / 8>(dp, sDO, ST, 16 * w, sV, ST, g, t);
  __syncthreads();   // sP / sDS reuse the K / V tiles' shared memory: every warp is done reading them
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = 16 * w + g + 8 * r, i = i0 + row;
    const bool row_ok = i < p.Sq;
    const float L = Lr[r], D = Dr[r];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      float pr[2], dsc[2], f[2] = {1.f, 1.f};
      bool live[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = k0 + n * 8 + 2 * t + e;
        live[e] = row_ok && !((kbits >> (2 * n + e)) & 1u) && !(p.causal && j > i);
      }
      if (drop && (live[0] || live[1])) drop_factor2(p, seed, step, bh, i, k0 + n * 8 + 2 * t, f[0], f[1]);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        pr[e] = dsc[e] = 0.f;
        if (live[e]) {
          pr[e] = expf(s[n][2 * r + e] - L);
          dsc[e] = pr[e] * (dp[n][2 * r + e] * f[e] - D);
          pr[e] *= f[e];
        }
      }
      if (sP) *reinterpret_cast<float2*>(sP + row * SP + n * 8 + 2 * t) = make_float2(pr[0], pr[1]);
      *reinterpret_cast<float2*>(sDS + row * SP + n * 8 + 2 * t) = make_float2(dsc[0], dsc[1]);
    }
  }
}

// DKV = false: grid (ceil(Sq/64), nhead, B), writes D and dQ.  DKV = true: grid (ceil(Sk/64), nhead, B),
// writes dK and dV.  block 128.  smem: Q, dO [64][DH+4] and two more tiles of max([64][DH+4],
// [64][68]) floats that hold K / V for the tile products and then P / dS (dQ pass: K stays, dS
// takes V's place) - four tiles, so that three CTAs fit an SM and 400 (sequence, head) CTAs are
// one wave.
template <int DH, bool DKV>
__global__ void __launch_bounds__(128, 3) mha_tc_bwd_kernel(MhaArgs p) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int N = TC_N, ST = DH + 4, SP = TC_SP, NTD = DH / 8;
  constexpr int TS = N * (ST > SP ? ST : SP);
  extern __shared__ float sm[];
  float* sQ = sm;
  float* sDO = sQ + N * ST;
  float* sK = sDO + N * ST;
  float* sV = sK + TS;
  float* sP = DKV ? sK : nullptr;
  float* sDS = sV;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int h = blockIdx.y, b = blockIdx.z, bh = b * p.nhead + h;
  const float* q = p.q + (int64_t)b * p.Sq * p.ldq + h * DH;
  const float* k = p.k + (int64_t)b * p.Sk * p.ldk + h * DH;
  const float* v = p.v + (int64_t)b * p.Sk * p.ldv + h * DH;
  const float* dO = p.dout + (int64_t)b * p.Sq * p.ldo + h * DH;
  const float* lse = p.lse + (int64_t)bh * p.Sq;
  float* dvec = p.dvec + (int64_t)bh * p.Sq;
  const int64_t* ktok = p.key_tokens ? p.key_tokens + (int64_t)b * p.Sk : nullptr;
  const bool drop = p.p_drop > 0.f;
  const uint64_t seed = drop ? p.rng[0] : 0, step = drop ? p.rng[1] : 0;
  float acc[NTD][4];
#pragma unroll
  for (int n = 0; n < NTD; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;

  if (!DKV) {
    const int i0 = blockIdx.x * N;
    // D_i = <dO_i, O_i>: two threads per row, each half a row of independent float4 loads
    {
      const float* O = p.o + (int64_t)b * p.Sq * p.ldo + h * DH;
      const int r = threadIdx.x >> 1, half = threadIdx.x & 1;
      constexpr int HW = DH / 2;
      float sacc = 0.f;
      if (i0 + r < p.Sq) {
        const float4* a4 = reinterpret_cast<const float4*>(dO + (int64_t)(i0 + r) * p.ldo + half * HW);
        const float4* b4 = reinterpret_cast<const float4*>(O + (int64_t)(i0 + r) * p.ldo + half * HW);
        float4 av[HW / 4], bv[HW / 4];
#pragma unroll
        for (int c = 0; c < HW / 4; ++c) {
          av[c] = a4[c];
          bv[c] = b4[c];
        }
#pragma unroll
        for (int c = 0; c < HW / 4; ++c)
          sacc += av[c].x * bv[c].x + av[c].y * bv[c].y + av[c].z * bv[c].z + av[c].w * bv[c].w;
      }
      sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
      if (half == 0 && i0 + r < p.Sq) dvec[i0 + r] = sacc;
    }
    __syncthreads();   // dvec rows of this tile are re-read below (same CTA wrote them)
    for (int k0 = 0; k0 < p.Sk; k0 += N) {
      if (p.causal && k0 > i0 + N - 1) break;
      __syncthreads();
      if (k0 == 0) {
        const TileSrc ts[4] = {{sQ, q, p.ldq, i0, p.Sq, p.scale}, {sDO, dO, p.ldo, i0, p.Sq, 1.f},
                               {sK, k, p.ldk, k0, p.Sk, 1.f}, {sV, v, p.ldv, k0, p.Sk, 1.f}};
        load_tiles_v<4, 2>(ts, DH, ST);
      } else {
        const TileSrc ts[2] = {{sK, k, p.ldk, k0, p.Sk, 1.f}, {sV, v, p.ldv, k0, p.Sk, 1.f}};
        load_tiles_v<2, 4>(ts, DH, ST);
      }
      __syncthreads();
      tc_bwd_tile<DH>(p, sQ, sK, sV, sDO, sP, sDS, lse, dvec, ktok, bh, i0, k0, seed, step, w, g, t);
      __syncwarp();
      mma_kn<NTD, 8>(acc, sDS, SP, 16 * w, sK, ST, g, t);   // dQ += dS K
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int i = i0 + 16 * w + g + 8 * r;
      if (i >= p.Sq) continue;
      float* gq = p.dq + ((int64_t)b * p.Sq + i) * p.ldq + h * DH;
#pragma unroll
      for (int n = 0; n < NTD; ++n)
        *reinterpret_cast<float2*>(gq + n * 8 + 2 * t) = make_float2(acc[n][2 * r] * p.scale, acc[n][2 * r + 1] * p.scale);
    }
  } else {
    float acc2[NTD][4];
#pragma unroll
    for (int n = 0; n < NTD; ++n) acc2[n][0] = acc2[n][1] = acc2[n][2] = acc2[n][3] = 0.f;
    const int k0 = blockIdx.x * N;
    for (int i0 = 0; i0 < p.Sq; i0 += N) {
      if (p.causal && i0 + N - 1 < k0) continue;
      __syncthreads();
      {   // K / V again for every query tile: P / dS of the previous one lived in their place
        const TileSrc ts[4] = {{sQ, q, p.ldq, i0, p.Sq, p.scale}, {sDO, dO, p.ldo, i0, p.Sq, 1.f},
                               {sK, k, p.ldk, k0, p.Sk, 1.f}, {sV, v, p.ldv, k0, p.Sk, 1.f}};
        load_tiles_v<4, 2>(ts, DH, ST);
      }
      __syncthreads();
      tc_bwd_tile<DH>(p, sQ, sK, sV, sDO, sP, sDS, lse, dvec, ktok, bh, i0, k0, seed, step, w, g, t);
      __syncthreads();
      mma_tkn<NTD, 8>(acc, sP, SP, 16 * w, sDO, ST, g, t);    // dV += P^T dO
      mma_tkn<NTD, 8>(acc2, sDS, SP, 16 * w, sQ, ST, g, t);   // dK += dS^T Q  (Q carries the scale)
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int j = k0 + 16 * w + g + 8 * r;
      if (j >= p.Sk) continue;
      float* gv = p.dv + ((int64_t)b * p.Sk + j) * p.ldv + h * DH;
      float* gk = p.dk + ((int64_t)b * p.Sk + j) * p.ldk + h * DH;
#pragma unroll
      for (int n = 0; n < NTD; ++n) {
        *reinterpret_cast<float2*>(gv + n * 8 + 2 * t) = make_float2(acc[n][2 * r], acc[n][2 * r + 1]);
        *reinterpret_cast<float2*>(gk + n * 8 + 2 * t) = make_float2(acc2[n][2 * r], acc2[n][2 * r + 1]);
      }
    }
  }
}

static bool mha_tc_ok(const MhaArgs& a) { return a.dh == 16 || a.dh == 32 || a.dh == 64; }

template <int DH>
static int launch_mha_tc_fwd(const MhaArgs& a, cudaStream_t s) {
  constexpr size_t smem = (size_t)(3 * TC_N * (DH + 4) + TC_N * TC_SP) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(mha_tc_fwd_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return fail("mha_tf32_fwd: cannot raise the shared-memory limit");
    attr = true;
  }
  launch_pdl(mha_tc_fwd_kernel<DH>, dim3(ceil_div(a.Sq, TC_N), a.nhead, a.B), dim3(128), smem, s, a);
  SLNLP_LAUNCH_OK("mha_tf32_fwd");
  return 0;
}

template <int DH>
static int launch_mha_tc_bwd(const MhaArgs& a, cudaStream_t s) {
  constexpr size_t smem = (size_t)(2 * TC_N * (DH + 4) + 2 * TC_N * (DH + 4 > TC_SP ? DH + 4 : TC_SP)) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(mha_tc_bwd_kernel<DH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaFuncSetAttribute(mha_tc_bwd_kernel<DH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return fail("mha_tf32_bwd: cannot raise the shared-memory limit");
    attr = true;
  }
  launch_pdl(mha_tc_bwd_kernel<DH, false>, dim3(ceil_div(a.Sq, TC_N), a.nhead, a.B), dim3(128), smem, s, a);
  SLNLP_LAUNCH_OK("mha_tf32_bwd(dq)");
  launch_pdl(mha_tc_bwd_kernel<DH, true>, dim3(ceil_div(a.Sk, TC_N), a.nhead, a.B), dim3(128), smem, s, a);
  SLNLP_LAUNCH_OK("mha_tf32_bwd(dkv)");
  return 0;
}

// ------------------------------------------------------------------ few-query attention
// The reference's decoder sees ONE target position (the label, model/transformer.py:82-87), so its
// self-attention is 1 x 1 and its cross-attention 1 x S: a 64-query tile would be 63/64 padding.
// One CTA per (sequence, head): all 128 threads stage 64-key K / V tiles in shared memory, warp i
// then owns query row i (Sq <= 4) - lanes over keys for the scores, lanes over head columns for
// P.V and dQ - and all threads write the dK / dV tile.  Same masks, dropout stream, log-sum-exp
// and all-masked-row NaN as the tile kernels; fp32 throughout.
constexpr int MHA_SMALL_SQ = 4;
constexpr int MHA_SMALL_C = 8;   // head columns per lane: dh <= 256

// rows [k0, k0+nk) of K and V -> smem tiles [64][dh+1]; the loads of both tiles are in flight
// together (one L2 round trip), rows past nk are left untouched (never read)
__device__ __forceinline__ void stage_kv(float* sK, float* sV, const float* k, const float* v, int ldk, int ldv,
                                         int k0, int nk, int dh) {
  const int st = dh + 1, c4n = dh >> 2, total = nk * c4n;
  for (int e0 = threadIdx.x; e0 < total; e0 += 4 * blockDim.x) {
    float4 a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * blockDim.x;
      if (e < total) {
        const int r = e / c4n, c4 = (e - r * c4n) << 2;
        a[u] = *reinterpret_cast<const float4*>(k + (int64_t)(k0 + r) * ldk + c4);
        b[u] = *reinterpret_cast<const float4*>(v + (int64_t)(k0 + r) * ldv + c4);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * blockDim.x;
      if (e < total) {
        const int r = e / c4n, c4 = (e - r * c4n) << 2;
        float* dk = sK + r * st + c4;
        float* dv = sV + r * st + c4;
        dk[0] = a[u].x; dk[1] = a[u].y; dk[2] = a[u].z; dk[3] = a[u].w;
        dv[0] = b[u].x; dv[1] = b[u].y; dv[2] = b[u].z; dv[3] = b[u].w;
      }
    }
  }
}

// smem: K, V [64][dh+1] | q [4][dh] | dO [4][dh] | p [4][64] | ds [4][64]
template <bool BWD>
__global__ void __launch_bounds__(128) mha_small_kernel(MhaArgs p) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ float sm[];
  const int dh = p.dh, st = dh + 1;
  float* sK = sm;
  float* sV = sK + 64 * st;
  float* sq = sV + 64 * st;
  float* sdo = sq + 4 * dh;
  float* sp = sdo + 4 * dh;
  float* sds = sp + 4 * 64;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int bh = blockIdx.x, b = bh / p.nhead, h = bh % p.nhead;
  const float* q = p.q + (int64_t)b * p.Sq * p.ldq + h * dh;
  const float* k = p.k + (int64_t)b * p.Sk * p.ldk + h * dh;
  const float* v = p.v + (int64_t)b * p.Sk * p.ldv + h * dh;
  const int64_t* ktok = p.key_tokens ? p.key_tokens + (int64_t)b * p.Sk : nullptr;
  const bool drop = p.p_drop > 0.f;
  const uint64_t seed = drop ? p.rng[0] : 0, step = drop ? p.rng[1] : 0;
  for (int idx = threadIdx.x; idx < p.Sq * dh; idx += blockDim.x) {
    const int i = idx / dh, d = idx - i * dh;
    sq[idx] = q[(int64_t)i * p.ldq + d] * p.scale;
    if (BWD) sdo[idx] = p.dout[((int64_t)b * p.Sq + i) * p.ldo + h * dh + d];
  }
  const bool active = w < p.Sq;
  const int i = w;   // this warp's query row
  float m = -INFINITY, l = 0.f, L = 0.f, D = 0.f;
  float acc[MHA_SMALL_C];
#pragma unroll
  for (int c = 0; c < MHA_SMALL_C; ++c) acc[c] = 0.f;
  if (BWD && active) {
    const float* dO = p.dout + ((int64_t)b * p.Sq + i) * p.ldo + h * dh;
    const float* O = p.o + ((int64_t)b * p.Sq + i) * p.ldo + h * dh;
    float ds = 0.f;
    for (int d = lane; d < dh; d += 32) ds = fmaf(dO[d], O[d], ds);
    D = warp_sum(ds);
    L = p.lse[(int64_t)bh * p.Sq + i];
    if (lane == 0) p.dvec[(int64_t)bh * p.Sq + i] = D;
  }
  for (int k0 = 0; k0 < p.Sk; k0 += 64) {
    const int nk = min(64, p.Sk - k0);   // keys present in this chunk (the 1 x 1 self-attention has one)
    const int j0 = k0 + lane, j1 = k0 + lane + 32;
    // key-padding tokens are fetched with the tiles, not when the mask is applied
    const int64_t tok0 = (ktok && j0 < p.Sk) ? ktok[j0] : p.pad_idx + 1;
    const int64_t tok1 = (ktok && j1 < p.Sk) ? ktok[j1] : p.pad_idx + 1;
    __syncthreads();
    stage_kv(sK, sV, k, v, p.ldk, p.ldv, k0, nk, dh);
    __syncthreads();
    if (active) {
      const float* qi = sq + i * dh;
      const float* di = sdo + i * dh;
      float s0 = 0.f, s1 = 0.f, dp0 = 0.f, dp1 = 0.f;
      const float* ka = sK + lane * st;
      const float* kb = sK + (lane + 32) * st;
      if (nk > 32) {
#pragma unroll 4
        for (int d = 0; d < dh; ++d) {
          s0 = fmaf(qi[d], ka[d], s0);
          s1 = fmaf(qi[d], kb[d], s1);
        }
      } else if (lane < nk) {
#pragma unroll 4
        for (int d = 0; d < dh; ++d) s0 = fmaf(qi[d], ka[d], s0);
      }
      if (BWD) {
        const float* va = sV + lane * st;
        const float* vb = sV + (lane + 32) * st;
        if (nk > 32) {
#pragma unroll 4
          for (int d = 0; d < dh; ++d) {
            dp0 = fmaf(di[d], va[d], dp0);
            dp1 = fmaf(di[d], vb[d], dp1);
          }
        } else if (lane < nk) {
#pragma unroll 4
          for (int d = 0; d < dh; ++d) dp0 = fmaf(di[d], va[d], dp0);
        }
      }
      const bool mk0 = j0 >= p.Sk || (p.causal && j0 > i) || tok0 == p.pad_idx;
      const bool mk1 = j1 >= p.Sk || (p.causal && j1 > i) || tok1 == p.pad_idx;
      float corr = 1.f;
      if (!BWD) {
        if (mk0) s0 = -INFINITY;
        if (mk1) s1 = -INFINITY;
        const float mx = warp_max(fmaxf(s0, s1));
        const float mn = fmaxf(m, mx);
        corr = mn == -INFINITY ? 1.f : expf(m - mn);
        float e0 = mn == -INFINITY ? 0.f : expf(s0 - mn);
        float e1 = mn == -INFINITY ? 0.f : expf(s1 - mn);
        const float sum = warp_sum(e0 + e1);
        if (drop && e0 != 0.f) e0 *= drop_factor(p, seed, step, bh, i, j0);
        if (drop && e1 != 0.f) e1 *= drop_factor(p, seed, step, bh, i, j1);
        sp[w * 64 + lane] = e0;
        sp[w * 64 + lane + 32] = e1;
        l = l * corr + sum;
        m = mn;
      } else {
        float pr0 = 0.f, pr1 = 0.f, ds0 = 0.f, ds1 = 0.f;
        if (!mk0) {
          pr0 = expf(s0 - L);
          const float f = drop ? drop_factor(p, seed, step, bh, i, j0) : 1.f;
          ds0 = pr0 * (dp0 * f - D);
          pr0 *= f;
        }
        if (!mk1) {
          pr1 = expf(s1 - L);
          const float f = drop ? drop_factor(p, seed, step, bh, i, j1) : 1.f;
          ds1 = pr1 * (dp1 * f - D);
          pr1 *= f;
        }
        sp[w * 64 + lane] = pr0;
        sp[w * 64 + lane + 32] = pr1;
        sds[w * 64 + lane] = ds0;
        sds[w * 64 + lane + 32] = ds1;
      }
      __syncwarp();
      // fwd: acc = acc * corr + P V;  bwd: acc += dS K   (lanes over head columns)
      const float* wrow = (BWD ? sds : sp) + w * 64;
      const float* tile = BWD ? sK : sV;
#pragma unroll
      for (int c = 0; c < MHA_SMALL_C; ++c) {
        const int d = lane + 32 * c;
        if (d < dh) {
          float a0 = acc[c] * corr, a1 = 0.f;
          int j = 0;
#pragma unroll 4
          for (; j + 1 < nk; j += 2) {
            a0 = fmaf(wrow[j], tile[j * st + d], a0);
            a1 = fmaf(wrow[j + 1], tile[(j + 1) * st + d], a1);
          }
          if (j < nk) a0 = fmaf(wrow[j], tile[j * st + d], a0);
          acc[c] = a0 + a1;
        }
      }
    }
    if (BWD) {
      __syncthreads();
      // dK[j] = sum_i dS[i][j] q_i (q carries the 1/sqrt(dh) scale), dV[j] = sum_i P[i][j] dO_i
      for (int idx = threadIdx.x; idx < 64 * dh; idx += blockDim.x) {
        const int j = idx / dh, d = idx - j * dh;
        if (k0 + j >= p.Sk) break;
        float gk = 0.f, gv = 0.f;
        for (int ii = 0; ii < p.Sq; ++ii) {
          gk = fmaf(sds[ii * 64 + j], sq[ii * dh + d], gk);
          gv = fmaf(sp[ii * 64 + j], sdo[ii * dh + d], gv);
        }
        p.dk[((int64_t)b * p.Sk + k0 + j) * p.ldk + h * dh + d] = gk;
        p.dv[((int64_t)b * p.Sk + k0 + j) * p.ldv + h * dh + d] = gv;
      }
    }
  }
  if (!active) return;
  if (!BWD) {
    const float inv = 1.f / l;   // every key masked: 0/0 = NaN, as torch's softmax of all -inf
    float* o = p.o + ((int64_t)b * p.Sq + i) * p.ldo + h * dh;
#pragma unroll
    for (int c = 0; c < MHA_SMALL_C; ++c)
      if (lane + 32 * c < dh) o[lane + 32 * c] = acc[c] * inv;
    if (lane == 0) p.lse[(int64_t)bh * p.Sq + i] = m + logf(l);
  } else {
    float* gq = p.dq + ((int64_t)b * p.Sq + i) * p.ldq + h * dh;
#pragma unroll
    for (int c = 0; c < MHA_SMALL_C; ++c)
      if (lane + 32 * c < dh) gq[lane + 32 * c] = acc[c] * p.scale;
  }
}

static bool mha_small_ok(const MhaArgs& a) { return a.Sq <= MHA_SMALL_SQ; }

template <bool BWD>
static int launch_mha_small(const MhaArgs& a, cudaStream_t s) {
  const size_t smem = (size_t)(2 * 64 * (a.dh + 1) + 8 * a.dh + 512) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(mha_small_kernel<BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return fail("mha(small): cannot raise the shared-memory limit");
    attr = true;
  }
  SLNLP_CHECK_ARG((int64_t)a.B * a.nhead <= 0x7fffffff, "mha: grid too large");
  launch_pdl(mha_small_kernel<BWD>, dim3(a.B * a.nhead), dim3(128), smem, s, a);
  SLNLP_LAUNCH_OK(BWD ? "mha_bwd(small)" : "mha_fwd(small)");
  return 0;
}

// ------------------------------------------------------------------ LayerNorm
constexpr int LN_MAX_VPL = 32;   // values per lane: E <= 1024
constexpr int LN_WARPS = 4;
// the kernels are instantiated for VPLT = 4, 8, 16, 32 values per lane (E <= 128, 256, 512, 1024) so that
// the per-lane arrays fit the register file with room for several CTAs per SM

// Every global load of a row is issued before the first use (loads in separate unrolled loops,
// out-of-range columns read as 0 through a select, no branch per element): interleaving load and
// use serialises VPLT DRAM round trips per row.

// y = LN(x + res) * gamma + beta; one warp per row.  mean / rstd [rows] saved for backward.
template <int VPLT>
__global__ void __launch_bounds__(LN_WARPS * 32) add_layernorm_fwd_kernel(const float* __restrict__ x,
                                                                        const float* __restrict__ res,
                                                                        const float* __restrict__ gamma,
                                                                        const float* __restrict__ beta,
                                                                        float* __restrict__ y, float* __restrict__ mean,
                                                                        float* __restrict__ rstd, int rows, int E, float eps) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int row = blockIdx.x * LN_WARPS + w;
  if (row >= rows) return;
  const float* xr = x + (int64_t)row * E;
  const float* rr = res ? res + (int64_t)row * E : xr;
  float v[VPLT], rv[VPLT], gm[VPLT], bt[VPLT];
#pragma unroll
  for (int k = 0; k < VPLT; ++k) v[k] = (lane + 32 * k < E) ? xr[lane + 32 * k] : 0.f;
#pragma unroll
  for (int k = 0; k < VPLT; ++k) rv[k] = (res && lane + 32 * k < E) ? rr[lane + 32 * k] : 0.f;
#pragma unroll
  for (int k = 0; k < VPLT; ++k) {
    gm[k] = (lane + 32 * k < E) ? gamma[lane + 32 * k] : 0.f;
    bt[k] = (lane + 32 * k < E) ? beta[lane + 32 * k] : 0.f;
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < VPLT; ++k) {
    v[k] += rv[k];
    s += v[k];
  }
  const float mu = warp_sum(s) / E;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < VPLT; ++k)
    if (lane + 32 * k < E) {
      const float d = v[k] - mu;
      q = fmaf(d, d, q);
    }
  const float rs = rsqrtf(warp_sum(q) / E + eps);
  float* yr = y + (int64_t)row * E;
#pragma unroll
  for (int k = 0; k < VPLT; ++k)
    if (lane + 32 * k < E) yr[lane + 32 * k] = (v[k] - mu) * rs * gm[k] + bt[k];
  if (lane == 0) {
    if (mean) mean[row] = mu;
    if (rstd) rstd[row] = rs;
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma, xhat = (x + res - mean) * rstd.
// dx is accumulated into when `accumulate` (the residual branch already holds a gradient).
// partials [gridDim.x, 2, E]: per-block sums of dy * xhat (d gamma) and dy (d beta).
template <int VPLT>
__global__ void __launch_bounds__(LN_WARPS * 32) layernorm_bwd_kernel(const float* __restrict__ dy,
                                                                    const float* __restrict__ x,
                                                                    const float* __restrict__ res,
                                                                    const float* __restrict__ gamma,
                                                                    const float* __restrict__ mean,
                                                                    const float* __restrict__ rstd, float* __restrict__ dx,
                                                                    float* __restrict__ partials, int rows, int E,
                                                                    int accumulate) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ float sred[];   // [LN_WARPS][2][E]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float dg[VPLT], db[VPLT], gm[VPLT];
#pragma unroll
  for (int k = 0; k < VPLT; ++k) {
    dg[k] = db[k] = 0.f;
    gm[k] = (lane + 32 * k < E) ? gamma[lane + 32 * k] : 0.f;
  }
  for (int row = blockIdx.x * LN_WARPS + w; row < rows; row += gridDim.x * LN_WARPS) {
    const float* xr = x + (int64_t)row * E;
    const float* rr = res ? res + (int64_t)row * E : xr;
    const float* gr = dy + (int64_t)row * E;
    float* o = dx + (int64_t)row * E;
    const float mu = mean[row], rs = rstd[row];
    float xh[VPLT], g[VPLT], rv[VPLT];
#pragma unroll
    for (int k = 0; k < VPLT; ++k) g[k] = (lane + 32 * k < E) ? gr[lane + 32 * k] : 0.f;
#pragma unroll
    for (int k = 0; k < VPLT; ++k) xh[k] = (lane + 32 * k < E) ? xr[lane + 32 * k] : 0.f;
#pragma unroll
    for (int k = 0; k < VPLT; ++k) rv[k] = (res && lane + 32 * k < E) ? rr[lane + 32 * k] : 0.f;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < VPLT; ++k) {
      const float d = g[k];
      xh[k] = (lane + 32 * k < E) ? (xh[k] + rv[k] - mu) * rs : 0.f;
      g[k] = d * gm[k];
      s1 += g[k];
      s2 = fmaf(g[k], xh[k], s2);
      dg[k] = fmaf(d, xh[k], dg[k]);
      db[k] += d;
    }
    if (accumulate) {   // old dx read before any of this row's stores
#pragma unroll
      for (int k = 0; k < VPLT; ++k) rv[k] = (lane + 32 * k < E) ? o[lane + 32 * k] : 0.f;
    } else {
#pragma unroll
      for (int k = 0; k < VPLT; ++k) rv[k] = 0.f;
    }
    s1 = warp_sum(s1) / E;
    s2 = warp_sum(s2) / E;
#pragma unroll
    for (int k = 0; k < VPLT; ++k)
      if (lane + 32 * k < E) o[lane + 32 * k] = rv[k] + rs * (g[k] - s1 - xh[k] * s2);
  }
#pragma unroll
  for (int k = 0; k < VPLT; ++k)
    if (lane + 32 * k < E) {
      sred[(w * 2 + 0) * E + lane + 32 * k] = dg[k];
      sred[(w * 2 + 1) * E + lane + 32 * k] = db[k];
    }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * E; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int ww = 0; ww < LN_WARPS; ++ww) s += sred[ww * 2 * E + c];
    partials[(int64_t)blockIdx.x * 2 * E + c] = s;
  }
}

static int ln_bwd_blocks(int rows) {
  int nb = ceil_div(rows, LN_WARPS);
  const int cap = 4 * (sm_count() > 0 ? sm_count() : 148);
  return nb < cap ? nb : cap;
}

template <int RPT, int NDC>
static int launch_mha_fwd(const MhaArgs& a, cudaStream_t s) {
  constexpr int N = Tile<RPT>::N;
  const size_t smem = (size_t)(3 * N * (a.dh + 1) + N * (N + 1)) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(mha_fwd_kernel<RPT, NDC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return fail("mha_fwd: cannot raise the shared-memory limit");
    attr = true;
  }
  dim3 grid(ceil_div(a.Sq, N), a.nhead, a.B);
  launch_pdl(mha_fwd_kernel<RPT, NDC>, dim3(grid), dim3(256), smem, s, a);
  SLNLP_LAUNCH_OK("mha_fwd");
  return 0;
}

template <int RPT, int NDC>
static int launch_mha_bwd(const MhaArgs& a, cudaStream_t s) {
  constexpr int N = Tile<RPT>::N;
  const size_t smem = (size_t)(4 * N * (a.dh + 1) + 2 * N * (N + 1)) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(mha_bwd_kernel<RPT, NDC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(mha_bwd_kernel<RPT, NDC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return fail("mha_bwd: cannot raise the shared-memory limit");
    attr = true;
  }
  launch_pdl(mha_bwd_kernel<RPT, NDC, false>, dim3(dim3(ceil_div(a.Sq, N), a.nhead, a.B)), dim3(256), smem, s, a);
  SLNLP_LAUNCH_OK("mha_bwd(dq)");
  launch_pdl(mha_bwd_kernel<RPT, NDC, true>, dim3(dim3(ceil_div(a.Sk, N), a.nhead, a.B)), dim3(256), smem, s, a);
  SLNLP_LAUNCH_OK("mha_bwd(dkv)");
  return 0;
}

static int check_mha(const MhaArgs& a, const char* who) {
  SLNLP_CHECK_ARG(a.B > 0 && a.Sq > 0 && a.Sk > 0 && a.nhead > 0, "%s: bad shape", who);
  SLNLP_CHECK_ARG(a.dh >= 4 && a.dh <= 256 && a.dh % 4 == 0, "%s: head dimension %d must be a multiple of 4 in [4, 256]", who, a.dh);
  SLNLP_CHECK_ARG(a.ldq % 4 == 0 && a.ldk % 4 == 0 && a.ldv % 4 == 0 && a.ldo % 4 == 0, "%s: leading dimensions must be multiples of 4", who);
  SLNLP_CHECK_ARG(a.p_drop >= 0.f && a.p_drop < 1.f && (a.p_drop == 0.f || a.rng), "%s: bad dropout arguments", who);
  SLNLP_CHECK_ARG(a.B <= 65535 && a.nhead <= 65535, "%s: grid too large", who);
  return 0;
}

}  // namespace slnlp

using namespace slnlp;

// head-dim columns per thread: the smallest of 1, 2, 4, 8 (64-row tiles) or 16 (32-row tiles) covering dh / 16
#define MHA_DISPATCH(FN, a, s)                          \
  if ((a).dh <= 16) return FN<4, 1>(a, s);              \
  if ((a).dh <= 32) return FN<4, 2>(a, s);              \
  if ((a).dh <= 64) return FN<4, 4>(a, s);              \
  if ((a).dh <= 128) return FN<4, 8>(a, s);             \
  return FN<2, 16>(a, s);

static int mha_fwd_entry(bool tc, const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                         float* o, int ldo, float* lse, int B, int Sq, int Sk, int nhead, int dh,
                         int causal, const int64_t* key_tokens, int64_t pad_idx, float p_drop,
                         const uint64_t* rng, uint32_t site, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(q && k && v && o && lse, "mha_fwd: null pointer");
  MhaArgs a{};
  a.q = q; a.k = k; a.v = v; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.o = o; a.ldo = ldo; a.lse = lse;
  a.B = B; a.Sq = Sq; a.Sk = Sk; a.nhead = nhead; a.dh = dh; a.causal = causal;
  a.key_tokens = key_tokens; a.pad_idx = pad_idx; a.scale = 1.f / sqrtf((float)dh);
  a.p_drop = p_drop; a.rng = rng; a.site = site;
  if (int rc = check_mha(a, "mha_fwd")) return rc;
  if (mha_small_ok(a)) return launch_mha_small<false>(a, as_stream(stream));
  if (tc && mha_tc_ok(a)) {
    if (dh == 16) return launch_mha_tc_fwd<16>(a, as_stream(stream));
    if (dh == 32) return launch_mha_tc_fwd<32>(a, as_stream(stream));
    return launch_mha_tc_fwd<64>(a, as_stream(stream));
  }
  MHA_DISPATCH(launch_mha_fwd, a, as_stream(stream));
}

static int mha_bwd_entry(bool tc, const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                         const float* o, const float* dout, int ldo, const float* lse, float* dvec,
                         float* dq, float* dk, float* dv, int B, int Sq, int Sk, int nhead, int dh,
                         int causal, const int64_t* key_tokens, int64_t pad_idx, float p_drop,
                         const uint64_t* rng, uint32_t site, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(q && k && v && o && dout && lse && dvec && dq && dk && dv, "mha_bwd: null pointer");
  MhaArgs a{};
  a.q = q; a.k = k; a.v = v; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.o = const_cast<float*>(o); a.dout = dout;
  a.ldo = ldo; a.lse = const_cast<float*>(lse); a.dvec = dvec; a.dq = dq; a.dk = dk; a.dv = dv;
  a.B = B; a.Sq = Sq; a.Sk = Sk; a.nhead = nhead; a.dh = dh; a.causal = causal;
  a.key_tokens = key_tokens; a.pad_idx = pad_idx; a.scale = 1.f / sqrtf((float)dh);
  a.p_drop = p_drop; a.rng = rng; a.site = site;
  if (int rc = check_mha(a, "mha_bwd")) return rc;
  if (mha_small_ok(a)) return launch_mha_small<true>(a, as_stream(stream));
  if (tc && mha_tc_ok(a)) {
    if (dh == 16) return launch_mha_tc_bwd<16>(a, as_stream(stream));
    if (dh == 32) return launch_mha_tc_bwd<32>(a, as_stream(stream));
    return launch_mha_tc_bwd<64>(a, as_stream(stream));
  }
  MHA_DISPATCH(launch_mha_bwd, a, as_stream(stream));
}

extern "C" int slnlp_mha_fwd(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                             float* o, int ldo, float* lse, int B, int Sq, int Sk, int nhead, int dh,
                             int causal, const int64_t* key_tokens, int64_t pad_idx, float p_drop,
                             const uint64_t* rng, uint32_t site, slnlp_stream_t stream) {
  return mha_fwd_entry(false, q, ldq, k, ldk, v, ldv, o, ldo, lse, B, Sq, Sk, nhead, dh, causal, key_tokens, pad_idx,
                       p_drop, rng, site, stream);
}
extern "C" int slnlp_mha_tf32_fwd(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                                  float* o, int ldo, float* lse, int B, int Sq, int Sk, int nhead, int dh,
                                  int causal, const int64_t* key_tokens, int64_t pad_idx, float p_drop,
                                  const uint64_t* rng, uint32_t site, slnlp_stream_t stream) {
  return mha_fwd_entry(true, q, ldq, k, ldk, v, ldv, o, ldo, lse, B, Sq, Sk, nhead, dh, causal, key_tokens, pad_idx,
                       p_drop, rng, site, stream);
}
extern "C" int slnlp_mha_bwd(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                             const float* o, const float* dout, int ldo, const float* lse, float* dvec,
                             float* dq, float* dk, float* dv, int B, int Sq, int Sk, int nhead, int dh,
                             int causal, const int64_t* key_tokens, int64_t pad_idx, float p_drop,
                             const uint64_t* rng, uint32_t site, slnlp_stream_t stream) {
  return mha_bwd_entry(false, q, ldq, k, ldk, v, ldv, o, dout, ldo, lse, dvec, dq, dk, dv, B, Sq, Sk, nhead, dh, causal,
                       key_tokens, pad_idx, p_drop, rng, site, stream);
}
extern "C" int slnlp_mha_tf32_bwd(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                                  const float* o, const float* dout, int ldo, const float* lse, float* dvec,
                                  float* dq, float* dk, float* dv, int B, int Sq, int Sk, int nhead, int dh,
                                  int causal, const int64_t* key_tokens, int64_t pad_idx, float p_drop,
                                  const uint64_t* rng, uint32_t site, slnlp_stream_t stream) {
  return mha_bwd_entry(true, q, ldq, k, ldk, v, ldv, o, dout, ldo, lse, dvec, dq, dk, dv, B, Sq, Sk, nhead, dh, causal,
                       key_tokens, pad_idx, p_drop, rng, site, stream);
}

extern "C" int slnlp_add_layernorm_fwd(const float* x, const float* res, const float* gamma, const float* beta,
                                       float* y, float* mean, float* rstd, int rows, int E, float eps,
                                       slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(x && gamma && beta && y, "add_layernorm_fwd: null pointer");
  SLNLP_CHECK_ARG(rows > 0 && E >= 1 && E <= 32 * LN_MAX_VPL, "add_layernorm_fwd: E must be in [1, 1024]");
#define SLNLP_LN_FWD(V) launch_pdl(add_layernorm_fwd_kernel<V>, dim3(ceil_div(rows, LN_WARPS)), dim3(LN_WARPS * 32), 0, as_stream(stream), x, res, gamma, beta, y, mean, rstd, rows, E, eps)
  if (E <= 128) SLNLP_LN_FWD(4);
  else if (E <= 256) SLNLP_LN_FWD(8);
  else if (E <= 512) SLNLP_LN_FWD(16);
  else SLNLP_LN_FWD(32);
#undef SLNLP_LN_FWD
  SLNLP_LAUNCH_OK("add_layernorm_fwd");
  return 0;
}

extern "C" int slnlp_ln_bwd_blocks(int rows) { return ln_bwd_blocks(rows); }

extern "C" int slnlp_layernorm_bwd(const float* dy, const float* x, const float* res, const float* gamma,
                                   const float* mean, const float* rstd, float* dx, float* partials,
                                   int rows, int E, int accumulate, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(dy && x && gamma && mean && rstd && dx && partials, "layernorm_bwd: null pointer");
  SLNLP_CHECK_ARG(rows > 0 && E >= 1 && E <= 32 * LN_MAX_VPL, "layernorm_bwd: E must be in [1, 1024]");
  const int nb = ln_bwd_blocks(rows);
#define SLNLP_LN_BWD(V) launch_pdl(layernorm_bwd_kernel<V>, dim3(nb), dim3(LN_WARPS * 32), LN_WARPS * 2 * E * sizeof(float), as_stream(stream),        dy, x, res, gamma, mean, rstd, dx, partials, rows, E, accumulate)
  if (E <= 128) SLNLP_LN_BWD(4);
  else if (E <= 256) SLNLP_LN_BWD(8);
  else if (E <= 512) SLNLP_LN_BWD(16);
  else SLNLP_LN_BWD(32);
#undef SLNLP_LN_BWD
  SLNLP_LAUNCH_OK("layernorm_bwd");
  return 0;
}
