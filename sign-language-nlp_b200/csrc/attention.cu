// K7: Bahdanau (additive) attention for the single decode step of the reference
// (BahdanauAttention.forward, bkp:304-327): score, masked softmax and context fused
// in one kernel per direction of autograd; warp-shuffle reductions, one CTA per
// sequence.  HBM-bound: reads proj_key [T,B,H] and enc_out [T,B,W] exactly once.
#include "common.cuh"

namespace slnlp {

// grid = B, block = 256 or 1024 (one warp per 2-8 timesteps: every phase is a short chain of
// dependent global loads, so more warps = fewer serial round trips).  Dynamic smem: T floats
// (scores -> alphas) + [warps][W] partial contexts.
__global__ void __launch_bounds__(1024) attn_fwd_kernel(const float* __restrict__ q,
                                                       const float* __restrict__ pk,
                                                       const float* __restrict__ v,
                                                       const float* __restrict__ val,
                                                       const int64_t* __restrict__ X, int64_t pad_idx,
                                                       int T, int B, int H, int W,
                                                       float* __restrict__ alpha, float* __restrict__ ctx) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ float sc[];
  __shared__ float red[33];
  const int b = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* qb = q + (int64_t)b * H;
  // e_t = sum_h v_h tanh(q_h + pk[t,b,h]); one warp per t
  for (int t = w; t < T; t += nw) {
    const float* k = pk + ((int64_t)t * B + b) * H;
    float s = 0.f;
    for (int h = lane; h < H; h += 32) s += v[h] * tanhf(qb[h] + k[h]);
    s = warp_sum(s);
    if (lane == 0) sc[t] = X[(int64_t)b * T + t] == pad_idx ? -INFINITY : s;
  }
  __syncthreads();
  float m = -INFINITY;
  for (int t = threadIdx.x; t < T; t += blockDim.x) m = fmaxf(m, sc[t]);
  m = block_max(m, red);
  float sum = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float e = expf(sc[t] - m);  // every position masked: -inf - -inf = NaN, as torch
    sc[t] = e;
    sum += e;
  }
  sum = block_sum(sum, red);
  const float inv = 1.f / sum;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float a = sc[t] * inv;
    sc[t] = a;
    alpha[(int64_t)b * T + t] = a;
  }
  __syncthreads();
  // ctx[b,:] = sum_t alpha_t val[t,b,:]: warp w takes t = w, w+nw, ... (T/nw independent row loads in
  // flight per warp instead of one T-long dependent chain per thread), partials meet in shared memory
  float* part = sc + T;                       // [nw][W]
  for (int c0 = 0; c0 < W; c0 += 32 * 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int c = c0 + lane * 4;
    if (c < W) {
      for (int t = w; t < T; t += nw) {
        const float a = sc[t];
        const float* vr = val + ((int64_t)t * B + b) * W + c;
        if (c + 3 < W && (((uintptr_t)vr & 15) == 0)) {
          const float4 x = *reinterpret_cast<const float4*>(vr);
          acc[0] = fmaf(a, x.x, acc[0]); acc[1] = fmaf(a, x.y, acc[1]);
          acc[2] = fmaf(a, x.z, acc[2]); acc[3] = fmaf(a, x.w, acc[3]);
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (c + u < W) acc[u] = fmaf(a, vr[u], acc[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (c + u < W) part[w * W + c + u] = acc[u];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < W; c += blockDim.x) {
    float s = 0.f;
    for (int ww = 0; ww < nw; ++ww) s += part[ww * W + c];
    ctx[(int64_t)b * W + c] = s;
  }
}

// grid = B, block = 256 or 1024.  Dynamic smem: T floats (dscore) + [warps][2][H] partial sums.
__global__ void __launch_bounds__(1024) attn_bwd_kernel(const float* __restrict__ dctx,
                                                       const float* __restrict__ q,
                                                       const float* __restrict__ pk,
                                                       const float* __restrict__ v,
                                                       const float* __restrict__ val,
                                                       const float* __restrict__ alpha, int T, int B,
                                                       int H, int W, float* __restrict__ dval,
                                                       float* __restrict__ dpk, float* __restrict__ dq,
                                                       float* __restrict__ dv_part) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ float ds[];
  __shared__ float red[33];
  const int b = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* dc = dctx + (int64_t)b * W;
  const float* al = alpha + (int64_t)b * T;
  // dalpha_t = <dctx, val_t>;  dval_t = alpha_t * dctx
  for (int t = w; t < T; t += nw) {
    const int64_t o = ((int64_t)t * B + b) * W;
    const float a = al[t];
    float s = 0.f;
    for (int c = lane; c < W; c += 32) {
      const float g = dc[c];
      s = fmaf(g, val[o + c], s);
      dval[o + c] = a * g;
    }
    s = warp_sum(s);
    if (lane == 0) ds[t] = s;
  }
  __syncthreads();
  float dot = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) dot += al[t] * ds[t];
  dot = block_sum(dot, red);
  for (int t = threadIdx.x; t < T; t += blockDim.x) ds[t] = al[t] * (ds[t] - dot);  // dscore
  __syncthreads();
  // u = tanh(q + pk); dpre = dscore * v * (1-u^2); dpk = dpre; dq = sum_t dpre; dv = sum_t dscore*u.
  // Warp w takes t = w, w+nw, ...; the two sums over t meet in shared memory.
  const float* qb = q + (int64_t)b * H;
  float* part = ds + T;                       // [nw][2][H]
  for (int h = lane; h < H; h += 32) {
    const float qh = qb[h], vh = v[h];
    float sq = 0.f, sv = 0.f;
    for (int t = w; t < T; t += nw) {
      const int64_t o = ((int64_t)t * B + b) * H + h;
      const float u = tanhf(qh + pk[o]);
      const float dsc = ds[t];
      const float dp = dsc * vh * (1.f - u * u);
      dpk[o] = dp;
      sq += dp;
      sv = fmaf(dsc, u, sv);
    }
    part[(w * 2 + 0) * H + h] = sq;
    part[(w * 2 + 1) * H + h] = sv;
  }
  __syncthreads();
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    float sq = 0.f, sv = 0.f;
    for (int ww = 0; ww < nw; ++ww) {
      sq += part[(ww * 2 + 0) * H + h];
      sv += part[(ww * 2 + 1) * H + h];
    }
    dq[(int64_t)b * H + h] = sq;
    dv_part[(int64_t)b * H + h] = sv;
  }
}

}  // namespace slnlp

using namespace slnlp;

extern "C" int slnlp_attn_step_fwd(const float* q, const float* pk, const float* v, const float* val,
                                   const int64_t* X, int64_t pad_idx, int T, int B, int H, int W,
                                   float* alpha, float* ctx, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(q && pk && v && val && X && alpha && ctx, "attn_step_fwd: null pointer");
  SLNLP_CHECK_ARG(T > 0 && B > 0 && H > 0 && W > 0 && T <= 12000, "attn_step_fwd: bad shape");
  // 1024 threads shorten the dependent-load chains of ONE sequence (the reference's batch: 50 CTAs on 148 SMs); once the
  // batch fills the GPU several times over, 256-thread CTAs keep more sequences resident (measured at batch 4096:
  // 38 % -> of the HBM peak with 1024 threads, 61 % with 256)
  int threads = (T >= 32 && B < 4 * (sm_count() > 0 ? sm_count() : 148)) ? 1024 : 256;
  if ((size_t)(T + (threads / 32) * W) * sizeof(float) > 200 * 1024) threads = 256;
  const size_t smf = (size_t)(T + (threads / 32) * W) * sizeof(float);
  SLNLP_CHECK_ARG(smf <= 200 * 1024, "attn_step_fwd: T + 8*W too large for shared memory");
  if (smf > 48 * 1024) cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  launch_pdl(attn_fwd_kernel, dim3(B), dim3(threads), smf, as_stream(stream), q, pk, v, val, X, pad_idx, T, B, H, W, alpha, ctx);
  SLNLP_LAUNCH_OK("attn_step_fwd");
  return 0;
}

extern "C" int slnlp_attn_step_bwd(const float* dctx, const float* q, const float* pk, const float* v,
                                   const float* val, const float* alpha, int T, int B, int H, int W,
                                   float* dval, float* dpk, float* dq, float* dv_part,
                                   slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(dctx && q && pk && v && val && alpha && dval && dpk && dq && dv_part, "attn_step_bwd: null pointer");
  SLNLP_CHECK_ARG(T > 0 && B > 0 && H > 0 && W > 0 && T <= 12000, "attn_step_bwd: bad shape");
  int threads = (T >= 32 && B < 4 * (sm_count() > 0 ? sm_count() : 148)) ? 1024 : 256;
  if ((size_t)(T + (threads / 32) * 2 * H) * sizeof(float) > 200 * 1024) threads = 256;
  const size_t smb = (size_t)(T + (threads / 32) * 2 * H) * sizeof(float);
  SLNLP_CHECK_ARG(smb <= 200 * 1024, "attn_step_bwd: T + 16*H too large for shared memory");
  if (smb > 48 * 1024) cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  launch_pdl(attn_bwd_kernel, dim3(B), dim3(threads), smb, as_stream(stream), dctx, q, pk, v, val, alpha, T, B, H, W, dval, dpk, dq, dv_part);
  SLNLP_LAUNCH_OK("attn_step_bwd");
  return 0;
}
