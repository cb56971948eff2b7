// K7b: the decoder "head" of the single decode step, one kernel per direction of autograd, one CTA per sequence
// (bkp:304-327 attention, bkp:202-216 decoder input; optionally bkp:268-280 bridge and bkp:312 query).
//   default form : attention (scores, masked softmax, context) + [bos embedding || context] in one launch; backward twin =
//                  the attention backward reading d(ctx) in place from d(decoder input).
//   optional     : the bridge rows (tanh(W_b enc_final + b_b), every layer) and the query row (W_q hidden0[L-1]) of the
//                  sequence as warp-per-output dot products by the same CTA (w_bridge / w_query != NULL), and their
//                  backward (d hidden0 += dq W_q, tanh', d enc_final = d hidden0 W_b).
// Every stage is independent per sequence, which is what made the one-CTA-per-sequence form attractive: at the reference's
// batch (50) bridge GEMM -> tanh -> query GEMM -> attention -> decoder input are five dependent launches of 2-8 us, all
// launch and round-trip latency.  MEASURED (profiles/r02_dec_fused_ab.txt): folding the two products in is SLOWER than the
// launches it removes (cfg1 0.381 -> 0.398 ms/step; 16.0 / 20.0 us per launch against 7.9 / 11.3 without them): the head then
// waits for the key projection, which ran next to bridge -> tanh -> query before, and 32 warps re-reading 192 KB of weights
// per sequence and replicating the scalar work cost what two 6 us GEMM launches did.  The callers therefore pass
// w_bridge = w_query = NULL by default ($SLNLP_DEC_HEAD_FUSE=1 turns the products on; parity-tested either way).
// fp32 FMA with full-precision tanhf / expf: serves both precision paths.
#include "common.cuh"

namespace slnlp {

// U dot products at once: out[u] = sum_c g[u][c] * s[u][c], c < n; g[u] rows in global memory (weights), s[u] in shared
// memory; the U rows' loads are issued together (one L2 round trip per 128 columns instead of U); result in every lane
template <int U>
__device__ __forceinline__ void warp_dots(const float* const (&g)[U], const float* const (&s)[U], int n, int lane, bool vec,
                                          float (&out)[U]) {
  float acc[U];
#pragma unroll
  for (int u = 0; u < U; ++u) acc[u] = 0.f;
  if (vec) {
#pragma unroll 2
    for (int c = lane * 4; c < n; c += 128) {
      float4 a[U];
#pragma unroll
      for (int u = 0; u < U; ++u) a[u] = __ldg(reinterpret_cast<const float4*>(g[u] + c));
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float4 x = *reinterpret_cast<const float4*>(s[u] + c);
        acc[u] = fmaf(a[u].x, x.x, acc[u]);
        acc[u] = fmaf(a[u].y, x.y, acc[u]);
        acc[u] = fmaf(a[u].z, x.z, acc[u]);
        acc[u] = fmaf(a[u].w, x.w, acc[u]);
      }
    }
  } else {
    for (int c = lane; c < n; c += 32) {
#pragma unroll
      for (int u = 0; u < U; ++u) acc[u] = fmaf(__ldg(g[u] + c), s[u][c], acc[u]);
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) out[u] = warp_sum(acc[u]);
}
constexpr int HD_U = 4;   // rows per warp and pass

struct DecHeadFwd {
  int T, B, H, W, L, E;
  const float* enc_final;   // [L,B,W] final encoder states, directions concatenated (bridge input)
  const float* w_bridge;    // [H,W] or null
  const float* b_bridge;    // [H]
  const float* w_query;     // [H,H] or null
  const float* pk;          // [T,B,H] key projection
  const float* v;           // [H] energy weights
  const float* val;         // [T,B,W] pad-filled encoder output
  const int64_t* X;         // [B,T] tokens (mask = X != pad_idx)
  int64_t pad_idx;
  const float* bos_row;     // [E] target embedding of the decoder's first input
  float* hidden0;           // [L,B,H]: written (bridge fused) or read
  float* q;                 // [B,H]: written (query fused) or read
  float* alpha;             // [B,T]
  float* ctx;               // [B,W] or null
  float* dec_xin;           // [B,E+W] = [bos_row || ctx]
};

// grid = B, block = 256 or 1024.  Dynamic smem (floats): xs[L*W] (bridge only) | hid[H] | qs[H] | sc[Tp] | part[nw*W]
__global__ void __launch_bounds__(1024) dec_head_fwd_kernel(DecHeadFwd p) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ __align__(16) float sm[];
  __shared__ float red[33];
  const int T = p.T, B = p.B, H = p.H, W = p.W, L = p.L;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
  float* xs = sm;
  float* hid = xs + (p.w_bridge ? L * W : 0);
  float* qs = hid + H;
  float* sc = qs + H;
  float* part = sc + ((T + 3) & ~3);

  // ---- bridge: hidden0[l,b,:] = tanh(W_b enc_final[l,b,:] + b_b)   (bkp:268-280)
  if (p.w_bridge) {
    for (int i = tid; i < L * W; i += blockDim.x) xs[i] = p.enc_final[((int64_t)(i / W) * B + b) * W + i % W];
    __syncthreads();
    const bool vec = (W % 4 == 0) && (((uintptr_t)p.w_bridge & 15) == 0);
    const int NO = L * H;
    for (int o0 = w; o0 < NO; o0 += nw * HD_U) {
      const float* g[HD_U];
      const float* x[HD_U];
      float s[HD_U];
#pragma unroll
      for (int u = 0; u < HD_U; ++u) {
        const int o = min(o0 + u * nw, NO - 1);      // past the end: a valid row, result dropped
        g[u] = p.w_bridge + (int64_t)(o % H) * W;
        x[u] = xs + (o / H) * W;
      }
      warp_dots<HD_U>(g, x, W, lane, vec, s);
      if (lane == 0) {
#pragma unroll
        for (int u = 0; u < HD_U; ++u) {
          const int o = o0 + u * nw;
          if (o < NO) {
            const int l = o / H, h = o % H;
            const float y = tanhf(s[u] + p.b_bridge[h]);
            p.hidden0[((int64_t)l * B + b) * H + h] = y;
            if (l == L - 1) hid[h] = y;
          }
        }
      }
    }
  } else if (p.w_query) {
    for (int h = tid; h < H; h += blockDim.x) hid[h] = p.hidden0[((int64_t)(L - 1) * B + b) * H + h];
  }
  __syncthreads();
  // ---- query: q[b,:] = W_q hidden0[L-1,b,:]   (bkp:312)
  if (p.w_query) {
    const bool vec = (H % 4 == 0) && (((uintptr_t)p.w_query & 15) == 0);
    for (int o0 = w; o0 < H; o0 += nw * HD_U) {
      const float* g[HD_U];
      const float* x[HD_U];
      float s[HD_U];
#pragma unroll
      for (int u = 0; u < HD_U; ++u) {
        g[u] = p.w_query + (int64_t)min(o0 + u * nw, H - 1) * H;
        x[u] = hid;
      }
      warp_dots<HD_U>(g, x, H, lane, vec, s);
      if (lane == 0) {
#pragma unroll
        for (int u = 0; u < HD_U; ++u) {
          const int o = o0 + u * nw;
          if (o < H) {
            qs[o] = s[u];
            p.q[(int64_t)b * H + o] = s[u];
          }
        }
      }
    }
  } else {
    for (int h = tid; h < H; h += blockDim.x) qs[h] = p.q[(int64_t)b * H + h];
  }
  __syncthreads();
  // ---- scores, masked softmax, context: attention.cu's forward with q in shared memory
  for (int t = w; t < T; t += nw) {
    const float* k = p.pk + ((int64_t)t * B + b) * H;
    float s = 0.f;
    for (int h = lane; h < H; h += 32) s += p.v[h] * tanhf(qs[h] + k[h]);
    s = warp_sum(s);
    if (lane == 0) sc[t] = p.X[(int64_t)b * T + t] == p.pad_idx ? -INFINITY : s;
  }
  __syncthreads();
  float m = -INFINITY;
  for (int t = tid; t < T; t += blockDim.x) m = fmaxf(m, sc[t]);
  m = block_max(m, red);
  float sum = 0.f;
  for (int t = tid; t < T; t += blockDim.x) {
    const float e = expf(sc[t] - m);
    sc[t] = e;
    sum += e;
  }
  sum = block_sum(sum, red);
  const float inv = 1.f / sum;
  for (int t = tid; t < T; t += blockDim.x) {
    const float a = sc[t] * inv;
    sc[t] = a;
    p.alpha[(int64_t)b * T + t] = a;
  }
  __syncthreads();
  for (int c0 = 0; c0 < W; c0 += 32 * 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int c = c0 + lane * 4;
    if (c < W) {
      for (int t = w; t < T; t += nw) {
        const float a = sc[t];
        const float* vr = p.val + ((int64_t)t * B + b) * W + c;
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
  // ---- decoder input of the one step: [trg_embed[bos] || ctx]   (bkp:202-216)
  float* xrow = p.dec_xin + (int64_t)b * (p.E + W);
  for (int c = tid; c < W; c += blockDim.x) {
    float s = 0.f;
    for (int ww = 0; ww < nw; ++ww) s += part[ww * W + c];
    if (p.ctx) p.ctx[(int64_t)b * W + c] = s;
    xrow[p.E + c] = s;
  }
  for (int c = tid; c < p.E; c += blockDim.x) xrow[c] = p.bos_row[c];
}

struct DecHeadBwd {
  int T, B, H, W, L, E;
  const float* d_decx;      // [B,E+W]: d(decoder input); its last W columns are d(ctx)
  const float* q;           // [B,H]
  const float* pk;          // [T,B,H]
  const float* v;           // [H]
  const float* val;         // [T,B,W]
  const float* alpha;       // [B,T]
  const float* hidden0;     // [L,B,H] (bridge fused)
  const float* w_query;     // [H,H] or null
  const float* w_bridge;    // [H,W] or null
  float* dval;              // [T,B,W]
  float* dpk;               // [T,B,H]
  float* dq;                // [B,H]
  float* dv_part;           // [B,H]
  float* d_hidden0;         // [L,B,H]: in = the decoder cells' d(h0) (+ d(c0)); out: + query path (top layer), x tanh' (bridge fused)
  float* d_enc_final;       // [L,B,W] (bridge fused)
};

constexpr int HB_NGQ = 8;   // split of the h range of the query-backward product
constexpr int HB_NGB = 4;   // ... of the bridge-backward product

// grid = B, block = 256 or 1024.  Dynamic smem (floats): ds[Tp] | dqs[H] | dhs[L*H] | scratch[max(nw*2*H, 8*H, 4*L*W)]
__global__ void __launch_bounds__(1024) dec_head_bwd_kernel(DecHeadBwd p) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ __align__(16) float sm[];
  __shared__ float red[33];
  const int T = p.T, B = p.B, H = p.H, W = p.W, L = p.L;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
  float* ds = sm;
  float* dqs = ds + ((T + 3) & ~3);
  float* dhs = dqs + H;
  float* part = dhs + L * H;
  const float* dc = p.d_decx + (int64_t)b * (p.E + W) + p.E;
  const float* al = p.alpha + (int64_t)b * T;
  // ---- attention backward (attention.cu): dalpha_t = <dctx, val_t>; dval_t = alpha_t dctx
  for (int t = w; t < T; t += nw) {
    const int64_t o = ((int64_t)t * B + b) * W;
    const float a = al[t];
    float s = 0.f;
    for (int c = lane; c < W; c += 32) {
      const float g = dc[c];
      s = fmaf(g, p.val[o + c], s);
      p.dval[o + c] = a * g;
    }
    s = warp_sum(s);
    if (lane == 0) ds[t] = s;
  }
  __syncthreads();
  float dot = 0.f;
  for (int t = tid; t < T; t += blockDim.x) dot += al[t] * ds[t];
  dot = block_sum(dot, red);
  for (int t = tid; t < T; t += blockDim.x) ds[t] = al[t] * (ds[t] - dot);
  __syncthreads();
  const float* qb = p.q + (int64_t)b * H;
  for (int h = lane; h < H; h += 32) {
    const float qh = qb[h], vh = p.v[h];
    float sq = 0.f, sv = 0.f;
    for (int t = w; t < T; t += nw) {
      const int64_t o = ((int64_t)t * B + b) * H + h;
      const float u = tanhf(qh + p.pk[o]);
      const float dsc = ds[t];
      const float dp = dsc * vh * (1.f - u * u);
      p.dpk[o] = dp;
      sq += dp;
      sv = fmaf(dsc, u, sv);
    }
    part[(w * 2 + 0) * H + h] = sq;
    part[(w * 2 + 1) * H + h] = sv;
  }
  __syncthreads();
  for (int h = tid; h < H; h += blockDim.x) {
    float sq = 0.f, sv = 0.f;
    for (int ww = 0; ww < nw; ++ww) {
      sq += part[(ww * 2 + 0) * H + h];
      sv += part[(ww * 2 + 1) * H + h];
    }
    p.dq[(int64_t)b * H + h] = sq;
    p.dv_part[(int64_t)b * H + h] = sv;
    dqs[h] = sq;
  }
  if (!p.w_query) return;
  __syncthreads();
  // ---- query backward: d hidden0[L-1,b,k] += sum_h dq[h] W_q[h,k]; thread = (slice of h, k): coalesced rows of W_q
  for (int item = tid; item < HB_NGQ * H; item += blockDim.x) {
    const int g = item / H, k = item % H;
    float s = 0.f;
#pragma unroll 4
    for (int h = g; h < H; h += HB_NGQ) s = fmaf(dqs[h], __ldg(p.w_query + (int64_t)h * H + k), s);
    part[item] = s;
  }
  __syncthreads();
  // ... and tanh' of the bridge for every layer's row of this sequence
  for (int idx = tid; idx < L * H; idx += blockDim.x) {
    const int l = idx / H, k = idx % H;
    const int64_t e = ((int64_t)l * B + b) * H + k;
    float d = p.d_hidden0[e];
    if (l == L - 1) {
#pragma unroll
      for (int g = 0; g < HB_NGQ; ++g) d += part[g * H + k];
    }
    if (p.w_bridge) {
      const float y = p.hidden0[e];
      d *= 1.f - y * y;
      dhs[idx] = d;
    }
    if (l == L - 1 || p.w_bridge) p.d_hidden0[e] = d;
  }
  if (!p.w_bridge) return;
  __syncthreads();
  // ---- bridge backward: d enc_final[l,b,c] = sum_h d hidden0[l,b,h] W_b[h,c]
  const int LW = L * W;
  for (int item = tid; item < HB_NGB * LW; item += blockDim.x) {
    const int g = item / LW, o = item % LW, l = o / W, c = o % W;
    const float* dh = dhs + l * H;
    float s = 0.f;
#pragma unroll 8
    for (int h = g; h < H; h += HB_NGB) s = fmaf(dh[h], __ldg(p.w_bridge + (int64_t)h * W + c), s);
    part[item] = s;
  }
  __syncthreads();
  for (int o = tid; o < LW; o += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < HB_NGB; ++g) s += part[g * LW + o];
    p.d_enc_final[((int64_t)(o / W) * B + b) * W + o % W] = s;
  }
}

static size_t head_fwd_smem(int T, int H, int W, int L, bool bridge, int threads) {
  return ((size_t)(bridge ? L * W : 0) + 2 * H + ((T + 3) & ~3) + (size_t)(threads / 32) * W) * sizeof(float);
}
static size_t head_bwd_smem(int T, int H, int W, int L, bool query, bool bridge, int threads) {
  size_t scratch = (size_t)(threads / 32) * 2 * H;
  if (query && (size_t)HB_NGQ * H > scratch) scratch = (size_t)HB_NGQ * H;
  if (bridge && (size_t)HB_NGB * L * W > scratch) scratch = (size_t)HB_NGB * L * W;
  return ((size_t)((T + 3) & ~3) + H + (size_t)L * H + scratch) * sizeof(float);
}
constexpr size_t HEAD_SMEM_MAX = 200 * 1024;

static int head_threads(int T, int B) {
  // as attention.cu: 1024 threads shorten the dependent-load chains of one sequence while the batch leaves SMs idle
  return (T >= 32 && B < 4 * (sm_count() > 0 ? sm_count() : 148)) ? 1024 : 256;
}

}  // namespace slnlp

using namespace slnlp;

extern "C" int slnlp_dec_head_supported(int T, int B, int H, int L, int fuse_query, int fuse_bridge) {
  if (T <= 0 || B <= 0 || H <= 0 || L <= 0 || T > 12000) return 0;
  if (fuse_bridge && !fuse_query) return 0;
  const int W = 2 * H;
  return head_fwd_smem(T, H, W, L, fuse_bridge != 0, 256) <= HEAD_SMEM_MAX &&
                 head_bwd_smem(T, H, W, L, fuse_query != 0, fuse_bridge != 0, 256) <= HEAD_SMEM_MAX
             ? 1 : 0;
}

extern "C" int slnlp_dec_head_fwd(int T, int B, int H, int L, int E, const float* enc_final, const float* w_bridge,
                                  const float* b_bridge, const float* w_query, const float* pk, const float* v,
                                  const float* val, const int64_t* X, int64_t pad_idx, const float* bos_row,
                                  float* hidden0, float* q, float* alpha, float* ctx, float* dec_xin,
                                  slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(pk && v && val && X && bos_row && hidden0 && q && alpha && dec_xin, "dec_head_fwd: null pointer");
  SLNLP_CHECK_ARG(T > 0 && B > 0 && H > 0 && L > 0 && E > 0 && T <= 12000, "dec_head_fwd: bad shape");
  SLNLP_CHECK_ARG(!w_bridge || (enc_final && b_bridge && w_query), "dec_head_fwd: the fused bridge needs enc_final, b_bridge and the fused query");
  const int W = 2 * H;
  int threads = head_threads(T, B);
  if (head_fwd_smem(T, H, W, L, w_bridge != nullptr, threads) > HEAD_SMEM_MAX) threads = 256;
  const size_t smf = head_fwd_smem(T, H, W, L, w_bridge != nullptr, threads);
  SLNLP_CHECK_ARG(smf <= HEAD_SMEM_MAX, "dec_head_fwd: shape too large for shared memory (ask slnlp_dec_head_supported)");
  if (smf > 48 * 1024) cudaFuncSetAttribute(dec_head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEAD_SMEM_MAX);
  DecHeadFwd p{T, B, H, W, L, E, enc_final, w_bridge, b_bridge, w_query, pk, v, val, X, pad_idx, bos_row,
               hidden0, q, alpha, ctx, dec_xin};
  launch_pdl(dec_head_fwd_kernel, dim3(B), dim3(threads), smf, as_stream(stream), p);
  SLNLP_LAUNCH_OK("dec_head_fwd");
  return 0;
}

extern "C" int slnlp_dec_head_bwd(int T, int B, int H, int L, int E, const float* d_decx, const float* q, const float* pk,
                                  const float* v, const float* val, const float* alpha, const float* hidden0,
                                  const float* w_query, const float* w_bridge, float* dval, float* dpk, float* dq,
                                  float* dv_part, float* d_hidden0, float* d_enc_final, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(d_decx && q && pk && v && val && alpha && dval && dpk && dq && dv_part, "dec_head_bwd: null pointer");
  SLNLP_CHECK_ARG(T > 0 && B > 0 && H > 0 && L > 0 && E > 0 && T <= 12000, "dec_head_bwd: bad shape");
  SLNLP_CHECK_ARG(!w_query || d_hidden0, "dec_head_bwd: the fused query needs d_hidden0");
  SLNLP_CHECK_ARG(!w_bridge || (w_query && hidden0 && d_enc_final), "dec_head_bwd: the fused bridge needs hidden0, d_enc_final and the fused query");
  const int W = 2 * H;
  int threads = head_threads(T, B);
  if (head_bwd_smem(T, H, W, L, w_query != nullptr, w_bridge != nullptr, threads) > HEAD_SMEM_MAX) threads = 256;
  const size_t smb = head_bwd_smem(T, H, W, L, w_query != nullptr, w_bridge != nullptr, threads);
  SLNLP_CHECK_ARG(smb <= HEAD_SMEM_MAX, "dec_head_bwd: shape too large for shared memory (ask slnlp_dec_head_supported)");
  if (smb > 48 * 1024) cudaFuncSetAttribute(dec_head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEAD_SMEM_MAX);
  DecHeadBwd p{T, B, H, W, L, E, d_decx, q, pk, v, val, alpha, hidden0, w_query, w_bridge, dval, dpk, dq, dv_part,
               d_hidden0, d_enc_final};
  launch_pdl(dec_head_bwd_kernel, dim3(B), dim3(threads), smb, as_stream(stream), p);
  SLNLP_LAUNCH_OK("dec_head_bwd");
  return 0;
}
