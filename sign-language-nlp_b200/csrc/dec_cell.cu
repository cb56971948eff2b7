// K8: one decoder step of one RNN layer as ONE kernel (the reference decodes exactly one target position,
// MAX_OUTPUT_LEN = 1: bkp:215-216,332).  Replaces, per decoder layer, the chain
//     input-projection GEMM (x W_ih^T + b_ih) -> single-step cell kernel (h0 W_hh^T + gates) -> dropout
// of three dependent launches (5.8 + 7.3 + 3.2 us at the reference's batch of 50, all of it launch and
// round-trip latency: the step is 26 MFLOP) with a skinny fp32-FMA product over K = [x | h0] whose
// epilogue is the cell update and, between layers, the inter-layer dropout of nn.LSTM / nn.GRU.
//
// Tiling: a CTA owns DJ hidden units (all G gates of them: the cell update is local) and DB batch rows;
// 128 threads = DJ x DB, one (unit, row) pair each, G accumulators (GRU: r, z and the x / h halves of n kept
// apart, nn.GRU applies r to the h half only).  K is walked in chunks of DK over the two segments
// (x with W_ih, then h0 with W_hh); the next chunk's global loads are issued before the current chunk's
// FMAs (register staging), so that each chunk costs one L2 round trip overlapped with the math.
// Full-precision expf / tanhf: this kernel serves the fp32 (1e-5) path and the tensor-core path alike.
#include "common.cuh"

namespace slnlp {

constexpr int DJ = 4;      // hidden units per CTA
constexpr int DB = 32;     // batch rows per CTA
constexpr int DK = 128;    // K chunk
constexpr int DLD = DK + 4;

struct DecCellFwd {
  int B, H, D;
  const float* x;        // [B, D]
  const float* h0;       // [B, H]
  const float* c0;       // [B, H] (LSTM) or null
  const float* w_ih;     // [G*H, D]
  const float* w_hh;     // [G*H, H]
  const float* b_ih;     // [G*H]
  const float* b_hh;     // [G*H]
  float* gates;          // [B, G, H] activated gates (BPTT stash)
  float* stash;          // [B, H]: LSTM c_1; GRU W_hn h0 + b_hn
  float* h;              // [B, H]
  float* h_drop;         // [B, H] dropout(h) for the next layer, or null
  float p_drop;
  const uint64_t* rng;   // {seed, step}
  uint32_t site;
};

template <int G>
__global__ void __launch_bounds__(DJ * DB) dec_cell_fwd_kernel(DecCellFwd p) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ __align__(16) float Ws[G * DJ][DLD];
  __shared__ __align__(16) float Xs[DB][DLD];
  const int H = p.H, B = p.B;
  const int tid = threadIdx.x, tj = tid & (DJ - 1), tb = tid / DJ;
  const int j0 = blockIdx.x * DJ, b0 = blockIdx.y * DB;

  // chunk c of the concatenated K range: segment 0 = (x, W_ih), segment 1 = (h0, W_hh)
  const int nc0 = (p.D + DK - 1) / DK, nc1 = (H + DK - 1) / DK;
  constexpr int NW = G * DJ * (DK / 4) / (DJ * DB), NX = DB * (DK / 4) / (DJ * DB);
  float4 wv[NW], xv[NX];
  auto fetch = [&](int c) {
    const bool s1 = c >= nc0;
    const int k0 = (s1 ? c - nc0 : c) * DK, K = s1 ? H : p.D;
    const float* Wm = s1 ? p.w_hh : p.w_ih;
    const float* Xm = s1 ? p.h0 : p.x;
    const bool vec = (K % 4 == 0) && ((((uintptr_t)Wm | (uintptr_t)Xm) & 15) == 0);
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      const int e = tid + i * (DJ * DB), k4 = e % (DK / 4), row = e / (DK / 4), g = row / DJ, j = j0 + row % DJ;
      const int k = k0 + k4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < H && k < K) {
        const float* src = Wm + ((int64_t)g * H + j) * K + k;
        if (vec) v = __ldg(reinterpret_cast<const float4*>(src));
        else {
          v.x = __ldg(src);
          if (k + 1 < K) v.y = __ldg(src + 1);
          if (k + 2 < K) v.z = __ldg(src + 2);
          if (k + 3 < K) v.w = __ldg(src + 3);
        }
      }
      wv[i] = v;
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      const int e = tid + i * (DJ * DB), k4 = e % (DK / 4), bb = e / (DK / 4), b = b0 + bb;
      const int k = k0 + k4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b < B && k < K) {
        const float* src = Xm + (int64_t)b * K + k;
        if (vec) v = *reinterpret_cast<const float4*>(src);
        else {
          v.x = src[0];
          if (k + 1 < K) v.y = src[1];
          if (k + 2 < K) v.z = src[2];
          if (k + 3 < K) v.w = src[3];
        }
      }
      xv[i] = v;
    }
  };
  auto stage = [&]() {
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      const int e = tid + i * (DJ * DB), k4 = e % (DK / 4), row = e / (DK / 4);
      *reinterpret_cast<float4*>(&Ws[row][k4 * 4]) = wv[i];
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) {
      const int e = tid + i * (DJ * DB), k4 = e % (DK / 4), bb = e / (DK / 4);
      *reinterpret_cast<float4*>(&Xs[bb][k4 * 4]) = xv[i];
    }
  };

  float acc[G];
#pragma unroll
  for (int g = 0; g < G; ++g) acc[g] = 0.f;
  float xn = 0.f;       // GRU: the x half of the candidate gate
  const int nc = nc0 + nc1;
  fetch(0);
  for (int c = 0; c < nc; ++c) {
    if (c) __syncthreads();
    stage();
    __syncthreads();
    if (c + 1 < nc) fetch(c + 1);
    if (G == 3 && c == nc0) {   // the h segment starts: keep the two halves of gate n apart
      xn = acc[2];
      acc[2] = 0.f;
    }
#pragma unroll 8
    for (int kk = 0; kk < DK; kk += 4) {
      const float4 h4 = *reinterpret_cast<const float4*>(&Xs[tb][kk]);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float4 w4 = *reinterpret_cast<const float4*>(&Ws[g * DJ + tj][kk]);
        acc[g] = fmaf(w4.x, h4.x, acc[g]);
        acc[g] = fmaf(w4.y, h4.y, acc[g]);
        acc[g] = fmaf(w4.z, h4.z, acc[g]);
        acc[g] = fmaf(w4.w, h4.w, acc[g]);
      }
    }
  }

  const int j = j0 + tj, b = b0 + tb;
  if (j >= H || b >= B) return;
  float* gt = p.gates + (int64_t)b * G * H;
  float h;
  if (G == 4) {
    const float gi = sigmoidf_(acc[0] + p.b_ih[j] + p.b_hh[j]);
    const float gf = sigmoidf_(acc[1] + p.b_ih[H + j] + p.b_hh[H + j]);
    const float gg = tanhf(acc[2] + p.b_ih[2 * H + j] + p.b_hh[2 * H + j]);
    const float go = sigmoidf_(acc[G - 1] + p.b_ih[3 * H + j] + p.b_hh[3 * H + j]);
    const float c = gf * (p.c0 ? p.c0[(int64_t)b * H + j] : 0.f) + gi * gg;
    h = go * tanhf(c);
    gt[j] = gi; gt[H + j] = gf; gt[2 * H + j] = gg; gt[3 * H + j] = go;
    p.stash[(int64_t)b * H + j] = c;
  } else {
    const float hn = acc[2] + p.b_hh[2 * H + j];
    const float gr = sigmoidf_(acc[0] + p.b_ih[j] + p.b_hh[j]);
    const float gz = sigmoidf_(acc[1] + p.b_ih[H + j] + p.b_hh[H + j]);
    const float gn = tanhf(xn + p.b_ih[2 * H + j] + gr * hn);
    h = (1.f - gz) * gn + gz * p.h0[(int64_t)b * H + j];
    gt[j] = gr; gt[H + j] = gz; gt[2 * H + j] = gn;
    p.stash[(int64_t)b * H + j] = hn;
  }
  const int64_t e = (int64_t)b * H + j;
  p.h[e] = h;
  if (p.h_drop) {
    // the mask of slnlp_dropout(site) for element e: Philox block e / 4, lane e % 4
    float u[4];
    philox_uniform4(p.rng[0], p.rng[1], p.site, (uint64_t)(e >> 2), u);
    const float uu = (e & 3) == 0 ? u[0] : (e & 3) == 1 ? u[1] : (e & 3) == 2 ? u[2] : u[3];
    p.h_drop[e] = uu < 1.f - p.p_drop ? h * (1.f / (1.f - p.p_drop)) : 0.f;
  }
}


// ---------------------------------------------------------------------------------------------------------
// Backward of the same single step as ONE kernel per layer.  Replaces the chain
//     gate-gradient step kernel -> final-state step kernel (d h0 = dG W_hh) -> axpy (LSTM: c0 = h0, bkp:278-279)
//     -> d(input) GEMM (dG W_ih) -> split-K reduce -> dropout of the gradient
// (five to six dependent launches of 3-6 us).  The output is the row-concatenation [dx | dh0] = dG [W_ih | W_hh]:
// a CTA owns CB_NC of its D + H columns and CB_BB batch rows, RECOMPUTES the element-wise gate gradients of its
// rows (7 loads and one tanhf per unit: cheaper than a grid-wide exchange) into shared memory next to its
// column slice of the weights, and contracts over K = G*H in chunks of CB_KC.  256 threads = 4 slices of the
// chunk x (8 row groups x 8 column pairs): 4 x 2 outputs per thread (rows rq, rq + 8, ..: conflict-free) from float4 shared-memory reads, slices met
// in shared memory.  The CTAs of column block 0 also write dG (x side; GRU: and d(W_hn h0 + b_hn)) for the
// weight-gradient GEMMs - to buffers of their own, the activated gates stay intact for the other CTAs.
constexpr int CB_NC = 16;
constexpr int CB_BB = 32;
constexpr int CB_KC = 512;
constexpr int CB_LD = CB_KC + 4;
constexpr int CB_THREADS = 256;

struct DecCellBwd {
  int B, H, D;
  const float* gates;    // [B,G,H] activated gates
  const float* stash;    // [B,H]: LSTM c_1; GRU W_hn h0 + b_hn
  const float* h0;       // [B,H]
  const float* c0;       // [B,H] (LSTM) or null
  const float* dh;       // [B,H] d(h_1)
  const float* w_ih;     // [G*H,D]
  const float* w_hh;     // [G*H,H]
  float* dgx;            // [B,G,H] d(x-side pre-activations)
  float* dnh;            // [B,H] GRU: d(W_hn h0 + b_hn); LSTM: null
  float* dx;             // [B,D] d(input) (x dropout keep/scale of `site` when rng)
  float* dh0;            // [B,H] d(h0) (LSTM: + d(c0))
  float p_drop;
  const uint64_t* rng;
  uint32_t site;
};

// the forward's values of unit j of sequence b that its gate gradients need (loaded apart from the math: a thread
// issues the loads of several units before it uses any of them)
template <int G>
struct CellIn {
  float g[G];        // activated gates
  float st, hp, dh;  // stash (LSTM c_1; GRU W_hn h0 + b_hn), c0 (LSTM) / h0 (GRU), d(h_1)
};
template <int G>
__device__ __forceinline__ void dec_cell_load(const DecCellBwd& p, int b, int j, CellIn<G>& in) {
  const int H = p.H;
  const float* gt = p.gates + (int64_t)b * G * H;
  const int64_t e = (int64_t)b * H + j;
#pragma unroll
  for (int g = 0; g < G; ++g) in.g[g] = gt[g * H + j];
  in.st = p.stash[e];
  in.dh = p.dh[e];
  in.hp = G == 4 ? (p.c0 ? p.c0[e] : 0.f) : p.h0[e];
}
// gx: x-side d(pre-activations); nh: the h-side candidate gate (GRU; LSTM: = gx[2]); extra: the direct path into d(h0)
// (LSTM: d c0 = dc f, c0 aliases h0; GRU: dh z).  All zero for dh = 0 (rows past the batch).
template <int G>
__device__ __forceinline__ void dec_cell_math(const CellIn<G>& in, float (&gx)[G], float& nh, float& extra) {
  const float dh = in.dh;
  if (G == 4) {
    const float gi = in.g[0], gf = in.g[1], gg = in.g[2], go = in.g[G - 1];
    const float tc = tanhf(in.st);
    const float dc = dh * go * (1.f - tc * tc);
    gx[0] = dc * gg * gi * (1.f - gi);
    gx[1] = dc * in.hp * gf * (1.f - gf);
    gx[2] = dc * gi * (1.f - gg * gg);
    gx[G - 1] = dh * tc * go * (1.f - go);
    nh = gx[2];
    extra = dc * gf;
  } else {
    const float gr = in.g[0], gz = in.g[1], gn = in.g[2];
    const float da_n = dh * (1.f - gz) * (1.f - gn * gn);
    gx[0] = da_n * in.st * gr * (1.f - gr);
    gx[1] = dh * (in.hp - gn) * gz * (1.f - gz);
    gx[2] = da_n;
    nh = da_n * gr;
    extra = dh * gz;
  }
}
constexpr int CB_PU = 8;    // units whose loads a thread keeps in flight

template <int G>
__global__ void __launch_bounds__(CB_THREADS) dec_cell_bwd_kernel(DecCellBwd p) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ __align__(16) float smem[];
  float* Dsm = smem;                          // [CB_BB][CB_LD] gate gradients of this CTA's rows, one K chunk
  float* Wsm = Dsm + CB_BB * CB_LD;           // [CB_NC][CB_LD] this CTA's weight columns, K-major
  float* red = Wsm + CB_NC * CB_LD;           // [4][CB_BB*CB_NC]
  float* Esm = red + 4 * CB_BB * CB_NC;       // [CB_BB][CB_NC] the direct path into d(h0) of this CTA's columns
  const int H = p.H, B = p.B, D = p.D, GH = G * p.H;
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * CB_NC, b0 = blockIdx.y * CB_BB;
  const bool hside = n0 >= D;                 // columns of d(h0): contract with W_hh and the h-side candidate gate
  const float* Wm = hside ? p.w_hh : p.w_ih;
  const int ldw = hside ? H : D, nc0 = hside ? n0 - D : n0;
  const bool writer = blockIdx.x == 0;
  const bool wvec = (((uintptr_t)Wm) & 15) == 0;   // rows and column blocks are multiples of 16 floats by construction
  const int ks = tid >> 6, rq = tid & 7, cp = (tid & 63) >> 3;

  float acc[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = 0.f;

  for (int kc = 0; kc < GH; kc += CB_KC) {
    const int kn = min(CB_KC, GH - kc);
    if (kc) __syncthreads();
    // this CTA's weight columns of the chunk: eight 16-byte loads per thread, all issued before the first is used
    constexpr int NWV = CB_KC * CB_NC / 4 / CB_THREADS;
    float4 wr[NWV];
    if (wvec) {
#pragma unroll
      for (int i = 0; i < NWV; ++i) {
        const int e4 = tid + i * CB_THREADS, c4 = e4 & 3, kk = e4 >> 2;
        wr[i] = kk < kn ? __ldg(reinterpret_cast<const float4*>(Wm + (int64_t)(kc + kk) * ldw + nc0) + c4)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
      for (int e = tid; e < CB_KC * CB_NC; e += CB_THREADS) {
        const int c = e % CB_NC, kk = e / CB_NC;
        Wsm[c * CB_LD + kk] = (kk < kn && nc0 + c < ldw) ? __ldg(Wm + (int64_t)(kc + kk) * ldw + nc0 + c) : 0.f;
      }
    }
    bool w_pending = wvec;
    auto store_w = [&]() {
#pragma unroll
      for (int i = 0; i < NWV; ++i) {
        const int e4 = tid + i * CB_THREADS, c4 = e4 & 3, kk = e4 >> 2;
        Wsm[(c4 * 4 + 0) * CB_LD + kk] = wr[i].x;
        Wsm[(c4 * 4 + 1) * CB_LD + kk] = wr[i].y;
        Wsm[(c4 * 4 + 2) * CB_LD + kk] = wr[i].z;
        Wsm[(c4 * 4 + 3) * CB_LD + kk] = wr[i].w;
      }
      w_pending = false;
    };
    if (kn < CB_KC) {
      const int nz = CB_KC - kn;
      for (int e = tid; e < CB_BB * nz; e += CB_THREADS) Dsm[(e / nz) * CB_LD + kn + e % nz] = 0.f;
    }
    // gate gradients of this CTA's rows: the loads of CB_PU units in flight per thread
    for (int pr0 = tid; pr0 < CB_BB * H; pr0 += CB_PU * CB_THREADS) {
      CellIn<G> in[CB_PU];
#pragma unroll
      for (int u = 0; u < CB_PU; ++u) {
        const int pr = pr0 + u * CB_THREADS, bb = pr / H, j = pr - bb * H, b = b0 + bb;
        if (pr < CB_BB * H && b < B) {
          dec_cell_load<G>(p, b, j, in[u]);
        } else {
#pragma unroll
          for (int g = 0; g < G; ++g) in[u].g[g] = 0.f;
          in[u].st = in[u].hp = in[u].dh = 0.f;
        }
      }
      if (w_pending) store_w();
#pragma unroll
      for (int u = 0; u < CB_PU; ++u) {
        const int pr = pr0 + u * CB_THREADS, bb = pr / H, j = pr - bb * H, b = b0 + bb;
        if (pr >= CB_BB * H) continue;
        float gx[G], nh, extra;
        dec_cell_math<G>(in[u], gx, nh, extra);
        if (writer && kc == 0 && b < B) {
#pragma unroll
          for (int g = 0; g < G; ++g) p.dgx[((int64_t)b * G + g) * H + j] = gx[g];
          if (G == 3 && p.dnh) p.dnh[(int64_t)b * H + j] = nh;
        }
        if (hside) {
          gx[2] = nh;
          if (j >= nc0 && j < nc0 + CB_NC) Esm[bb * CB_NC + j - nc0] = extra;
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const int k = g * H + j - kc;
          if (k >= 0 && k < CB_KC) Dsm[bb * CB_LD + k] = gx[g];
        }
      }
    }
    if (w_pending) store_w();
    __syncthreads();
    const int kbeg = ks * (CB_KC / 4);
#pragma unroll 4
    for (int kk = kbeg; kk < kbeg + CB_KC / 4; kk += 4) {
      float4 d[4], wv[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) d[i] = *reinterpret_cast<const float4*>(&Dsm[(i * 8 + rq) * CB_LD + kk]);
#pragma unroll
      for (int c = 0; c < 2; ++c) wv[c] = *reinterpret_cast<const float4*>(&Wsm[(cp * 2 + c) * CB_LD + kk]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          acc[i][c] = fmaf(d[i].x, wv[c].x, acc[i][c]);
          acc[i][c] = fmaf(d[i].y, wv[c].y, acc[i][c]);
          acc[i][c] = fmaf(d[i].z, wv[c].z, acc[i][c]);
          acc[i][c] = fmaf(d[i].w, wv[c].w, acc[i][c]);
        }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 2; ++c) red[ks * (CB_BB * CB_NC) + (i * 8 + rq) * CB_NC + cp * 2 + c] = acc[i][c];
  __syncthreads();
  for (int o = tid; o < CB_BB * CB_NC; o += CB_THREADS) {
    const int bb = o / CB_NC, c = o % CB_NC, b = b0 + bb;
    if (b >= B || nc0 + c >= ldw) continue;
    float s = red[o] + red[CB_BB * CB_NC + o] + red[2 * CB_BB * CB_NC + o] + red[3 * CB_BB * CB_NC + o];
    if (hside) {
      p.dh0[(int64_t)b * H + nc0 + c] = s + Esm[bb * CB_NC + c];
    } else {
      const int64_t e = (int64_t)b * D + nc0 + c;
      if (p.rng) {   // the input was dropout(h of the layer below): the keep / scale factor slnlp_dropout(site) drew for e
        float u[4];
        philox_uniform4(p.rng[0], p.rng[1], p.site, (uint64_t)(e >> 2), u);
        const float uu = (e & 3) == 0 ? u[0] : (e & 3) == 1 ? u[1] : (e & 3) == 2 ? u[2] : u[3];
        s = uu < 1.f - p.p_drop ? s * (1.f / (1.f - p.p_drop)) : 0.f;
      }
      p.dx[e] = s;
    }
  }
}

}  // namespace slnlp

extern "C" int slnlp_dec_cell_fwd(int mode, int B, int H, int D, const float* x, const float* h0, const float* c0,
                                  const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                                  float* gates, float* stash, float* h, float* h_drop, float p_drop,
                                  const uint64_t* rng, uint32_t site, slnlp_stream_t stream) {
  using namespace slnlp;
  SLNLP_CHECK_ARG(B > 0 && H > 0 && D > 0 && x && h0 && w_ih && w_hh && b_ih && b_hh && gates && stash && h,
                  "dec_cell_fwd: bad arguments");
  SLNLP_CHECK_ARG(mode == SLNLP_MODE_LSTM || mode == SLNLP_MODE_GRU, "dec_cell_fwd: bad mode");
  SLNLP_CHECK_ARG(!h_drop || (rng && p_drop >= 0.f && p_drop < 1.f), "dec_cell_fwd: dropout needs rng and 0 <= p < 1");
  DecCellFwd p{B, H, D, x, h0, c0, w_ih, w_hh, b_ih, b_hh, gates, stash, h, h_drop, p_drop, rng, site};
  const dim3 grid(ceil_div(H, DJ), ceil_div(B, DB));
  if (mode == SLNLP_MODE_LSTM) launch_pdl(dec_cell_fwd_kernel<4>, grid, dim3(DJ * DB), 0, as_stream(stream), p);
  else launch_pdl(dec_cell_fwd_kernel<3>, grid, dim3(DJ * DB), 0, as_stream(stream), p);
  SLNLP_LAUNCH_OK("dec_cell_fwd");
  return 0;
}

extern "C" int slnlp_dec_cell_bwd_supported(int mode, int B, int H, int D) {
  (void)mode;
  return (B > 0 && H > 0 && D > 0 && H % slnlp::CB_NC == 0 && D % slnlp::CB_NC == 0) ? 1 : 0;
}

extern "C" int slnlp_dec_cell_bwd(int mode, int B, int H, int D, const float* gates, const float* stash, const float* h0,
                                  const float* c0, const float* dh, const float* w_ih, const float* w_hh, float* dgx,
                                  float* dnh, float* dx, float* dh0, float p_drop, const uint64_t* rng, uint32_t site,
                                  slnlp_stream_t stream) {
  using namespace slnlp;
  SLNLP_CHECK_ARG(mode == SLNLP_MODE_LSTM || mode == SLNLP_MODE_GRU, "dec_cell_bwd: bad mode");
  SLNLP_CHECK_ARG(B > 0 && H > 0 && D > 0 && gates && stash && h0 && dh && w_ih && w_hh && dgx && dx && dh0,
                  "dec_cell_bwd: bad arguments");
  SLNLP_CHECK_ARG(mode == SLNLP_MODE_LSTM || dnh, "dec_cell_bwd: the GRU needs dnh");
  SLNLP_CHECK_ARG(H % CB_NC == 0 && D % CB_NC == 0, "dec_cell_bwd: H and D must be multiples of %d (ask slnlp_dec_cell_bwd_supported)", CB_NC);
  SLNLP_CHECK_ARG(!rng || (p_drop >= 0.f && p_drop < 1.f), "dec_cell_bwd: dropout needs 0 <= p < 1");
  DecCellBwd p{B, H, D, gates, stash, h0, c0, dh, w_ih, w_hh, dgx, dnh, dx, dh0, p_drop, rng, site};
  const dim3 grid((D + H) / CB_NC, ceil_div(B, CB_BB));
  const size_t sm = (size_t)((CB_BB + CB_NC) * CB_LD + 5 * CB_BB * CB_NC) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(dec_cell_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaFuncSetAttribute(dec_cell_bwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    attr = true;
  }
  if (mode == SLNLP_MODE_LSTM) launch_pdl(dec_cell_bwd_kernel<4>, grid, dim3(CB_THREADS), sm, as_stream(stream), p);
  else launch_pdl(dec_cell_bwd_kernel<3>, grid, dim3(CB_THREADS), sm, as_stream(stream), p);
  SLNLP_LAUNCH_OK("dec_cell_bwd");
  return 0;
}
