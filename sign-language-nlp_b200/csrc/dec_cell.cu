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
