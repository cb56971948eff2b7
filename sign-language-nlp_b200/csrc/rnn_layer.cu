// K4/K8: recurrent layer, general fp32 path.  One launch per timestep (both
// directions in the same grid) of a fused "h_{t-1} W_hh^T + gate nonlinearities +
// cell update + length freeze" kernel, and its BPTT twin "dG_{t+1} W_hh + cell
// backward".  The host entry points loop over T on the caller's stream so that
// a whole layer is one C call and is capturable into a CUDA graph.
//
// Semantics follow torch.nn.LSTM / nn.GRU on a packed sequence as the reference
// uses them (bkp:95-100,110-123): gate order i,f,g,o / r,z,n; state frozen and
// output 0 for t >= len_b; direction 1 walks t = len_b-1 .. 0.
//
// This is the shape-general path (any H, B); the W_hh-resident persistent
// tcgen05 kernel in rnn_persistent.cu takes over for the shapes it supports.
#include <stdlib.h>
#include "common.cuh"

namespace slnlp {

constexpr int FJ = 8;     // hidden units per CTA (forward)
constexpr int FTB = 16;   // batch groups per CTA (forward); batches per CTA = FTB*RB
constexpr int FKC = 128;  // K chunk (forward)

struct StepFwd {
  int T, B, H, ndir, step;
  float* gates;         // [T,B,ndir,G,H]
  const float* w_hh;    // [ndir,G*H,H]
  const float* b_hh;    // [ndir,G*H]
  const int64_t* lengths;
  const float* h0;      // [ndir,B,H] or null
  const float* c0;
  float* out;           // [T,B,ndir*H]
  float* stash;         // [T,B,ndir,H]
  float* h_final;       // [ndir,B,H] or null
};

template <int G, int RB>
__global__ void __launch_bounds__(FJ * FTB) rnn_step_fwd_kernel(StepFwd p) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ __align__(16) float smem[];
  constexpr int BB = FTB * RB;
  constexpr int LDS_ = FKC + 4;
  float* Ws = smem;                    // [G*FJ][LDS_]
  float* Hs = smem + G * FJ * LDS_;    // [BB][LDS_]
  const int H = p.H, B = p.B, T = p.T;
  const int d = blockIdx.z;
  const int t = d == 0 ? p.step : T - 1 - p.step;
  const int tp = d == 0 ? t - 1 : t + 1;
  const bool has_prev = tp >= 0 && tp < T;
  const int j0 = blockIdx.x * FJ, b0 = blockIdx.y * BB;
  const int tid = threadIdx.x, tj = tid & (FJ - 1), tb = tid / FJ;
  const float* W = p.w_hh + (int64_t)d * G * H * H;
  const int64_t out_ld = (int64_t)p.ndir * H;
  const float* hprev = has_prev ? p.out + ((int64_t)tp * B) * out_ld + (int64_t)d * H
                                : (p.h0 ? p.h0 + (int64_t)d * B * H : nullptr);
  const int64_t hprev_ld = has_prev ? out_ld : H;

  float acc[G][RB];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int r = 0; r < RB; ++r) acc[g][r] = 0.f;

  if (hprev) {  // first step with zero state: the recurrent product is exactly 0
    const bool vec = (H % 4 == 0) && ((((uintptr_t)W | (uintptr_t)hprev) & 15) == 0);
    for (int k0 = 0; k0 < H; k0 += FKC) {
      const int kc = min(FKC, H - k0);
      if (k0) __syncthreads();
      if (vec) {
        // all global loads of the tile are issued before the first shared store (one L2 round trip)
        constexpr int NW = G * FJ * (FKC / 4) / (FJ * FTB), NH = BB * (FKC / 4) / (FJ * FTB);
        float4 wv[NW], hv[NH];
#pragma unroll
        for (int i = 0; i < NW; ++i) {
          const int e = tid + i * (FJ * FTB), k4 = e % (FKC / 4), row = e / (FKC / 4), g = row / FJ, j = j0 + row % FJ;
          wv[i] = (k4 * 4 < kc && j < H) ? __ldg(reinterpret_cast<const float4*>(W + ((int64_t)g * H + j) * H + k0) + k4)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < NH; ++i) {
          const int e = tid + i * (FJ * FTB), k4 = e % (FKC / 4), bb = e / (FKC / 4), b = b0 + bb;
          hv[i] = (k4 * 4 < kc && b < B) ? *(reinterpret_cast<const float4*>(hprev + (int64_t)b * hprev_ld + k0) + k4)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < NW; ++i) {
          const int e = tid + i * (FJ * FTB), k4 = e % (FKC / 4), row = e / (FKC / 4);
          *reinterpret_cast<float4*>(&Ws[row * LDS_ + k4 * 4]) = wv[i];
        }
#pragma unroll
        for (int i = 0; i < NH; ++i) {
          const int e = tid + i * (FJ * FTB), k4 = e % (FKC / 4), bb = e / (FKC / 4);
          *reinterpret_cast<float4*>(&Hs[bb * LDS_ + k4 * 4]) = hv[i];
        }
      } else {
        for (int e = tid; e < G * FJ * FKC; e += FJ * FTB) {
          const int kk = e % FKC, row = e / FKC, g = row / FJ, jj = row % FJ;
          const int j = j0 + jj;
          Ws[row * LDS_ + kk] = (kk < kc && j < H) ? __ldg(W + ((int64_t)g * H + j) * H + k0 + kk) : 0.f;
        }
        for (int e = tid; e < BB * FKC; e += FJ * FTB) {
          const int kk = e % FKC, bb = e / FKC, b = b0 + bb;
          Hs[bb * LDS_ + kk] = (kk < kc && b < B) ? hprev[(int64_t)b * hprev_ld + k0 + kk] : 0.f;
        }
      }
      __syncthreads();
      const int kc4 = (kc + 3) & ~3;
      for (int kk = 0; kk < kc4; kk += 4) {
        float4 w[G], h[RB];
#pragma unroll
        for (int g = 0; g < G; ++g) w[g] = *reinterpret_cast<const float4*>(&Ws[(g * FJ + tj) * LDS_ + kk]);
#pragma unroll
        for (int r = 0; r < RB; ++r) h[r] = *reinterpret_cast<const float4*>(&Hs[(tb * RB + r) * LDS_ + kk]);
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
          for (int r = 0; r < RB; ++r) {
            acc[g][r] = fmaf(w[g].x, h[r].x, acc[g][r]);
            acc[g][r] = fmaf(w[g].y, h[r].y, acc[g][r]);
            acc[g][r] = fmaf(w[g].z, h[r].z, acc[g][r]);
            acc[g][r] = fmaf(w[g].w, h[r].w, acc[g][r]);
          }
      }
    }
  }

  const int j = j0 + tj;
  if (j >= H) return;
  const float* bh = p.b_hh + (int64_t)d * G * H;
#pragma unroll
  for (int r = 0; r < RB; ++r) {
    const int b = b0 + tb * RB + r;
    if (b >= B) continue;
    const int len = p.lengths ? (int)p.lengths[b] : T;
    const bool active = t < len;
    const int64_t row = (int64_t)t * B + b;
    float* gt = p.gates + (row * p.ndir + d) * G * H;
    float* o = p.out + row * out_ld + (int64_t)d * H + j;
    float* st = p.stash + (row * p.ndir + d) * H + j;
    if (!active) {
      *o = 0.f;
      *st = 0.f;
      continue;
    }
    float h;
    if (G == 4) {
      const float gi = sigmoidf_(gt[j] + acc[0][r] + bh[j]);
      const float gf = sigmoidf_(gt[H + j] + acc[1][r] + bh[H + j]);
      const float gg = tanhf(gt[2 * H + j] + acc[2][r] + bh[2 * H + j]);
      const float go = sigmoidf_(gt[3 * H + j] + acc[G - 1][r] + bh[3 * H + j]);
      const float cp = has_prev ? p.stash[(((int64_t)tp * B + b) * p.ndir + d) * H + j]
                                : (p.c0 ? p.c0[((int64_t)d * B + b) * H + j] : 0.f);
      const float c = gf * cp + gi * gg;
      h = go * tanhf(c);
      gt[j] = gi; gt[H + j] = gf; gt[2 * H + j] = gg; gt[3 * H + j] = go;
      *st = c;
    } else {
      const float hr = acc[0][r] + bh[j], hz = acc[1][r] + bh[H + j], hn = acc[2][r] + bh[2 * H + j];
      const float gr = sigmoidf_(gt[j] + hr);
      const float gz = sigmoidf_(gt[H + j] + hz);
      const float gn = tanhf(gt[2 * H + j] + gr * hn);
      const float hp = hprev ? hprev[(int64_t)b * hprev_ld + j] : 0.f;
      h = (1.f - gz) * gn + gz * hp;
      gt[j] = gr; gt[H + j] = gz; gt[2 * H + j] = gn;
      *st = hn;
    }
    *o = h;
    if (p.h_final && (d == 0 ? t == len - 1 : t == 0)) p.h_final[((int64_t)d * B + b) * H + j] = h;
  }
}

// ------------------------------------------------------------------ backward step
constexpr int BKT = 16;    // k columns per CTA
constexpr int BJS = 8;     // split of the reduction (j') range across warps
constexpr int BJC = 512;   // j' chunk resident in shared memory

struct StepBwd {
  int T, B, H, ndir, step, final_only;
  float* gates;          // in: activated gates; out: d pre-activations (x side)
  float* stash;          // LSTM: c_t (read); GRU: hn (read) -> d hn (written)
  const float* out;      // h_t
  const float* w_hh;
  const int64_t* lengths;
  const float* h0;
  const float* c0;
  const float* dout;     // [T,B,ndir*H] or null
  const float* dh_final; // [ndir,B,H] or null
  const float* dc_final;
  float* dh0;            // final_only outputs
  float* dc0;
  float* carry;          // [ndir,B,H]: LSTM dc carry / GRU direct dh carry
};

template <int G, int RB>
__global__ void __launch_bounds__(BJS * 32) rnn_step_bwd_kernel(StepBwd p) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ __align__(16) float smem[];
  constexpr int BB = 8 * RB;
  constexpr int WLD = BKT + 4;
  const int H = p.H, B = p.B, T = p.T, GH = G * p.H;
  const int jc_max = (min(GH, BJC) + 3) & ~3;  // chunk of the reduction range, padded to 4
  const int DLD = jc_max + 4;
  float* Ws = smem;                         // [jc_max][WLD]
  float* Ds = smem + (size_t)jc_max * WLD;  // [BB][DLD]
  // reduction scratch aliases Ws after the main loop
  const int d = blockIdx.z;
  const int t = p.final_only ? (d == 0 ? -1 : T) : (d == 0 ? T - 1 - p.step : p.step);
  const int tn = d == 0 ? t + 1 : t - 1;  // the step processed just before this one
  const bool has_next = tn >= 0 && tn < T;
  const int k0 = blockIdx.x * BKT, b0 = blockIdx.y * BB;
  const int tid = threadIdx.x, js = tid >> 5, lane = tid & 31;
  const int kq = lane & 3, bq = lane >> 2;  // 4 k-quads x 8 batch groups
  const float* W = p.w_hh + (int64_t)d * GH * H;

  float acc[RB][4];
#pragma unroll
  for (int r = 0; r < RB; ++r)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[r][q] = 0.f;

  if (has_next) {
    const bool vec = (H % 4 == 0) &&
                     ((((uintptr_t)W | (uintptr_t)p.gates | (uintptr_t)p.stash) & 15) == 0);
    for (int jbase = 0; jbase < GH; jbase += jc_max) {
      const int jc = min(jc_max, GH - jbase);
      if (jbase) __syncthreads();
      if (vec) {
        constexpr int NV = BJC * (BKT / 4) / (BJS * 32);       // float4 per thread for a full chunk
        constexpr int ND = 8 * RB * (BJC / 4) / (BJS * 32);
        float4 wv[NV], dv[ND];
        const int nw4 = jc_max * (BKT / 4), jq = jc_max / 4, nd4 = BB * jq;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int e = tid + i * (BJS * 32), k4 = e % (BKT / 4), jj = e / (BKT / 4);
          wv[i] = (e < nw4 && jj < jc && k0 + k4 * 4 < H)
                      ? __ldg(reinterpret_cast<const float4*>(W + (int64_t)(jbase + jj) * H + k0) + k4)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < ND; ++i) {
          const int e = tid + i * (BJS * 32), j4 = e % jq, bb = e / jq, b = b0 + bb, j = jbase + j4 * 4;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (e < nd4 && j4 * 4 < jc && b < B) {
            const int64_t row = ((int64_t)tn * B + b) * p.ndir + d;
            v = (G == 4 || j < 2 * H) ? *reinterpret_cast<const float4*>(p.gates + row * GH + j)
                                      : *reinterpret_cast<const float4*>(p.stash + row * H + (j - 2 * H));
          }
          dv[i] = v;
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int e = tid + i * (BJS * 32), k4 = e % (BKT / 4), jj = e / (BKT / 4);
          if (e < nw4) *reinterpret_cast<float4*>(&Ws[jj * WLD + k4 * 4]) = wv[i];
        }
#pragma unroll
        for (int i = 0; i < ND; ++i) {
          const int e = tid + i * (BJS * 32), j4 = e % jq, bb = e / jq;
          if (e < nd4) *reinterpret_cast<float4*>(&Ds[bb * DLD + j4 * 4]) = dv[i];
        }
      } else {
        for (int e = tid; e < jc_max * BKT; e += BJS * 32) {
          const int kk = e % BKT, jj = e / BKT;
          Ws[jj * WLD + kk] = (jj < jc && k0 + kk < H) ? __ldg(W + (int64_t)(jbase + jj) * H + k0 + kk) : 0.f;
        }
        for (int e = tid; e < BB * jc_max; e += BJS * 32) {
          const int jj = e % jc_max, bb = e / jc_max, b = b0 + bb, j = jbase + jj;
          float v = 0.f;
          if (jj < jc && b < B) {
            const int64_t row = ((int64_t)tn * B + b) * p.ndir + d;
            if (G == 4 || j < 2 * H) v = p.gates[row * GH + j];
            else v = p.stash[row * H + (j - 2 * H)];
          }
          Ds[bb * DLD + jj] = v;
        }
      }
      __syncthreads();
      const int per = ((jc + BJS * 4 - 1) / (BJS * 4)) * 4;  // j' per split, multiple of 4
      const int jlo = js * per, jhi = min(jc_max, jlo + per);
      for (int jj = jlo; jj < jhi; jj += 4) {
        float4 w[4], dv[RB];
#pragma unroll
        for (int u = 0; u < 4; ++u) w[u] = *reinterpret_cast<const float4*>(&Ws[(jj + u) * WLD + kq * 4]);
#pragma unroll
        for (int r = 0; r < RB; ++r) dv[r] = *reinterpret_cast<const float4*>(&Ds[(bq * RB + r) * DLD + jj]);
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          const float dd[4] = {dv[r].x, dv[r].y, dv[r].z, dv[r].w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            acc[r][0] = fmaf(dd[u], w[u].x, acc[r][0]);
            acc[r][1] = fmaf(dd[u], w[u].y, acc[r][1]);
            acc[r][2] = fmaf(dd[u], w[u].z, acc[r][2]);
            acc[r][3] = fmaf(dd[u], w[u].w, acc[r][3]);
          }
        }
      }
    }
  }
  // cross-split reduction through shared memory: red[js][bb][k]
  __syncthreads();
  float* red = smem;  // BJS*BB*BKT floats <= jc_max*WLD whenever jc_max >= BB*BJS*BKT/WLD
#pragma unroll
  for (int r = 0; r < RB; ++r)
#pragma unroll
    for (int q = 0; q < 4; ++q) red[(js * BB + bq * RB + r) * BKT + kq * 4 + q] = acc[r][q];
  __syncthreads();

  for (int e = tid; e < BB * BKT; e += BJS * 32) {
    const int kk = e % BKT, bb = e / BKT, k = k0 + kk, b = b0 + bb;
    if (k >= H || b >= B) continue;
    float m = 0.f;
#pragma unroll
    for (int s = 0; s < BJS; ++s) m += red[(s * BB + bb) * BKT + kk];
    const int64_t cidx = ((int64_t)d * B + b) * H + k;
    if (p.final_only) {
      if (G == 4) {
        if (p.dh0) p.dh0[cidx] = m;
        if (p.dc0) p.dc0[cidx] = p.carry[cidx];
      } else if (p.dh0) {
        p.dh0[cidx] = m + p.carry[cidx];
      }
      continue;
    }
    const int len = p.lengths ? (int)p.lengths[b] : T;
    const int64_t row = ((int64_t)t * B + b) * p.ndir + d;
    float* gt = p.gates + row * GH;
    float* st = p.stash + row * H + k;
    if (t >= len) {
#pragma unroll
      for (int g = 0; g < G; ++g) gt[g * H + k] = 0.f;
      if (G == 3) *st = 0.f;
      continue;
    }
    const bool inject = d == 0 ? t == len - 1 : t == 0;
    float dh = p.dout ? p.dout[((int64_t)t * B + b) * p.ndir * H + (int64_t)d * H + k] : 0.f;
    const int tp = d == 0 ? t - 1 : t + 1;  // forward-time predecessor
    const bool has_prev = tp >= 0 && tp < T;
    if (G == 4) {
      float dc_in;
      if (inject) {
        dh += p.dh_final ? p.dh_final[cidx] : 0.f;
        dc_in = p.dc_final ? p.dc_final[cidx] : 0.f;
      } else {
        dh += m;
        dc_in = p.carry[cidx];
      }
      const float gi = gt[k], gf = gt[H + k], gg = gt[2 * H + k], go = gt[3 * H + k];
      const float c = *st;
      const float cp = has_prev ? p.stash[(((int64_t)tp * B + b) * p.ndir + d) * H + k]
                                : (p.c0 ? p.c0[cidx] : 0.f);
      const float tc = tanhf(c);
      const float dc = dh * go * (1.f - tc * tc) + dc_in;
      gt[k] = dc * gg * gi * (1.f - gi);
      gt[H + k] = dc * cp * gf * (1.f - gf);
      gt[2 * H + k] = dc * gi * (1.f - gg * gg);
      gt[3 * H + k] = dh * tc * go * (1.f - go);
      p.carry[cidx] = dc * gf;
    } else {
      if (inject) dh += p.dh_final ? p.dh_final[cidx] : 0.f;
      else dh += m + p.carry[cidx];
      const float gr = gt[k], gz = gt[H + k], gn = gt[2 * H + k];
      const float hn = *st;
      const float hp = has_prev ? p.out[((int64_t)tp * B + b) * p.ndir * H + (int64_t)d * H + k]
                                : (p.h0 ? p.h0[cidx] : 0.f);
      const float da_n = dh * (1.f - gz) * (1.f - gn * gn);
      gt[k] = da_n * hn * gr * (1.f - gr);
      gt[H + k] = dh * (hp - gn) * gz * (1.f - gz);
      gt[2 * H + k] = da_n;
      *st = da_n * gr;
      p.carry[cidx] = dh * gz;
    }
  }
}

static size_t fwd_smem(int G, int RB) { return (size_t)(G * FJ + FTB * RB) * (FKC + 4) * sizeof(float); }
static size_t bwd_smem(int G, int RB, int H) {
  const int jc = ((G * H < BJC ? G * H : BJC) + 3) & ~3;
  size_t a = (size_t)jc * (BKT + 4) + (size_t)8 * RB * (jc + 4);
  size_t r = (size_t)BJS * 8 * RB * BKT;
  return (a > r ? a : r) * sizeof(float);
}

template <int G, int RB>
static int launch_fwd(const StepFwd& p, cudaStream_t s) {
  const size_t sm = fwd_smem(G, RB);
  if (sm > 48 * 1024)
    cudaFuncSetAttribute(rnn_step_fwd_kernel<G, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  dim3 grid(ceil_div(p.H, FJ), ceil_div(p.B, FTB * RB), p.ndir);
  launch_pdl(rnn_step_fwd_kernel<G, RB>, dim3(grid), dim3(FJ * FTB), sm, s, p);
  return 0;
}
template <int G, int RB>
static int launch_bwd(const StepBwd& p, cudaStream_t s) {
  const size_t sm = bwd_smem(G, RB, p.H);
  if (sm > 48 * 1024)
    cudaFuncSetAttribute(rnn_step_bwd_kernel<G, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  dim3 grid(ceil_div(p.H, BKT), ceil_div(p.B, 8 * RB), p.ndir);
  launch_pdl(rnn_step_bwd_kernel<G, RB>, dim3(grid), dim3(BJS * 32), sm, s, p);
  return 0;
}

// persistent tcgen05 path (rnn_persistent.cu); returns -1 when the shape is not supported
int rnn_layer_fwd_tc(int mode, int T, int B, int H, int ndir, float* gates, const float* w_hh,
                     const float* b_hh, const int64_t* lengths, const float* h0, const float* c0,
                     float* out, float* stash, float* h_final, const slnlp_rnn_extras* ex, cudaStream_t s);
int rnn_layer_bwd_tc(int mode, int T, int B, int H, int ndir, float* gates, float* stash, const float* out,
                     const float* w_hh, const int64_t* lengths, const float* h0, const float* c0,
                     const float* dout, const float* dh_final, const float* dc_final, float* dh0, float* dc0,
                     const slnlp_rnn_extras* ex, cudaStream_t s);

// cluster-persistent path, W_hh slices resident in the TMEM of a thread-block cluster, H = 256 / 512
// (rnn_cluster.cu); -1 = unsupported
int rnn_layer_fwd_cluster(int mode, int T, int B, int H, int ndir, float* gates, const float* w_hh, const float* b_hh,
                          const int64_t* lengths, const float* h0, const float* c0, float* out, float* stash,
                          float* h_final, cudaStream_t s);
int rnn_layer_bwd_cluster(int mode, int T, int B, int H, int ndir, float* gates, float* stash, const float* out,
                          const float* w_hh, const int64_t* lengths, const float* h0, const float* c0,
                          const float* dout, const float* dh_final, const float* dc_final, float* dh0, float* dc0,
                          cudaStream_t s);
// per-timestep TMA + tcgen05 kind::tf32 path for any H % 32 == 0 (rnn_step_tc.cu); -1 = unsupported
int rnn_layer_fwd_tcstep(int mode, int T, int B, int H, int ndir, float* gates, const float* w_hh, const float* b_hh,
                         const int64_t* lengths, const float* h0, const float* c0, float* out, float* stash,
                         float* h_final, cudaStream_t s);
int rnn_layer_bwd_tcstep(int mode, int T, int B, int H, int ndir, float* gates, float* stash, const float* out,
                         const float* w_hh, const int64_t* lengths, const float* h0, const float* c0,
                         const float* dout, const float* dh_final, const float* dc_final, float* dh0, float* dc0,
                         float* carry, cudaStream_t s);

// persistent fp32-FMA path for H = 128 (rnn_persistent_f32.cu): W_hh resident in registers + shared memory,
// one CTA per (sequence, direction); -1 = unsupported
int rnn_layer_fwd_pf32(int mode, int T, int B, int H, int ndir, float* gates, const float* w_hh, const float* b_hh,
                       const int64_t* lengths, const float* h0, const float* c0, float* out, float* stash,
                       float* h_final, const slnlp_rnn_extras* ex, cudaStream_t s);
int rnn_layer_bwd_pf32(int mode, int T, int B, int H, int ndir, float* gates, float* stash, const float* out,
                       const float* w_hh, const int64_t* lengths, const float* h0, const float* c0, const float* dout,
                       const float* dh_final, const float* dc_final, float* dh0, float* dc0,
                       const slnlp_rnn_extras* ex, cudaStream_t s);

}  // namespace slnlp

using namespace slnlp;

// the families that implement slnlp_rnn_extras (h_final in the concatenated layout, fused inter-layer dropout)
static bool extras_family(int precision, int T, int H, const float* w_hh) {
  return T > 1 && H == 128 && (((uintptr_t)w_hh & 15) == 0) && (precision == 1 || precision == 0);
}

extern "C" int slnlp_rnn_extras_supported(int precision, int T, int B, int H, int ndir) {
  (void)B; (void)ndir;
  if (precision == 0) {
    const char* e = getenv("SLNLP_PERSIST_F32");
    if (e && e[0] == '0') return 0;
  }
  return (T > 1 && H == 128) ? 1 : 0;
}

extern "C" int slnlp_rnn_layer_fwd_ex(int mode, int precision, int T, int B, int H, int ndir, float* gates,
                                      const float* w_hh, const float* b_hh, const int64_t* lengths,
                                      const float* h0, const float* c0, float* out, float* stash,
                                      float* h_final, const slnlp_rnn_extras* ex, slnlp_stream_t stream);

extern "C" int slnlp_rnn_layer_fwd(int mode, int precision, int T, int B, int H, int ndir, float* gates,
                                   const float* w_hh, const float* b_hh, const int64_t* lengths,
                                   const float* h0, const float* c0, float* out, float* stash,
                                   float* h_final, slnlp_stream_t stream) {
  return slnlp_rnn_layer_fwd_ex(mode, precision, T, B, H, ndir, gates, w_hh, b_hh, lengths, h0, c0, out, stash, h_final,
                                nullptr, stream);
}

extern "C" int slnlp_rnn_layer_fwd_ex(int mode, int precision, int T, int B, int H, int ndir, float* gates,
                                      const float* w_hh, const float* b_hh, const int64_t* lengths,
                                      const float* h0, const float* c0, float* out, float* stash,
                                      float* h_final, const slnlp_rnn_extras* ex, slnlp_stream_t stream) {
  const bool wants = ex && (ex->hfinal_cat || ex->out_drop);
  SLNLP_CHECK_ARG(!wants || (extras_family(precision, T, H, w_hh) && !h0 && !c0),
                  "rnn_layer_fwd_ex: extras need the persistent kernels (H = 128, T > 1, no initial state); "
                  "ask slnlp_rnn_extras_supported first");
  SLNLP_CHECK_ARG(!ex || !ex->out_drop || ((ex->rng || ex->mask) && ex->p_drop >= 0.f && ex->p_drop < 1.f),
                  "rnn_layer_fwd_ex: bad dropout arguments");
  SLNLP_CHECK_ARG(mode == SLNLP_MODE_LSTM || mode == SLNLP_MODE_GRU, "rnn_layer_fwd: bad mode %d", mode);
  SLNLP_CHECK_ARG(T > 0 && B > 0 && H > 0 && (ndir == 1 || ndir == 2), "rnn_layer_fwd: bad shape");
  SLNLP_CHECK_ARG(gates && w_hh && b_hh && out && stash, "rnn_layer_fwd: null pointer");
  cudaStream_t s = as_stream(stream);
  if (precision == 1) {
    // a single step (the decoder cell) is not worth staging W_hh into tensor memory: per-step kernel
    const int rc = T > 1 ? rnn_layer_fwd_tc(mode, T, B, H, ndir, gates, w_hh, b_hh, lengths, h0, c0, out, stash, h_final, ex, s) : -1;
    if (rc >= 0) return rc;
    SLNLP_CHECK_ARG(!wants, "rnn_layer_fwd_ex: extras requested but the persistent kernel rejected the shape");
    const int rc1 = T <= 1 ? -1 : rnn_layer_fwd_cluster(mode, T, B, H, ndir, gates, w_hh, b_hh, lengths, h0, c0, out, stash, h_final, s);
    if (rc1 >= 0) return rc1;
    const int rc2 = rnn_layer_fwd_tcstep(mode, T, B, H, ndir, gates, w_hh, b_hh, lengths, h0, c0, out, stash, h_final, s);
    if (rc2 >= 0) return rc2;
    // unsupported shape for the tensor-core kernels: the general path below is still CUDA
  }
  if (precision == 0) {
    const int rc = rnn_layer_fwd_pf32(mode, T, B, H, ndir, gates, w_hh, b_hh, lengths, h0, c0, out, stash, h_final, ex, s);
    if (rc >= 0) return rc;
    SLNLP_CHECK_ARG(!wants, "rnn_layer_fwd_ex: extras requested but the persistent fp32 kernel is off or rejected the shape");
  }
  StepFwd p{T, B, H, ndir, 0, gates, w_hh, b_hh, lengths, h0, c0, out, stash, h_final};
  const bool big = B > 256;
  for (int step = 0; step < T; ++step) {
    p.step = step;
    if (mode == SLNLP_MODE_LSTM) { if (big) launch_fwd<4, 4>(p, s); else launch_fwd<4, 2>(p, s); }
    else { if (big) launch_fwd<3, 4>(p, s); else launch_fwd<3, 2>(p, s); }
  }
  note_launches(T - 1);
  SLNLP_LAUNCH_OK("rnn_layer_fwd");
  return 0;
}

extern "C" int slnlp_rnn_layer_bwd_ex(int mode, int precision, int T, int B, int H, int ndir, float* gates,
                                      float* stash, const float* out, const float* w_hh,
                                      const int64_t* lengths, const float* h0, const float* c0,
                                      const float* dout, const float* dh_final, const float* dc_final,
                                      float* dh0, float* dc0, float* carry, const slnlp_rnn_extras* ex,
                                      slnlp_stream_t stream);

extern "C" int slnlp_rnn_layer_bwd(int mode, int precision, int T, int B, int H, int ndir, float* gates,
                                   float* stash, const float* out, const float* w_hh,
                                   const int64_t* lengths, const float* h0, const float* c0,
                                   const float* dout, const float* dh_final, const float* dc_final,
                                   float* dh0, float* dc0, float* carry, slnlp_stream_t stream) {
  return slnlp_rnn_layer_bwd_ex(mode, precision, T, B, H, ndir, gates, stash, out, w_hh, lengths, h0, c0, dout, dh_final,
                                dc_final, dh0, dc0, carry, nullptr, stream);
}

extern "C" int slnlp_rnn_layer_bwd_ex(int mode, int precision, int T, int B, int H, int ndir, float* gates,
                                      float* stash, const float* out, const float* w_hh,
                                      const int64_t* lengths, const float* h0, const float* c0,
                                      const float* dout, const float* dh_final, const float* dc_final,
                                      float* dh0, float* dc0, float* carry, const slnlp_rnn_extras* ex,
                                      slnlp_stream_t stream) {
  const bool wants = ex && (ex->hfinal_cat || (ex->dout_dropped && ex->p_drop > 0.f));
  SLNLP_CHECK_ARG(!wants || (extras_family(precision, T, H, w_hh) && !h0 && !c0 && !dh0 && !dc0),
                  "rnn_layer_bwd_ex: extras need the persistent kernels (H = 128, T > 1, no initial state)");
  SLNLP_CHECK_ARG(!ex || !ex->dout_dropped || ex->p_drop <= 0.f || ex->rng || ex->mask,
                  "rnn_layer_bwd_ex: dropout replay needs rng or a mask");
  SLNLP_CHECK_ARG(mode == SLNLP_MODE_LSTM || mode == SLNLP_MODE_GRU, "rnn_layer_bwd: bad mode %d", mode);
  SLNLP_CHECK_ARG(T > 0 && B > 0 && H > 0 && (ndir == 1 || ndir == 2), "rnn_layer_bwd: bad shape");
  SLNLP_CHECK_ARG(gates && stash && out && w_hh && carry, "rnn_layer_bwd: null pointer");
  SLNLP_CHECK_ARG(!(dh0 || dc0) || !lengths, "rnn_layer_bwd: dh0/dc0 need lengths == NULL");
  cudaStream_t s = as_stream(stream);
  if (precision == 1) {
    const int rc = T > 1 ? rnn_layer_bwd_tc(mode, T, B, H, ndir, gates, stash, out, w_hh, lengths, h0, c0, dout, dh_final,
                                            dc_final, dh0, dc0, ex, s) : -1;
    if (rc >= 0) return rc;
    SLNLP_CHECK_ARG(!wants, "rnn_layer_bwd_ex: extras requested but the persistent kernel rejected the shape");
    const int rc1 = T <= 1 ? -1 : rnn_layer_bwd_cluster(mode, T, B, H, ndir, gates, stash, out, w_hh, lengths, h0, c0, dout, dh_final,
                                          dc_final, dh0, dc0, s);
    if (rc1 >= 0) return rc1;
    const int rc2 = rnn_layer_bwd_tcstep(mode, T, B, H, ndir, gates, stash, out, w_hh, lengths, h0, c0, dout, dh_final,
                                         dc_final, dh0, dc0, carry, s);
    if (rc2 >= 0) return rc2;
  }
  if (precision == 0) {
    const int rc = rnn_layer_bwd_pf32(mode, T, B, H, ndir, gates, stash, out, w_hh, lengths, h0, c0, dout, dh_final,
                                      dc_final, dh0, dc0, ex, s);
    if (rc >= 0) return rc;
    SLNLP_CHECK_ARG(!wants, "rnn_layer_bwd_ex: extras requested but the persistent fp32 kernel is off or rejected the shape");
  }
  StepBwd p{T, B, H, ndir, 0, 0, gates, stash, out, w_hh, lengths, h0, c0, dout, dh_final, dc_final, dh0, dc0, carry};
  const bool big = B > 256;
  auto go = [&]() {
    if (mode == SLNLP_MODE_LSTM) { if (big) launch_bwd<4, 4>(p, s); else launch_bwd<4, 2>(p, s); }
    else { if (big) launch_bwd<3, 4>(p, s); else launch_bwd<3, 2>(p, s); }
  };
  for (int step = 0; step < T; ++step) {
    p.step = step;
    go();
  }
  if (dh0 || dc0) {
    p.final_only = 1;
    go();
  }
  note_launches(T - 1 + ((dh0 || dc0) ? 1 : 0));
  SLNLP_LAUNCH_OK("rnn_layer_bwd");
  return 0;
}
