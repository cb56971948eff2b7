// K4 on the fp32 (1e-5, identical-argmax) path: persistent recurrent layer for H = 128 in exact fp32 FMA.
//
// The general fp32 path launches one kernel per timestep (rnn_layer.cu: 7 us per dependent step at the
// reference's batch of 50 - launch + W_hh re-read every step).  Here one CTA owns ONE sequence and one
// direction for all timesteps and W_hh never leaves the SM:
//   * thread r = (gate g, unit j) owns row r of W_hh (forward) / column j of gate block g (BPTT): RW of its
//     128 weights live in REGISTERS, the other 128 - RW in shared memory as [k/4][thread] float4 (one
//     conflict-free 128-bit load per thread and k-quad) - 256 KB of fp32 weights do not fit either alone;
//   * a step is 128 FMAs per thread against h_{t-1} (BPTT: dG_t) read as broadcast float4 from shared
//     memory, the gate nonlinearity applied by the owning thread (full-precision expf / tanhf), a
//     shared-memory exchange, and the cell update by the 128 unit threads; two CTA barriers per step;
//   * a sequence's frozen steps (t >= length) cost nothing but the zero stores; results and the next
//     step's hoisted projection move one value per thread and step.
// fp32 throughout, fixed summation order (k ascending, two interleaved accumulators): deterministic, and
// within 1e-6 of the per-step kernels (different association only).
#include <stdlib.h>
#include "common.cuh"

namespace slnlp {

constexpr int FH = 128;    // hidden size handled here
constexpr int RW = 80;     // weights per thread kept in registers (multiple of 4); the rest in shared memory
constexpr int SWQ = (FH - RW) / 4;   // float4 per thread in shared memory

struct PF32Fwd {
  int T, B, ndir;
  float* gates;          // [T,B,ndir,G,H]: in x W_ih^T + b_ih, out activated gates
  const float* w_hh;     // [ndir,G*H,H]
  const float* b_hh;     // [ndir,G*H]
  const int64_t* lengths;
  float* out;            // [T,B,ndir*H]
  float* stash;          // [T,B,ndir,H]
  float* h_final;        // [ndir,B,H] or null
  int64_t hf_d, hf_b;    // h_final strides (direction, batch)
  float* out_drop;       // dropout(out) for the next layer's input, or null
  float p_drop;
  const uint64_t* rng;
  uint32_t site;
  const float* mask;     // precomputed factors instead of Philox, or null
};

__device__ __forceinline__ float dropout_factor32(uint64_t seed, uint64_t step, uint32_t site, int64_t e, float p) {
  float u[4];
  philox_uniform4(seed, step, site, (uint64_t)(e >> 2), u);
  const int l = (int)(e & 3);
  const float uu = l == 0 ? u[0] : l == 1 ? u[1] : l == 2 ? u[2] : u[3];
  return uu < 1.f - p ? 1.f / (1.f - p) : 0.f;
}

template <int G>
__global__ void __launch_bounds__(G * FH, 1) rnn_pf32_fwd_kernel(PF32Fwd p) {
  pdl_launch_dependents();
  constexpr int H = FH, NT = G * FH;
  extern __shared__ __align__(16) float smem[];
  float4* Ws = reinterpret_cast<float4*>(smem);                 // [SWQ][NT]
  float* hs = smem + (size_t)SWQ * NT * 4;                      // [H]   h_{t-1}
  float* act = hs + H;                                          // [NT]  activated gates (GRU row 2: W_hn h + b_hn)
  float* xn_s = act + NT;                                       // [H]   GRU: x half of the candidate gate
  const int tid = threadIdx.x, g = tid / H, j = tid % H;
  const int b = blockIdx.x, d = blockIdx.y;
  const int T = p.T, B = p.B;

  // ---- weights-only prologue (runs under the tail of the preceding kernel: programmatic dependent launch)
  const float* wrow = p.w_hh + ((int64_t)d * NT + tid) * H;
  float wr[RW];
#pragma unroll
  for (int q = 0; q < RW / 4; ++q) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(wrow) + q);
    wr[4 * q] = v.x; wr[4 * q + 1] = v.y; wr[4 * q + 2] = v.z; wr[4 * q + 3] = v.w;
  }
#pragma unroll
  for (int q = 0; q < SWQ; ++q) Ws[q * NT + tid] = __ldg(reinterpret_cast<const float4*>(wrow) + RW / 4 + q);
  const float bias = p.b_hh[(int64_t)d * NT + tid];
  if (tid < H) hs[tid] = 0.f;

  pdl_wait();   // ---- from here on: this step's hoisted projection and lengths

  const int len = p.lengths ? (int)p.lengths[b] : T;
  const int64_t gstride = (int64_t)B * p.ndir * NT, ostride = (int64_t)B * p.ndir * H;
  float* gp = p.gates + ((int64_t)b * p.ndir + d) * NT + tid;           // + t * gstride
  float* op = p.out + (int64_t)b * p.ndir * H + (int64_t)d * H + j;     // + t * ostride   (unit threads)
  float* sp = p.stash + ((int64_t)b * p.ndir + d) * H + j;
  float c_state = 0.f, h_state = 0.f;
  const bool drop = p.out_drop != nullptr;
  const uint64_t rseed = (drop && !p.mask) ? p.rng[0] : 0, rstep = (drop && !p.mask) ? p.rng[1] : 0;
  const int64_t ddelta = p.out_drop - p.out;
  // frozen steps: out = stash = 0 (gates untouched), no recurrence
  if (tid < H)
    for (int t = len; t < T; ++t) {
      op[(int64_t)t * ostride] = 0.f;
      sp[(int64_t)t * ostride] = 0.f;
      if (drop) op[(int64_t)t * ostride + ddelta] = 0.f;
    }
  __syncthreads();
  // the sequence's own steps: direction 0 walks t = 0 .. len-1, direction 1 walks t = len-1 .. 0
  float x = len > 0 ? gp[(int64_t)(d == 0 ? 0 : len - 1) * gstride] : 0.f;
  for (int step = 0; step < len; ++step) {
    const int t = d == 0 ? step : len - 1 - step;
    float xnext = 0.f;
    if (step + 1 < len) xnext = gp[(int64_t)(d == 0 ? t + 1 : t - 1) * gstride];
    float mk = 0.f;      // precomputed dropout factor of this step's output: loaded ahead of the recurrent product
    if (drop && p.mask && tid < H) mk = p.mask[(op + (int64_t)t * ostride) - p.out];
    // recurrent product: row tid of W_hh against h_{t-1}
    float a0 = 0.f, a1 = 0.f;
    if (step > 0) {
#pragma unroll
      for (int q = 0; q < RW / 4; ++q) {
        const float4 h4 = *reinterpret_cast<const float4*>(hs + 4 * q);
        a0 = fmaf(wr[4 * q], h4.x, a0);
        a1 = fmaf(wr[4 * q + 1], h4.y, a1);
        a0 = fmaf(wr[4 * q + 2], h4.z, a0);
        a1 = fmaf(wr[4 * q + 3], h4.w, a1);
      }
#pragma unroll
      for (int q = 0; q < SWQ; ++q) {
        const float4 w4 = Ws[q * NT + tid];
        const float4 h4 = *reinterpret_cast<const float4*>(hs + RW + 4 * q);
        a0 = fmaf(w4.x, h4.x, a0);
        a1 = fmaf(w4.y, h4.y, a1);
        a0 = fmaf(w4.z, h4.z, a0);
        a1 = fmaf(w4.w, h4.w, a1);
      }
    }
    const float hh = a0 + a1 + bias;
    float av;
    if (G == 4) av = g == 2 ? tanhf(x + hh) : sigmoidf_(x + hh);
    else if (g < 2) av = sigmoidf_(x + hh);
    else {
      av = hh;           // W_hn h + b_hn: r applies to it
      xn_s[j] = x;
    }
    act[tid] = av;
    if (G == 4 || g < 2) gp[(int64_t)t * gstride] = av;   // activated gate -> BPTT stash
    __syncthreads();
    if (tid < H) {
      float h;
      if (G == 4) {
        const float c = act[H + j] * c_state + act[j] * act[2 * H + j];
        h = act[3 * H + j] * tanhf(c);
        c_state = c;
        sp[(int64_t)t * ostride] = c;
      } else {
        const float hn = act[2 * H + j];
        const float gn = tanhf(xn_s[j] + act[j] * hn);
        const float gz = act[H + j];
        h = (1.f - gz) * gn + gz * h_state;
        gp[(int64_t)t * gstride + 2 * H] = gn;             // gp of thread j < H points at gate 0: + 2H = gate n
        sp[(int64_t)t * ostride] = hn;
      }
      h_state = h;
      hs[j] = h;
      op[(int64_t)t * ostride] = h;
      if (drop) {
        float* oq = op + (int64_t)t * ostride;
        oq[ddelta] = h * (p.mask ? mk : dropout_factor32(rseed, rstep, p.site, oq - p.out, p.p_drop));
      }
      if (p.h_final && step == len - 1) p.h_final[(int64_t)d * p.hf_d + (int64_t)b * p.hf_b + j] = h;
    }
    x = xnext;
    __syncthreads();
  }
}

// ---------------------------------------------------------------- BPTT twin
struct PF32Bwd {
  int T, B, ndir;
  float* gates;          // in: activated gates; out: d pre-activations (x side)
  float* stash;          // LSTM: c_t; GRU: hn -> d hn
  const float* out;
  const float* w_hh;
  const int64_t* lengths;
  const float* dout;
  const float* dh_final;
  const float* dc_final;
  int64_t hf_d, hf_b;    // dh_final / dc_final strides (direction, batch)
  float p_drop;          // > 0: dout is the gradient of dropout(out)
  const uint64_t* rng;
  uint32_t site;
  const float* mask;
};

template <int G>
__global__ void __launch_bounds__(G * FH, 1) rnn_pf32_bwd_kernel(PF32Bwd p) {
  pdl_launch_dependents();
  constexpr int H = FH, NT = G * FH;
  extern __shared__ __align__(16) float smem[];
  float4* Ws = reinterpret_cast<float4*>(smem);                 // [SWQ][NT]
  float* dgs = smem + (size_t)SWQ * NT * 4;                     // [NT]  h-side d(pre-activations) of the step
  float* part = dgs + NT;                                       // [G][H] partial dh per gate block
  const int tid = threadIdx.x, g = tid / H, k = tid % H;
  const int b = blockIdx.x, d = blockIdx.y;
  const int T = p.T, B = p.B;

  // thread (g, k): column k of gate block g of W_hh, i.e. W[(g*H + i)*H + k], i = 0..H-1
  const float* wcol = p.w_hh + ((int64_t)d * NT + (int64_t)g * H) * H + k;
  float wr[RW];
#pragma unroll
  for (int i = 0; i < RW; ++i) wr[i] = __ldg(wcol + (int64_t)i * H);
#pragma unroll
  for (int q = 0; q < SWQ; ++q) {
    float4 v;
    v.x = __ldg(wcol + (int64_t)(RW + 4 * q) * H);
    v.y = __ldg(wcol + (int64_t)(RW + 4 * q + 1) * H);
    v.z = __ldg(wcol + (int64_t)(RW + 4 * q + 2) * H);
    v.w = __ldg(wcol + (int64_t)(RW + 4 * q + 3) * H);
    Ws[q * NT + tid] = v;
  }

  pdl_wait();

  const int len = p.lengths ? (int)p.lengths[b] : T;
  const int64_t gstride = (int64_t)B * p.ndir * NT, ostride = (int64_t)B * p.ndir * H;
  float* gp = p.gates + ((int64_t)b * p.ndir + d) * NT + tid;
  float* g0 = p.gates + ((int64_t)b * p.ndir + d) * NT + k;            // gate 0 of unit k (unit threads)
  float* sp = p.stash + ((int64_t)b * p.ndir + d) * H + k;
  const float* op = p.out + (int64_t)b * p.ndir * H + (int64_t)d * H + k;
  const float* dp = p.dout ? p.dout + (int64_t)b * p.ndir * H + (int64_t)d * H + k : nullptr;
  const int64_t cidx = (int64_t)d * p.hf_d + (int64_t)b * p.hf_b + k;
  const bool undrop = p.p_drop > 0.f && dp != nullptr;
  const uint64_t rseed = (undrop && !p.mask) ? p.rng[0] : 0, rstep = (undrop && !p.mask) ? p.rng[1] : 0;
  // frozen steps: d(pre-activations) = 0 (the hoisted dW / dx GEMMs read every row)
  for (int t = len; t < T; ++t) gp[(int64_t)t * gstride] = 0.f;
  if (G == 3 && tid < H)
    for (int t = len; t < T; ++t) sp[(int64_t)t * ostride] = 0.f;
  float carry = 0.f;     // LSTM: dc carry; GRU: direct dh carry (dh * z)
  // BPTT walks against the forward direction: d = 0 from t = len-1 down, d = 1 from t = 0 up
  for (int step = 0; step < len; ++step) {
    const int t = d == 0 ? len - 1 - step : step;
    const int tp = d == 0 ? t - 1 : t + 1;            // the state forward step t started from
    const bool has_prev = d == 0 ? tp >= 0 : tp < len;
    if (tid < H) {
      float m = 0.f;
      if (step > 0) {
#pragma unroll
        for (int q = 0; q < G; ++q) m += part[q * H + k];
      }
      const bool inject = step == 0;
      float dh = 0.f;
      if (dp) {
        const float* dq = dp + (int64_t)t * ostride;
        dh = *dq;
        if (undrop) dh *= p.mask ? p.mask[dq - p.dout] : dropout_factor32(rseed, rstep, p.site, dq - p.dout, p.p_drop);
      }
      float* gt = g0 + (int64_t)t * gstride;
      if (G == 4) {
        float dc_in;
        if (inject) {
          dh += p.dh_final ? p.dh_final[cidx] : 0.f;
          dc_in = p.dc_final ? p.dc_final[cidx] : 0.f;
        } else {
          dh += m;
          dc_in = carry;
        }
        const float gi = gt[0], gf = gt[H], gg = gt[2 * H], go = gt[3 * H];
        const float tc = tanhf(sp[(int64_t)t * ostride]);
        const float cprev = has_prev ? sp[(int64_t)tp * ostride] : 0.f;
        const float dc = dh * go * (1.f - tc * tc) + dc_in;
        const float d0 = dc * gg * gi * (1.f - gi), d1 = dc * cprev * gf * (1.f - gf);
        const float d2 = dc * gi * (1.f - gg * gg), d3 = dh * tc * go * (1.f - go);
        gt[0] = d0; gt[H] = d1; gt[2 * H] = d2; gt[3 * H] = d3;
        dgs[k] = d0; dgs[H + k] = d1; dgs[2 * H + k] = d2; dgs[3 * H + k] = d3;
        carry = dc * gf;
      } else {
        if (inject) dh += p.dh_final ? p.dh_final[cidx] : 0.f;
        else dh += m + carry;
        const float gr = gt[0], gz = gt[H], gn = gt[2 * H];
        const float hn = sp[(int64_t)t * ostride];
        const float hprev = has_prev ? op[(int64_t)tp * ostride] : 0.f;
        const float da_n = dh * (1.f - gz) * (1.f - gn * gn);
        const float d0 = da_n * hn * gr * (1.f - gr), d1 = dh * (hprev - gn) * gz * (1.f - gz);
        gt[0] = d0; gt[H] = d1; gt[2 * H] = da_n;
        sp[(int64_t)t * ostride] = da_n * gr;
        dgs[k] = d0; dgs[H + k] = d1; dgs[2 * H + k] = da_n * gr;
        carry = dh * gz;
      }
    }
    __syncthreads();
    if (step + 1 < len) {
      // partial dh[k] of gate block g: sum_i W[(g*H + i), k] * dG_h[g*H + i]
      const float* dv = dgs + g * H;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int q = 0; q < RW / 4; ++q) {
        const float4 d4 = *reinterpret_cast<const float4*>(dv + 4 * q);
        a0 = fmaf(wr[4 * q], d4.x, a0);
        a1 = fmaf(wr[4 * q + 1], d4.y, a1);
        a0 = fmaf(wr[4 * q + 2], d4.z, a0);
        a1 = fmaf(wr[4 * q + 3], d4.w, a1);
      }
#pragma unroll
      for (int q = 0; q < SWQ; ++q) {
        const float4 w4 = Ws[q * NT + tid];
        const float4 d4 = *reinterpret_cast<const float4*>(dv + RW + 4 * q);
        a0 = fmaf(w4.x, d4.x, a0);
        a1 = fmaf(w4.y, d4.y, a1);
        a0 = fmaf(w4.z, d4.z, a0);
        a1 = fmaf(w4.w, d4.w, a1);
      }
      part[tid] = a0 + a1;
    }
    __syncthreads();
  }
}

static size_t pf32_fwd_smem(int G) { return ((size_t)SWQ * G * FH * 4 + FH + (size_t)G * FH + FH) * sizeof(float); }
static size_t pf32_bwd_smem(int G) { return ((size_t)SWQ * G * FH * 4 + 2 * (size_t)G * FH) * sizeof(float); }

static bool pf32_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SLNLP_PERSIST_F32");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

// -1 = shape not supported (the caller falls back to the per-step kernels)
int rnn_layer_fwd_pf32(int mode, int T, int B, int H, int ndir, float* gates, const float* w_hh, const float* b_hh,
                       const int64_t* lengths, const float* h0, const float* c0, float* out, float* stash,
                       float* h_final, const slnlp_rnn_extras* ex, cudaStream_t s) {
  if (!pf32_enabled() || H != FH || T <= 1 || h0 || c0 || ((uintptr_t)w_hh & 15)) return -1;
  const bool cat = ex && ex->hfinal_cat;
  PF32Fwd p{T, B, ndir, gates, w_hh, b_hh, lengths, out, stash, h_final,
            cat ? (int64_t)H : (int64_t)B * H, cat ? (int64_t)ndir * H : (int64_t)H,
            ex ? ex->out_drop : nullptr, ex ? ex->p_drop : 0.f, ex ? ex->rng : nullptr, ex ? ex->site : 0u,
            ex ? ex->mask : nullptr};
  const dim3 grid(B, ndir);
  if (mode == SLNLP_MODE_LSTM) {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(rnn_pf32_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pf32_fwd_smem(4)); attr = true; }
    launch_pdl(rnn_pf32_fwd_kernel<4>, grid, dim3(4 * FH), pf32_fwd_smem(4), s, p);
  } else {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(rnn_pf32_fwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pf32_fwd_smem(3)); attr = true; }
    launch_pdl(rnn_pf32_fwd_kernel<3>, grid, dim3(3 * FH), pf32_fwd_smem(3), s, p);
  }
  SLNLP_LAUNCH_OK("rnn_layer_fwd(persistent fp32)");
  return 0;
}

int rnn_layer_bwd_pf32(int mode, int T, int B, int H, int ndir, float* gates, float* stash, const float* out,
                       const float* w_hh, const int64_t* lengths, const float* h0, const float* c0, const float* dout,
                       const float* dh_final, const float* dc_final, float* dh0, float* dc0,
                       const slnlp_rnn_extras* ex, cudaStream_t s) {
  if (!pf32_enabled() || H != FH || T <= 1 || h0 || c0 || dh0 || dc0 || ((uintptr_t)w_hh & 15)) return -1;
  const bool cat = ex && ex->hfinal_cat;
  const bool undrop = ex && ex->dout_dropped && ex->p_drop > 0.f;
  PF32Bwd p{T, B, ndir, gates, stash, out, w_hh, lengths, dout, dh_final, dc_final,
            cat ? (int64_t)H : (int64_t)B * H, cat ? (int64_t)ndir * H : (int64_t)H,
            undrop ? ex->p_drop : 0.f, undrop ? ex->rng : nullptr, undrop ? ex->site : 0u, undrop ? ex->mask : nullptr};
  const dim3 grid(B, ndir);
  if (mode == SLNLP_MODE_LSTM) {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(rnn_pf32_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pf32_bwd_smem(4)); attr = true; }
    launch_pdl(rnn_pf32_bwd_kernel<4>, grid, dim3(4 * FH), pf32_bwd_smem(4), s, p);
  } else {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(rnn_pf32_bwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pf32_bwd_smem(3)); attr = true; }
    launch_pdl(rnn_pf32_bwd_kernel<3>, grid, dim3(3 * FH), pf32_bwd_smem(3), s, p);
  }
  SLNLP_LAUNCH_OK("rnn_layer_bwd(persistent fp32)");
  return 0;
}

}  // namespace slnlp
