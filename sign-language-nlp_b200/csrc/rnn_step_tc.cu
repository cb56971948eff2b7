// K4, tensor-core path for any hidden size that is a multiple of 32 (H = 256, 512 of the
// reference grid; batch 50 ... 4096): one launch per timestep of
//
//   forward : acc[b, (g,u)] = sum_k h_{t-1}[b,k] W_hh[gH+u, k]   (tcgen05.mma kind::tf32, M = 128
//             sequences, N = G*32 gate columns of 32 hidden units, K = H streamed by TMA), then
//             in the epilogue: + hoisted x W_ih^T + b_hh, gate nonlinearities, cell update, length
//             freeze, activated gates stashed in place for BPTT;
//   backward: m[b,k] = sum_j dG_{t+1}[b,j] W_hh[j,k]             (W_hh consumed in place as an
//             MN-major operand), then the cell backward of step t in the epilogue, writing
//             d(pre-activations) in place - the A operand of the next launch.
//
// Swap of roles w.r.t. the W_hh-resident kernel (rnn_persistent.cu): here a TMEM lane is a
// SEQUENCE and the columns are (gate, unit), so one thread owns all gates of its sequence for the
// CTA's 32 units and the whole cell update happens in registers.  Operands stay fp32 in HBM/L2
// (W_hh: 0.75 - 4 MB, L2-resident across the T launches) and are read as TF32 through the same
// 4..8-stage TMA/mbarrier ring as gemm_tma.cu.  Used when precision == 1 and the persistent
// kernel does not cover the shape.
#include "tma.cuh"

namespace slnlp {

constexpr int SB = 128;                 // sequences per CTA (MMA M)
constexpr int SU = 32;                  // hidden units per CTA
constexpr int SK = 32;                  // k-tile (floats) = one 128-byte swizzle row
constexpr int S_THREADS = 192;
constexpr int SA_STAGE = SB * SK * 4;   // 16 KB

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}

struct TcFwd {
  int T, B, H, ndir, step;
  float* gates;
  const float* b_hh;
  const int64_t* lengths;
  const float* h0;
  const float* c0;
  float* out;
  float* stash;
  float* h_final;
};

// shared pipeline prologue: carve smem, init barriers, allocate TMEM
struct Ring {
  uint8_t *sA, *sB;
  uint64_t *full, *empty, *acc_full;
  uint32_t tmem;
};
template <int STAGES, int B_STAGE, int TCOLS>
__device__ __forceinline__ Ring ring_setup(uint8_t* smem_dyn, int warp, const CUtensorMap* m0, const CUtensorMap* m1,
                                           const CUtensorMap* m2) {
  Ring r;
  uint8_t* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  r.sA = base;
  r.sB = base + STAGES * SA_STAGE;
  r.full = reinterpret_cast<uint64_t*>(r.sB + STAGES * B_STAGE);
  r.empty = r.full + STAGES;
  r.acc_full = r.empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r.acc_full + 1);
  if (warp == 0 && elect_one()) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(m1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(m2) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&r.full[s], 1);
      mbar_init(&r.empty[s], 1);
    }
    mbar_init(r.acc_full, 1);
  }
  if (warp == 1) tmem_alloc(tmem_slot, TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  r.tmem = *tmem_slot;
  return r;
}

constexpr int F_STAGES = 4;

// grid (H/32, ceil(B/128), ndir), block 192.  mapH: out as [T][B][ndir*H]; mapH0: h0 as [ndir][B][H];
// mapW: w_hh as [ndir][G*H][H], box {32 k, 32 rows}.
template <int G>
__global__ void __launch_bounds__(S_THREADS) rnn_step_fwd_tc_kernel(const __grid_constant__ CUtensorMap mapH,
                                                                    const __grid_constant__ CUtensorMap mapH0,
                                                                    const __grid_constant__ CUtensorMap mapW, TcFwd p) {
  extern __shared__ uint8_t smem_dyn[];
  constexpr int B_STAGE = G * SU * SK * 4, NCOL = G * SU;
  const int warp = warp_uniform(), lane = threadIdx.x & 31;
  const int H = p.H, B = p.B, T = p.T;
  const int d = blockIdx.z, u0 = blockIdx.x * SU, b0 = blockIdx.y * SB;
  const int t = d == 0 ? p.step : T - 1 - p.step;
  const int tp = d == 0 ? t - 1 : t + 1;
  const bool has_prev = tp >= 0 && tp < T;
  const int nk = (has_prev || p.h0) ? H / SK : 0;   // zero initial state: the recurrent product is exactly 0
  Ring r = ring_setup<F_STAGES, B_STAGE, 128>(smem_dyn, warp, &mapH, &mapH0, &mapW);

  if (warp == 0) {
    if (elect_one()) {
      for (int i = 0; i < nk; ++i) {
        const int s = i % F_STAGES, k0 = i * SK;
        if (i >= F_STAGES) mbar_wait(&r.empty[s], (uint32_t)(i / F_STAGES - 1) & 1u);
        mbar_expect_tx(&r.full[s], SA_STAGE + B_STAGE);
        if (has_prev) tma_load_3d(r.sA + s * SA_STAGE, &mapH, &r.full[s], d * H + k0, b0, tp);
        else tma_load_3d(r.sA + s * SA_STAGE, &mapH0, &r.full[s], k0, b0, d);
#pragma unroll
        for (int g = 0; g < G; ++g) tma_load_3d(r.sB + s * B_STAGE + g * 4096, &mapW, &r.full[s], k0, g * H + u0, d);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_tf32(SB, NCOL, 0, 0);
    const uint64_t dA = make_desc_sw128(smem_u32(r.sA), 16, 1024, 2), dB = make_desc_sw128(smem_u32(r.sB), 16, 1024, 2);
    for (int i = 0; i < nk; ++i) {
      const int s = i % F_STAGES;
      mbar_wait(&r.full[s], (uint32_t)(i / F_STAGES) & 1u);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < SK / UMMA_K; ++kk)
          umma_tf32(r.tmem, dA + (uint64_t)((s * SA_STAGE + kk * 32) >> 4), dB + (uint64_t)((s * B_STAGE + kk * 32) >> 4),
                    idesc, (i > 0 || kk > 0) ? 1u : 0u);
        umma_commit(&r.empty[s]);
        if (i == nk - 1) umma_commit(r.acc_full);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int b = b0 + q * 32 + lane;
    const bool valid = b < B;
    const int len = (valid && p.lengths) ? (int)p.lengths[b] : T;
    const bool active = valid && t < len;
    const int bb = valid ? b : 0;
    const int64_t row = (int64_t)t * B + bb;
    float* gt = p.gates + (row * p.ndir + d) * G * H + u0;
    float* o = p.out + row * p.ndir * H + (int64_t)d * H + u0;
    float* st = p.stash + (row * p.ndir + d) * H + u0;
    const float* bh = p.b_hh + (int64_t)d * G * H + u0;
    const int64_t cidx = ((int64_t)d * B + bb) * H + u0;
    // predecessor state of this sequence (fp32, exact): c_{t-1} (LSTM) / h_{t-1} (GRU)
    const float* prev = nullptr;
    if (G == 4) prev = has_prev ? p.stash + (((int64_t)tp * B + bb) * p.ndir + d) * H + u0 : (p.c0 ? p.c0 + cidx : nullptr);
    else prev = has_prev ? p.out + ((int64_t)tp * B + bb) * p.ndir * H + (int64_t)d * H + u0 : (p.h0 ? p.h0 + cidx : nullptr);
    if (nk > 0) {
      mbar_wait(r.acc_full, 0);
      tc_fence_after();
    }
    const bool fin = active && p.h_final && (d == 0 ? t == len - 1 : t == 0);
#pragma unroll 1
    for (int c = 0; c < SU; c += 8) {
      float acc[G][8];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (nk > 0) {
          tmem_ld8(r.tmem + ((uint32_t)(q * 32) << 16) + g * SU + c, acc[g]);
        } else {
#pragma unroll
          for (int x = 0; x < 8; ++x) acc[g][x] = 0.f;
        }
      }
      if (!valid) continue;
      float hv[8], sv[8];
      if (!active) {
#pragma unroll
        for (int x = 0; x < 8; ++x) hv[x] = 0.f;
        st8(o + c, hv);
        st8(st + c, hv);
        continue;
      }
      float xg[G][8], bv[G][8], pv[8];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        ld8(gt + g * H + c, xg[g]);
        ld8(bh + g * H + c, bv[g]);
      }
      if (prev) {
        ld8(prev + c, pv);
      } else {
#pragma unroll
        for (int x = 0; x < 8; ++x) pv[x] = 0.f;
      }
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        if (G == 4) {
          const float gi = sigmoid_fast(xg[0][x] + acc[0][x] + bv[0][x]);
          const float gf = sigmoid_fast(xg[1][x] + acc[1][x] + bv[1][x]);
          const float gg = tanh_fast(xg[2][x] + acc[2][x] + bv[2][x]);
          const float go = sigmoid_fast(xg[G - 1][x] + acc[G - 1][x] + bv[G - 1][x]);
          const float cc = gf * pv[x] + gi * gg;
          hv[x] = go * tanh_fast(cc);
          sv[x] = cc;
          xg[0][x] = gi; xg[1][x] = gf; xg[2][x] = gg; xg[G - 1][x] = go;
        } else {
          const float hn = acc[2][x] + bv[2][x];
          const float gr = sigmoid_fast(xg[0][x] + acc[0][x] + bv[0][x]);
          const float gz = sigmoid_fast(xg[1][x] + acc[1][x] + bv[1][x]);
          const float gn = tanh_fast(xg[2][x] + gr * hn);
          hv[x] = (1.f - gz) * gn + gz * pv[x];
          sv[x] = hn;
          xg[0][x] = gr; xg[1][x] = gz; xg[2][x] = gn;
        }
      }
#pragma unroll
      for (int g = 0; g < G; ++g) st8(gt + g * H + c, xg[g]);
      st8(st + c, sv);
      st8(o + c, hv);
      if (fin) st8(p.h_final + cidx + c, hv);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(r.tmem, 128);
}

struct TcBwd {
  int T, B, H, ndir, step, final_only;
  float* gates;
  float* stash;
  const float* out;
  const int64_t* lengths;
  const float *h0, *c0, *dout, *dh_final, *dc_final;
  float *dh0, *dc0, *carry;
};

constexpr int B_STAGES = 8;

// grid (H/32, ceil(B/128), ndir), block 192.  mapG: gates as [T][B][ndir*G*H]; mapS: stash as
// [T][B][ndir*H] (GRU: the d(W_hn h) part of the reduction range lives there); mapW: w_hh as
// [ndir][G*H (j)][H (k)] read MN-major, box {32 k, 32 j}.
template <int G>
__global__ void __launch_bounds__(S_THREADS) rnn_step_bwd_tc_kernel(const __grid_constant__ CUtensorMap mapG,
                                                                    const __grid_constant__ CUtensorMap mapS,
                                                                    const __grid_constant__ CUtensorMap mapW, TcBwd p) {
  extern __shared__ uint8_t smem_dyn[];
  constexpr int B_STAGE = SU * SK * 4;   // one {32 k, 32 j} box
  const int warp = warp_uniform(), lane = threadIdx.x & 31;
  const int H = p.H, B = p.B, T = p.T, GH = G * p.H;
  const int d = blockIdx.z, u0 = blockIdx.x * SU, b0 = blockIdx.y * SB;
  const int t = p.final_only ? (d == 0 ? -1 : T) : (d == 0 ? T - 1 - p.step : p.step);
  const int tn = d == 0 ? t + 1 : t - 1;   // the step processed just before this one
  const bool has_next = tn >= 0 && tn < T;
  const int nk = has_next ? GH / SK : 0;
  Ring r = ring_setup<B_STAGES, B_STAGE, 32>(smem_dyn, warp, &mapG, &mapS, &mapW);

  if (warp == 0) {
    if (elect_one()) {
      for (int i = 0; i < nk; ++i) {
        const int s = i % B_STAGES, j0 = i * SK;
        if (i >= B_STAGES) mbar_wait(&r.empty[s], (uint32_t)(i / B_STAGES - 1) & 1u);
        mbar_expect_tx(&r.full[s], SA_STAGE + B_STAGE);
        if (G == 4 || j0 < 2 * H) tma_load_3d(r.sA + s * SA_STAGE, &mapG, &r.full[s], d * GH + j0, b0, tn);
        else tma_load_3d(r.sA + s * SA_STAGE, &mapS, &r.full[s], d * H + (j0 - 2 * H), b0, tn);
        tma_load_3d(r.sB + s * B_STAGE, &mapW, &r.full[s], u0, j0, d);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_tf32(SB, SU, 0, 1);
    const uint64_t dA = make_desc_sw128(smem_u32(r.sA), 16, 1024, 2), dB = make_desc_sw128(smem_u32(r.sB), 4096, 512, 1);
    for (int i = 0; i < nk; ++i) {
      const int s = i % B_STAGES;
      mbar_wait(&r.full[s], (uint32_t)(i / B_STAGES) & 1u);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < SK / UMMA_K; ++kk)
          umma_tf32(r.tmem, dA + (uint64_t)((s * SA_STAGE + kk * 32) >> 4), dB + (uint64_t)((s * B_STAGE + kk * 1024) >> 4),
                    idesc, (i > 0 || kk > 0) ? 1u : 0u);
        umma_commit(&r.empty[s]);
        if (i == nk - 1) umma_commit(r.acc_full);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int b = b0 + q * 32 + lane;
    const bool valid = b < B;
    const int bb = valid ? b : 0;
    const int len = (valid && p.lengths) ? (int)p.lengths[b] : T;
    const int64_t cidx = ((int64_t)d * B + bb) * H + u0;
    if (nk > 0) {
      mbar_wait(r.acc_full, 0);
      tc_fence_after();
    }
    const int tt = p.final_only ? 0 : t;
    const int64_t row = ((int64_t)tt * B + bb) * p.ndir + d;
    float* gt = p.gates + row * GH + u0;
    float* st = p.stash + row * H + u0;
    const int tp = d == 0 ? t - 1 : t + 1;   // forward-time predecessor
    const bool has_prev = tp >= 0 && tp < T;
    const bool inject = d == 0 ? t == len - 1 : t == 0;
#pragma unroll 1
    for (int c = 0; c < SU; c += 8) {
      float m[8];
      if (nk > 0) {
        tmem_ld8(r.tmem + ((uint32_t)(q * 32) << 16) + c, m);
      } else {
#pragma unroll
        for (int x = 0; x < 8; ++x) m[x] = 0.f;
      }
      if (!valid) continue;
      float cr[8];
      ld8(p.carry + cidx + c, cr);
      if (p.final_only) {
        if (G == 4) {
          if (p.dh0) st8(p.dh0 + cidx + c, m);
          if (p.dc0) st8(p.dc0 + cidx + c, cr);
        } else if (p.dh0) {
#pragma unroll
          for (int x = 0; x < 8; ++x) m[x] += cr[x];
          st8(p.dh0 + cidx + c, m);
        }
        continue;
      }
      float z[8];
#pragma unroll
      for (int x = 0; x < 8; ++x) z[x] = 0.f;
      if (t >= len) {
#pragma unroll
        for (int g = 0; g < G; ++g) st8(gt + g * H + c, z);
        if (G == 3) st8(st + c, z);
        continue;
      }
      float dh[8], gv[G][8], sv[8], pv[8], fin[8], fc[8];
      if (p.dout) ld8(p.dout + ((int64_t)t * B + b) * p.ndir * H + (int64_t)d * H + u0 + c, dh);
      else {
#pragma unroll
        for (int x = 0; x < 8; ++x) dh[x] = 0.f;
      }
#pragma unroll
      for (int g = 0; g < G; ++g) ld8(gt + g * H + c, gv[g]);
      ld8(st + c, sv);
      {
        const float* pp;
        if (G == 4) pp = has_prev ? p.stash + (((int64_t)tp * B + b) * p.ndir + d) * H + u0 + c : (p.c0 ? p.c0 + cidx + c : nullptr);
        else pp = has_prev ? p.out + ((int64_t)tp * B + b) * p.ndir * H + (int64_t)d * H + u0 + c : (p.h0 ? p.h0 + cidx + c : nullptr);
        if (pp) ld8(pp, pv);
        else {
#pragma unroll
          for (int x = 0; x < 8; ++x) pv[x] = 0.f;
        }
      }
      if (inject && p.dh_final) ld8(p.dh_final + cidx + c, fin);
      else {
#pragma unroll
        for (int x = 0; x < 8; ++x) fin[x] = 0.f;
      }
      if (G == 4 && inject && p.dc_final) ld8(p.dc_final + cidx + c, fc);
      else {
#pragma unroll
        for (int x = 0; x < 8; ++x) fc[x] = 0.f;
      }
      float dst[8];
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        if (G == 4) {
          const float dhx = dh[x] + (inject ? fin[x] : m[x]);
          const float dc_in = inject ? fc[x] : cr[x];
          const float gi = gv[0][x], gf = gv[1][x], gg = gv[2][x], go = gv[G - 1][x];
          const float tc = tanh_fast(sv[x]);
          const float dc = dhx * go * (1.f - tc * tc) + dc_in;
          gv[0][x] = dc * gg * gi * (1.f - gi);
          gv[1][x] = dc * pv[x] * gf * (1.f - gf);
          gv[2][x] = dc * gi * (1.f - gg * gg);
          gv[G - 1][x] = dhx * tc * go * (1.f - go);
          cr[x] = dc * gf;
          dst[x] = 0.f;
        } else {
          const float dhx = dh[x] + (inject ? fin[x] : m[x] + cr[x]);
          const float gr = gv[0][x], gz = gv[1][x], gn = gv[2][x], hn = sv[x];
          const float da_n = dhx * (1.f - gz) * (1.f - gn * gn);
          gv[0][x] = da_n * hn * gr * (1.f - gr);
          gv[1][x] = dhx * (pv[x] - gn) * gz * (1.f - gz);
          gv[2][x] = da_n;
          dst[x] = da_n * gr;
          cr[x] = dhx * gz;
        }
      }
#pragma unroll
      for (int g = 0; g < G; ++g) st8(gt + g * H + c, gv[g]);
      if (G == 3) st8(st + c, dst);
      st8(p.carry + cidx + c, cr);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(r.tmem, 32);
}

static bool step_tc_supported(int B, int H, const void* a, const void* b, const void* c) {
  return H % 32 == 0 && H >= 64 && H <= 4096 && B >= 1 && encode_fn() != nullptr &&
         (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) == 0;
}

// returns -1 when the shape / alignment is not supported (the caller falls back to the fp32 step kernels)
int rnn_layer_fwd_tcstep(int mode, int T, int B, int H, int ndir, float* gates, const float* w_hh, const float* b_hh,
                         const int64_t* lengths, const float* h0, const float* c0, float* out, float* stash,
                         float* h_final, cudaStream_t s) {
  const int G = mode == SLNLP_MODE_LSTM ? 4 : 3;
  if (!step_tc_supported(B, H, gates, out, w_hh) || (((uintptr_t)stash | (uintptr_t)b_hh) & 15)) return -1;
  if ((h0 && ((uintptr_t)h0 & 15)) || (c0 && ((uintptr_t)c0 & 15)) || (h_final && ((uintptr_t)h_final & 15))) return -1;
  CUtensorMap mapH, mapH0, mapW;
  if (!tensor_map3(out, (uint64_t)ndir * H, B, T, (uint64_t)ndir * H, (uint64_t)B * ndir * H, SB, false, &mapH)) return -1;
  if (h0) {
    if (!tensor_map3(h0, H, B, ndir, H, (uint64_t)B * H, SB, false, &mapH0)) return -1;
  } else {
    mapH0 = mapH;
  }
  if (!tensor_map3(w_hh, H, (uint64_t)G * H, ndir, H, (uint64_t)G * H * H, SU, false, &mapW)) return -1;
  TcFwd p{T, B, H, ndir, 0, gates, b_hh, lengths, h0, c0, out, stash, h_final};
  dim3 grid(H / SU, ceil_div(B, SB), ndir);
  const size_t sm = F_STAGES * (SA_STAGE + G * SU * SK * 4) + (2 * F_STAGES + 1) * 8 + 16 + 1024;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(rnn_step_fwd_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(F_STAGES * (SA_STAGE + 4 * SU * SK * 4) + 2048));
    cudaFuncSetAttribute(rnn_step_fwd_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)(F_STAGES * (SA_STAGE + 4 * SU * SK * 4) + 2048));
    attr = true;
  }
  for (int step = 0; step < T; ++step) {
    p.step = step;
    if (G == 4) rnn_step_fwd_tc_kernel<4><<<grid, S_THREADS, sm, s>>>(mapH, mapH0, mapW, p);
    else rnn_step_fwd_tc_kernel<3><<<grid, S_THREADS, sm, s>>>(mapH, mapH0, mapW, p);
  }
  note_launches(T);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("rnn_layer_fwd(tc step): launch failed: %s", cudaGetErrorString(e));
  return 0;
}

int rnn_layer_bwd_tcstep(int mode, int T, int B, int H, int ndir, float* gates, float* stash, const float* out,
                         const float* w_hh, const int64_t* lengths, const float* h0, const float* c0,
                         const float* dout, const float* dh_final, const float* dc_final, float* dh0, float* dc0,
                         float* carry, cudaStream_t s) {
  const int G = mode == SLNLP_MODE_LSTM ? 4 : 3;
  if (!step_tc_supported(B, H, gates, stash, w_hh) || (((uintptr_t)out | (uintptr_t)carry) & 15)) return -1;
  const void* opt[] = {h0, c0, dout, dh_final, dc_final, dh0, dc0};
  for (const void* q : opt)
    if (q && ((uintptr_t)q & 15)) return -1;
  CUtensorMap mapG, mapS, mapW;
  if (!tensor_map3(gates, (uint64_t)ndir * G * H, B, T, (uint64_t)ndir * G * H, (uint64_t)B * ndir * G * H, SB, false, &mapG)) return -1;
  if (!tensor_map3(stash, (uint64_t)ndir * H, B, T, (uint64_t)ndir * H, (uint64_t)B * ndir * H, SB, false, &mapS)) return -1;
  if (!tensor_map3(w_hh, H, (uint64_t)G * H, ndir, H, (uint64_t)G * H * H, SK, true, &mapW)) return -1;
  TcBwd p{T, B, H, ndir, 0, 0, gates, stash, out, lengths, h0, c0, dout, dh_final, dc_final, dh0, dc0, carry};
  dim3 grid(H / SU, ceil_div(B, SB), ndir);
  const size_t sm = B_STAGES * (SA_STAGE + SU * SK * 4) + (2 * B_STAGES + 1) * 8 + 16 + 1024;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(rnn_step_bwd_tc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    cudaFuncSetAttribute(rnn_step_bwd_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    attr = true;
  }
  // the carry buffer starts at zero (the fp32 step kernels only read it after writing it)
  cudaMemsetAsync(carry, 0, sizeof(float) * (size_t)ndir * B * H, s);
  auto go = [&]() {
    if (G == 4) rnn_step_bwd_tc_kernel<4><<<grid, S_THREADS, sm, s>>>(mapG, mapS, mapW, p);
    else rnn_step_bwd_tc_kernel<3><<<grid, S_THREADS, sm, s>>>(mapG, mapS, mapW, p);
  };
  for (int step = 0; step < T; ++step) {
    p.step = step;
    go();
  }
  int n = T;
  if (dh0 || dc0) {
    p.final_only = 1;
    go();
    ++n;
  }
  note_launches(n);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("rnn_layer_bwd(tc step): launch failed: %s", cudaGetErrorString(e));
  return 0;
}

}  // namespace slnlp
