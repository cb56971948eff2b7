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
constexpr int S_THREADS = 320;          // producer, MMA, 8 epilogue warps (two per TMEM lane quadrant, 16 units each)
constexpr int SUH = SU / 2;             // hidden units per epilogue thread

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
__device__ __forceinline__ void st8_bf16(__nv_bfloat16* p, const float (&v)[8]) { *reinterpret_cast<uint4*>(p) = pack8_bf16(v); }

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
  __nv_bfloat16* out_bf;   // BF kernels: bf16 copy of out (the next step's A operand, and dW_hh's B operand)
};

// shared pipeline prologue: carve smem, init barriers, allocate TMEM
struct Ring {
  uint8_t *sA, *sB;
  uint64_t *full, *empty, *acc_full;
  uint32_t tmem;
};
template <int STAGES, int SA_STAGE, int B_STAGE, int TCOLS>
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

// grid (H/32, ceil(B/128), ndir), block 320.  mapH: out as [T][B][ndir*H]; mapH0: h0 as [ndir][B][H];
// mapW: w_hh as [ndir][G*H][H], box {32 k, 32 rows}.  AROWS = rows the A box actually loads (64 when
// the batch fits: the MMA still reads 128 rows, the upper 64 are whatever shared memory holds and only
// reach accumulator rows nobody reads) - half the bytes per stage buys twice the stages in flight.
// BF: operands are bf16 copies (out_bf, a bf16 W_hh; tcgen05 kind::f16, 64-element k-blocks - the same 128-byte rows,
// half the L2 -> shared-memory bytes per flop) and the epilogue also writes h_t as bf16.
template <int G, int F_STAGES, int AROWS, bool BF>
__global__ void __launch_bounds__(S_THREADS) rnn_step_fwd_tc_kernel(const __grid_constant__ CUtensorMap mapH,
                                                                    const __grid_constant__ CUtensorMap mapH0,
                                                                    const __grid_constant__ CUtensorMap mapW, TcFwd p) {
  extern __shared__ uint8_t smem_dyn[];
  constexpr int SA_STAGE = AROWS * SK * 4, B_STAGE = G * SU * SK * 4, NCOL = G * SU;
  const int warp = warp_uniform(), lane = threadIdx.x & 31;
  const int H = p.H, B = p.B, T = p.T;
  const int d = blockIdx.z, u0 = blockIdx.x * SU, b0 = blockIdx.y * SB;
  const int t = d == 0 ? p.step : T - 1 - p.step;
  const int tp = d == 0 ? t - 1 : t + 1;
  const bool has_prev = tp >= 0 && tp < T;
  constexpr int KE = BF ? 64 : SK;                  // elements per k-block (one 128-byte row)
  const int nk = (has_prev || p.h0) ? H / KE : 0;   // zero initial state: the recurrent product is exactly 0
  Ring r = ring_setup<F_STAGES, SA_STAGE, B_STAGE, 128>(smem_dyn, warp, &mapH, &mapH0, &mapW);

  // the next timestep's launch may start now: its prologue and its W_hh tiles overlap this step
  pdl_launch_dependents();
  if (warp == 0) {
    if (elect_one()) {
      // first ring round: W_hh tiles do not depend on the previous step - issue them, THEN wait for
      // the previous launch (h_{t-1} must be complete in HBM before the A tiles are read)
      const int first = nk < F_STAGES ? nk : F_STAGES;
      for (int i = 0; i < first; ++i) {
        mbar_expect_tx(&r.full[i], SA_STAGE + B_STAGE);
#pragma unroll
        for (int g = 0; g < G; ++g) tma_load_3d(r.sB + i * B_STAGE + g * 4096, &mapW, &r.full[i], i * KE, g * H + u0, d);
      }
      pdl_wait();
      for (int i = 0; i < first; ++i) {
        if (has_prev) tma_load_3d(r.sA + i * SA_STAGE, &mapH, &r.full[i], d * H + i * KE, b0, tp);
        else tma_load_3d(r.sA + i * SA_STAGE, &mapH0, &r.full[i], i * KE, b0, d);
      }
      for (int i = first; i < nk; ++i) {
        const int s = i % F_STAGES, k0 = i * KE;
        mbar_wait(&r.empty[s], (uint32_t)(i / F_STAGES - 1) & 1u);
        mbar_expect_tx(&r.full[s], SA_STAGE + B_STAGE);
        if (has_prev) tma_load_3d(r.sA + s * SA_STAGE, &mapH, &r.full[s], d * H + k0, b0, tp);
        else tma_load_3d(r.sA + s * SA_STAGE, &mapH0, &r.full[s], k0, b0, d);
#pragma unroll
        for (int g = 0; g < G; ++g) tma_load_3d(r.sB + s * B_STAGE + g * 4096, &mapW, &r.full[s], k0, g * H + u0, d);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = BF ? make_idesc_bf16(SB, NCOL, 0, 0) : make_idesc_tf32(SB, NCOL, 0, 0);
    const uint64_t dA = make_desc_sw128(smem_u32(r.sA), 16, 1024, 2), dB = make_desc_sw128(smem_u32(r.sB), 16, 1024, 2);
    for (int i = 0; i < nk; ++i) {
      const int s = i % F_STAGES;
      mbar_wait(&r.full[s], (uint32_t)(i / F_STAGES) & 1u);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {   // four MMAs of 32 bytes of K each (tf32: K = 8, bf16: K = 16)
          const uint64_t a = dA + (uint64_t)((s * SA_STAGE + kk * 32) >> 4), b = dB + (uint64_t)((s * B_STAGE + kk * 32) >> 4);
          if (BF) umma_bf16(r.tmem, a, b, idesc, (i > 0 || kk > 0) ? 1u : 0u);
          else umma_tf32(r.tmem, a, b, idesc, (i > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&r.empty[s]);
        if (i == nk - 1) umma_commit(r.acc_full);
      }
      __syncwarp();
    }
  } else {
    pdl_wait();   // everything below reads what the previous launch wrote
    const int q = warp & 3, uh = ((warp - 2) >> 2) * SUH;   // lane quadrant, first unit of this thread's half
    const int b = b0 + q * 32 + lane;
    const bool valid = b < B;
    const int len = (valid && p.lengths) ? (int)p.lengths[b] : T;
    const bool active = valid && t < len;
    const int bb = valid ? b : 0;
    const int64_t row = (int64_t)t * B + bb;
    float* gt = p.gates + (row * p.ndir + d) * G * H + u0 + uh;
    float* o = p.out + row * p.ndir * H + (int64_t)d * H + u0 + uh;
    float* st = p.stash + (row * p.ndir + d) * H + u0 + uh;
    __nv_bfloat16* ob = BF ? p.out_bf + row * p.ndir * H + (int64_t)d * H + u0 + uh : nullptr;
    const float* bh = p.b_hh + (int64_t)d * G * H + u0 + uh;
    const int64_t cidx = ((int64_t)d * B + bb) * H + u0 + uh;
    // predecessor state of this sequence (fp32, exact): c_{t-1} (LSTM) / h_{t-1} (GRU)
    const float* prev = nullptr;
    if (G == 4) prev = has_prev ? p.stash + (((int64_t)tp * B + bb) * p.ndir + d) * H + u0 + uh : (p.c0 ? p.c0 + cidx : nullptr);
    else prev = has_prev ? p.out + ((int64_t)tp * B + bb) * p.ndir * H + (int64_t)d * H + u0 + uh : (p.h0 ? p.h0 + cidx : nullptr);
    // every operand of the cell update is fetched BEFORE waiting for the accumulator, so the loads
    // overlap the TMA/MMA ring instead of extending the step by an L2 round trip per chunk
    float xg[G][SUH], pv[SUH];
#pragma unroll
    for (int c = 0; c < SUH; c += 8) {
      float tmp[8];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (active) {
          ld8(gt + g * H + c, tmp);
        } else {
#pragma unroll
          for (int x = 0; x < 8; ++x) tmp[x] = 0.f;
        }
#pragma unroll
        for (int x = 0; x < 8; ++x) xg[g][c + x] = tmp[x];
      }
      if (active && prev) {
        ld8(prev + c, tmp);
      } else {
#pragma unroll
        for (int x = 0; x < 8; ++x) tmp[x] = 0.f;
      }
#pragma unroll
      for (int x = 0; x < 8; ++x) pv[c + x] = tmp[x];
    }
    if (nk > 0) {
      mbar_wait(r.acc_full, 0);
      tc_fence_after();
    }
    const bool fin = active && p.h_final && (d == 0 ? t == len - 1 : t == 0);
#pragma unroll
    for (int c = 0; c < SUH; c += 8) {
      float acc[G][8];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        if (nk > 0) {
          tmem_ld8(r.tmem + ((uint32_t)(q * 32) << 16) + g * SU + uh + c, acc[g]);
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
        if (BF) st8_bf16(ob + c, hv);
        continue;
      }
      float bv[G][8], go_[G][8];
#pragma unroll
      for (int g = 0; g < G; ++g) ld8(bh + g * H + c, bv[g]);
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        if (G == 4) {
          const float gi = sigmoid_fast(xg[0][c + x] + acc[0][x] + bv[0][x]);
          const float gf = sigmoid_fast(xg[1][c + x] + acc[1][x] + bv[1][x]);
          const float gg = tanh_fast(xg[2][c + x] + acc[2][x] + bv[2][x]);
          const float go = sigmoid_fast(xg[G - 1][c + x] + acc[G - 1][x] + bv[G - 1][x]);
          const float cc = gf * pv[c + x] + gi * gg;
          hv[x] = go * tanh_fast(cc);
          sv[x] = cc;
          go_[0][x] = gi; go_[1][x] = gf; go_[2][x] = gg; go_[G - 1][x] = go;
        } else {
          const float hn = acc[2][x] + bv[2][x];
          const float gr = sigmoid_fast(xg[0][c + x] + acc[0][x] + bv[0][x]);
          const float gz = sigmoid_fast(xg[1][c + x] + acc[1][x] + bv[1][x]);
          const float gn = tanh_fast(xg[2][c + x] + gr * hn);
          hv[x] = (1.f - gz) * gn + gz * pv[c + x];
          sv[x] = hn;
          go_[0][x] = gr; go_[1][x] = gz; go_[2][x] = gn;
        }
      }
#pragma unroll
      for (int g = 0; g < G; ++g) st8(gt + g * H + c, go_[g]);
      st8(st + c, sv);
      st8(o + c, hv);
      if (BF) st8_bf16(ob + c, hv);
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
  __nv_bfloat16* dg_bf;    // BF kernels: bf16 copy of d(pre-activations) (the next step's A operand; dX / dW GEMMs)
};

// grid (H/32, ceil(B/128), ndir), block 192.  mapG: gates as [T][B][ndir*G*H]; mapS: stash as
// [T][B][ndir*H] (GRU: the d(W_hn h) part of the reduction range lives there); mapW: w_hh as
// [ndir][G*H (j)][H (k)] read MN-major, box {32 k, 32 j}.
// NCH = 32-unit output chunks per CTA: 1 at small batch (more CTAs share the W_hh stream), 4 at large
// batch (the dG tile, re-read by every unit tile of the same sequences, is fetched 4x less often)
// BF (LSTM): A = the bf16 copy of dG (mapG), B = a bf16 copy of W_hh^T as [ndir][H][G*H] (K-major, mapW); the epilogue
// also writes d(pre-activations) as bf16.
template <int G, int B_STAGES, int AROWS, int NCH, bool BF>
__global__ void __launch_bounds__(S_THREADS) rnn_step_bwd_tc_kernel(const __grid_constant__ CUtensorMap mapG,
                                                                    const __grid_constant__ CUtensorMap mapS,
                                                                    const __grid_constant__ CUtensorMap mapW, TcBwd p) {
  extern __shared__ uint8_t smem_dyn[];
  constexpr int SA_STAGE = AROWS * SK * 4, B_STAGE = NCH * SU * SK * 4;   // B: NCH {32 k, 32 j} boxes
  const int warp = warp_uniform(), lane = threadIdx.x & 31;
  const int H = p.H, B = p.B, T = p.T, GH = G * p.H;
  const int d = blockIdx.z, u0 = blockIdx.x * SU * NCH, b0 = blockIdx.y * SB;
  const int t = p.final_only ? (d == 0 ? -1 : T) : (d == 0 ? T - 1 - p.step : p.step);
  const int tn = d == 0 ? t + 1 : t - 1;   // the step processed just before this one
  const bool has_next = tn >= 0 && tn < T;
  constexpr int KE = BF ? 64 : SK;
  const int nk = has_next ? GH / KE : 0;
  Ring r = ring_setup<B_STAGES, SA_STAGE, B_STAGE, NCH * SU>(smem_dyn, warp, &mapG, &mapS, &mapW);

  pdl_launch_dependents();
  if (warp == 0) {
    if (elect_one()) {
      auto load_a = [&](int s, int j0) {
        if (BF || G == 4 || j0 < 2 * H) tma_load_3d(r.sA + s * SA_STAGE, &mapG, &r.full[s], d * GH + j0, b0, tn);
        else tma_load_3d(r.sA + s * SA_STAGE, &mapS, &r.full[s], d * H + (j0 - 2 * H), b0, tn);
      };
      const int first = nk < B_STAGES ? nk : B_STAGES;
      for (int i = 0; i < first; ++i) {
        mbar_expect_tx(&r.full[i], SA_STAGE + B_STAGE);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          if (BF) tma_load_3d(r.sB + i * B_STAGE + c * 4096, &mapW, &r.full[i], i * KE, u0 + c * SU, d);
          else tma_load_3d(r.sB + i * B_STAGE + c * 4096, &mapW, &r.full[i], u0 + c * SU, i * KE, d);
        }
      }
      pdl_wait();   // dG of the step processed just before must be complete in HBM
      for (int i = 0; i < first; ++i) load_a(i, i * KE);
      for (int i = first; i < nk; ++i) {
        const int s = i % B_STAGES, j0 = i * KE;
        mbar_wait(&r.empty[s], (uint32_t)(i / B_STAGES - 1) & 1u);
        mbar_expect_tx(&r.full[s], SA_STAGE + B_STAGE);
        load_a(s, j0);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          if (BF) tma_load_3d(r.sB + s * B_STAGE + c * 4096, &mapW, &r.full[s], j0, u0 + c * SU, d);
          else tma_load_3d(r.sB + s * B_STAGE + c * 4096, &mapW, &r.full[s], u0 + c * SU, j0, d);
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = BF ? make_idesc_bf16(SB, NCH * SU, 0, 0) : make_idesc_tf32(SB, NCH * SU, 0, 1);
    const uint64_t dA = make_desc_sw128(smem_u32(r.sA), 16, 1024, 2);
    const uint64_t dB = BF ? make_desc_sw128(smem_u32(r.sB), 16, 1024, 2) : make_desc_sw128(smem_u32(r.sB), 4096, 512, 1);
    for (int i = 0; i < nk; ++i) {
      const int s = i % B_STAGES;
      mbar_wait(&r.full[s], (uint32_t)(i / B_STAGES) & 1u);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t a = dA + (uint64_t)((s * SA_STAGE + kk * 32) >> 4);
          if (BF) umma_bf16(r.tmem, a, dB + (uint64_t)((s * B_STAGE + kk * 32) >> 4), idesc, (i > 0 || kk > 0) ? 1u : 0u);
          else umma_tf32(r.tmem, a, dB + (uint64_t)((s * B_STAGE + kk * 1024) >> 4), idesc, (i > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&r.empty[s]);
        if (i == nk - 1) umma_commit(r.acc_full);
      }
      __syncwarp();
    }
  } else {
    pdl_wait();
    for (int ch = 0; ch < NCH; ++ch) {
    const int ub = u0 + ch * SU;
    const int q = warp & 3, uh = ((warp - 2) >> 2) * SUH;
    const int b = b0 + q * 32 + lane;
    const bool valid = b < B;
    const int bb = valid ? b : 0;
    const int len = (valid && p.lengths) ? (int)p.lengths[b] : T;
    const int64_t cidx = ((int64_t)d * B + bb) * H + ub + uh;
    const int tt = p.final_only ? 0 : t;
    const int64_t row = ((int64_t)tt * B + bb) * p.ndir + d;
    float* gt = p.gates + row * GH + ub + uh;
    float* st = p.stash + row * H + ub + uh;
    __nv_bfloat16* gb = BF ? p.dg_bf + row * GH + ub + uh : nullptr;
    const int tp = d == 0 ? t - 1 : t + 1;   // forward-time predecessor
    const bool has_prev = tp >= 0 && tp < T;
    const bool inject = d == 0 ? t == len - 1 : t == 0;
    const bool live = valid && !p.final_only && t < len;   // this thread has a cell backward to do
    // operands of the cell backward, fetched before waiting for the accumulator
    float gv[G][SUH], sv[SUH], pv[SUH], dh[SUH], cr[SUH];
    {
      const float* pp = nullptr;
      if (live) {
        if (G == 4) pp = has_prev ? p.stash + (((int64_t)tp * B + b) * p.ndir + d) * H + ub + uh : (p.c0 ? p.c0 + cidx : nullptr);
        else pp = has_prev ? p.out + ((int64_t)tp * B + b) * p.ndir * H + (int64_t)d * H + ub + uh : (p.h0 ? p.h0 + cidx : nullptr);
      }
      const float* dop = (live && p.dout) ? p.dout + ((int64_t)t * B + b) * p.ndir * H + (int64_t)d * H + ub + uh : nullptr;
#pragma unroll
      for (int c = 0; c < SUH; c += 8) {
        float tmp[8];
        auto fetch = [&](const float* src, float* dst) {
          if (src) {
            ld8(src + c, tmp);
          } else {
#pragma unroll
            for (int x = 0; x < 8; ++x) tmp[x] = 0.f;
          }
#pragma unroll
          for (int x = 0; x < 8; ++x) dst[c + x] = tmp[x];
        };
#pragma unroll
        for (int g = 0; g < G; ++g) fetch(live ? gt + g * H : nullptr, gv[g]);
        fetch(live ? st : nullptr, sv);
        fetch(pp, pv);
        fetch(dop, dh);
        fetch(valid ? p.carry + cidx : nullptr, cr);
        if (live && inject) {   // final-state gradients enter here instead of the recurrent ones
          float fin[8];
          if (p.dh_final) {
            ld8(p.dh_final + cidx + c, fin);
#pragma unroll
            for (int x = 0; x < 8; ++x) dh[c + x] += fin[x];
          }
          if (G == 4) {
            if (p.dc_final) {
              ld8(p.dc_final + cidx + c, fin);
            } else {
#pragma unroll
              for (int x = 0; x < 8; ++x) fin[x] = 0.f;
            }
#pragma unroll
            for (int x = 0; x < 8; ++x) cr[c + x] = fin[x];   // dc_in of the injection step
          } else {
#pragma unroll
            for (int x = 0; x < 8; ++x) cr[c + x] = 0.f;
          }
        }
      }
    }
    if (nk > 0 && ch == 0) {
      mbar_wait(r.acc_full, 0);
      tc_fence_after();
    }
#pragma unroll
    for (int c = 0; c < SUH; c += 8) {
      float m[8];
      if (nk > 0) {
        tmem_ld8(r.tmem + ((uint32_t)(q * 32) << 16) + ch * SU + uh + c, m);
      } else {
#pragma unroll
        for (int x = 0; x < 8; ++x) m[x] = 0.f;
      }
      if (!valid) continue;
      if (p.final_only) {
        float o8[8];
        if (G == 4) {
          if (p.dh0) st8(p.dh0 + cidx + c, m);
#pragma unroll
          for (int x = 0; x < 8; ++x) o8[x] = cr[c + x];
          if (p.dc0) st8(p.dc0 + cidx + c, o8);
        } else if (p.dh0) {
#pragma unroll
          for (int x = 0; x < 8; ++x) o8[x] = m[x] + cr[c + x];
          st8(p.dh0 + cidx + c, o8);
        }
        continue;
      }
      float og[G][8], dst[8], oc[8];
      if (!live) {   // t >= len: zero gradients for the hoisted dW / dx GEMMs, carry untouched
#pragma unroll
        for (int x = 0; x < 8; ++x) dst[x] = 0.f;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          st8(gt + g * H + c, dst);
          if (BF) st8_bf16(gb + g * H + c, dst);
        }
        if (G == 3) st8(st + c, dst);
        continue;
      }
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        const int i = c + x;
        if (G == 4) {
          const float dhx = dh[i] + (inject ? 0.f : m[x]);
          const float gi = gv[0][i], gf = gv[1][i], gg = gv[2][i], go = gv[G - 1][i];
          const float tc = tanh_fast(sv[i]);
          const float dc = dhx * go * (1.f - tc * tc) + cr[i];
          og[0][x] = dc * gg * gi * (1.f - gi);
          og[1][x] = dc * pv[i] * gf * (1.f - gf);
          og[2][x] = dc * gi * (1.f - gg * gg);
          og[G - 1][x] = dhx * tc * go * (1.f - go);
          oc[x] = dc * gf;
          dst[x] = 0.f;
        } else {
          const float dhx = dh[i] + (inject ? 0.f : m[x] + cr[i]);
          const float gr = gv[0][i], gz = gv[1][i], gn = gv[2][i], hn = sv[i];
          const float da_n = dhx * (1.f - gz) * (1.f - gn * gn);
          og[0][x] = da_n * hn * gr * (1.f - gr);
          og[1][x] = dhx * (pv[i] - gn) * gz * (1.f - gz);
          og[2][x] = da_n;
          dst[x] = da_n * gr;
          oc[x] = dhx * gz;
        }
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        st8(gt + g * H + c, og[g]);
        if (BF) st8_bf16(gb + g * H + c, og[g]);
      }
      if (G == 3) st8(st + c, dst);
      st8(p.carry + cidx + c, oc);
    }
  }   // chunk loop
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(r.tmem, NCH * SU);
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
  const bool small = B <= 64;            // the A box loads 64 rows, twice the ring depth
  const uint32_t arows = small ? 64 : SB;
  if (!tensor_map3(out, (uint64_t)ndir * H, B, T, (uint64_t)ndir * H, (uint64_t)B * ndir * H, arows, false, &mapH)) return -1;
  if (h0) {
    if (!tensor_map3(h0, H, B, ndir, H, (uint64_t)B * H, arows, false, &mapH0)) return -1;
  } else {
    mapH0 = mapH;
  }
  if (!tensor_map3(w_hh, H, (uint64_t)G * H, ndir, H, (uint64_t)G * H * H, SU, false, &mapW)) return -1;
  TcFwd p{T, B, H, ndir, 0, gates, b_hh, lengths, h0, c0, out, stash, h_final, nullptr};
  dim3 grid(H / SU, ceil_div(B, SB), ndir);
  auto run = [&](auto kernel, int stages, int arows) {
    const size_t sm = (size_t)stages * (arows * SK * 4 + G * SU * SK * 4) + (2 * stages + 1) * 8 + 16 + 1024;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    for (int step = 0; step < T; ++step) {
      p.step = step;
      launch_pdl(kernel, grid, dim3(S_THREADS), sm, s, mapH, mapH0, mapW, p);
    }
  };
  if (small) {
    if (G == 4) run(rnn_step_fwd_tc_kernel<4, 8, 64, false>, 8, 64); else run(rnn_step_fwd_tc_kernel<3, 8, 64, false>, 8, 64);
  } else {
    if (G == 4) run(rnn_step_fwd_tc_kernel<4, 4, 128, false>, 4, 128); else run(rnn_step_fwd_tc_kernel<3, 4, 128, false>, 4, 128);
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
  const bool small = B <= 64;
  const uint32_t arows = small ? 64 : SB;
  if (!tensor_map3(gates, (uint64_t)ndir * G * H, B, T, (uint64_t)ndir * G * H, (uint64_t)B * ndir * G * H, arows, false, &mapG)) return -1;
  if (!tensor_map3(stash, (uint64_t)ndir * H, B, T, (uint64_t)ndir * H, (uint64_t)B * ndir * H, arows, false, &mapS)) return -1;
  if (!tensor_map3(w_hh, H, (uint64_t)G * H, ndir, H, (uint64_t)G * H * H, SK, true, &mapW)) return -1;
  TcBwd p{T, B, H, ndir, 0, 0, gates, stash, out, lengths, h0, c0, dout, dh_final, dc_final, dh0, dc0, carry, nullptr};
  // large batch: 128 output units per CTA, as long as the wide grid still fills the GPU
  const bool wide = H % (4 * SU) == 0 && ceil_div(B, SB) * (H / (4 * SU)) * ndir >= (sm_count() > 0 ? sm_count() : 148);
  dim3 grid(H / (wide ? 4 * SU : SU), ceil_div(B, SB), ndir);
  // `carry` needs no initialisation: a sequence's entry is written at its injection step (t = len-1
  // or 0) before any step consumes it
  int n = T;
  auto run = [&](auto kernel, int stages, int arows) {
    const size_t sm = (size_t)stages * (arows * SK * 4 + (wide ? 4 : 1) * SU * SK * 4) + (2 * stages + 1) * 8 + 16 + 1024;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    for (int step = 0; step < T; ++step) {
      p.step = step;
      launch_pdl(kernel, grid, dim3(S_THREADS), sm, s, mapG, mapS, mapW, p);
    }
    if (dh0 || dc0) {
      p.final_only = 1;
      launch_pdl(kernel, grid, dim3(S_THREADS), sm, s, mapG, mapS, mapW, p);
      ++n;
    }
  };
  if (small) {
    if (G == 4) run(rnn_step_bwd_tc_kernel<4, 16, 64, 1, false>, 16, 64); else run(rnn_step_bwd_tc_kernel<3, 16, 64, 1, false>, 16, 64);
  } else if (wide) {
    if (G == 4) run(rnn_step_bwd_tc_kernel<4, 6, 128, 4, false>, 6, 128); else run(rnn_step_bwd_tc_kernel<3, 6, 128, 4, false>, 6, 128);
  } else {
    if (G == 4) run(rnn_step_bwd_tc_kernel<4, 8, 128, 1, false>, 8, 128); else run(rnn_step_bwd_tc_kernel<3, 8, 128, 1, false>, 8, 128);
  }
  note_launches(n);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("rnn_layer_bwd(tc step): launch failed: %s", cudaGetErrorString(e));
  return 0;
}

// ---- bf16-operand launchers (large batch: the recurrence is L2-bandwidth bound on its operand tiles)
// w_hh_bf: [ndir][G*H][H] bf16 copy of W_hh; out_bf: [T][B][ndir*H] bf16, written here (h_t), read as the next step's A
int rnn_layer_fwd_bfstep(int mode, int T, int B, int H, int ndir, float* gates, const uint16_t* w_hh_bf, const float* b_hh,
                         const int64_t* lengths, float* out, uint16_t* out_bf, float* stash, float* h_final, cudaStream_t s) {
  const int G = mode == SLNLP_MODE_LSTM ? 4 : 3;
  if (H % 64 != 0 || H < 64 || H > 4096 || B < 1 || encode_fn() == nullptr) return -1;
  if ((((uintptr_t)gates | (uintptr_t)out | (uintptr_t)w_hh_bf | (uintptr_t)out_bf | (uintptr_t)stash | (uintptr_t)b_hh) & 15) ||
      (h_final && ((uintptr_t)h_final & 15)))
    return -1;
  CUtensorMap mapH, mapW;
  if (!tensor_map3_bf16(out_bf, (uint64_t)ndir * H, B, T, (uint64_t)ndir * H, (uint64_t)B * ndir * H, SB, &mapH)) return -1;
  if (!tensor_map3_bf16(w_hh_bf, H, (uint64_t)G * H, ndir, H, (uint64_t)G * H * H, SU, &mapW)) return -1;
  TcFwd p{T, B, H, ndir, 0, gates, b_hh, lengths, nullptr, nullptr, out, stash, h_final, reinterpret_cast<__nv_bfloat16*>(out_bf)};
  dim3 grid(H / SU, ceil_div(B, SB), ndir);
  auto run = [&](auto kernel, int stages) {
    const size_t sm = (size_t)stages * (SB * SK * 4 + G * SU * SK * 4) + (2 * stages + 1) * 8 + 16 + 1024;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    for (int step = 0; step < T; ++step) {
      p.step = step;
      launch_pdl(kernel, grid, dim3(S_THREADS), sm, s, mapH, mapH, mapW, p);
    }
  };
  if (G == 4) run(rnn_step_fwd_tc_kernel<4, 6, 128, true>, 6); else run(rnn_step_fwd_tc_kernel<3, 6, 128, true>, 6);
  note_launches(T);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("rnn_layer_fwd(bf16 step): launch failed: %s", cudaGetErrorString(e));
  return 0;
}

// LSTM only.  w_hhT_bf: [ndir][H][G*H] bf16 copy of W_hh^T; dg_bf: [T][B][ndir*G*H] bf16, written here
int rnn_layer_bwd_bfstep(int mode, int T, int B, int H, int ndir, float* gates, uint16_t* dg_bf, float* stash, const float* out,
                         const uint16_t* w_hhT_bf, const int64_t* lengths, const float* dout, const float* dh_final,
                         const float* dc_final, float* carry, cudaStream_t s) {
  if (mode != SLNLP_MODE_LSTM) return -1;
  constexpr int G = 4;
  if (H % (4 * SU) != 0 || H > 4096 || B < 1 || encode_fn() == nullptr) return -1;
  const void* ptrs[] = {gates, dg_bf, stash, out, w_hhT_bf, dout, dh_final, dc_final, carry};
  for (const void* q : ptrs)
    if (q && ((uintptr_t)q & 15)) return -1;
  CUtensorMap mapG, mapW;
  if (!tensor_map3_bf16(dg_bf, (uint64_t)ndir * G * H, B, T, (uint64_t)ndir * G * H, (uint64_t)B * ndir * G * H, SB, &mapG)) return -1;
  if (!tensor_map3_bf16(w_hhT_bf, (uint64_t)G * H, H, ndir, (uint64_t)G * H, (uint64_t)G * H * H, SU, &mapW)) return -1;
  TcBwd p{T, B, H, ndir, 0, 0, gates, stash, out, lengths, nullptr, nullptr, dout, dh_final, dc_final, nullptr, nullptr, carry,
          reinterpret_cast<__nv_bfloat16*>(dg_bf)};
  dim3 grid(H / (4 * SU), ceil_div(B, SB), ndir);
  constexpr int stages = 6;
  const size_t sm = (size_t)stages * (SB * SK * 4 + 4 * SU * SK * 4) + (2 * stages + 1) * 8 + 16 + 1024;
  auto kernel = rnn_step_bwd_tc_kernel<4, stages, 128, 4, true>;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  for (int step = 0; step < T; ++step) {
    p.step = step;
    launch_pdl(kernel, grid, dim3(S_THREADS), sm, s, mapG, mapG, mapW, p);
  }
  note_launches(T);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("rnn_layer_bwd(bf16 step): launch failed: %s", cudaGetErrorString(e));
  return 0;
}

// persistent CTA-pair form of the same step (rnn_step_pair.cu; LSTM); -1 = unsupported
int lstm_layer_fwd_pairstep(int T, int B, int H, int ndir, float* gates, const uint16_t* w_hh_bf, const float* b_hh,
                            const int64_t* lengths, float* out, uint16_t* out_bf, float* stash, float* h_final, uint16_t* gact_bf,
                            cudaStream_t s);
int lstm_layer_bwd_pairstep(int T, int B, int H, int ndir, float* gates, uint16_t* dg_bf, const float* stash,
                            const uint16_t* w_hhT_bf, const int64_t* lengths, const float* dout, const float* dh_final,
                            const float* dc_final, float* carry, int write_f32, const uint32_t* dout_keep, float dout_scale,
                            int gates_in_dg, cudaStream_t s);
static bool pair_step_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SLNLP_PAIR_STEP");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

}  // namespace slnlp

extern "C" int slnlp_rnn_layer_fwd_bf16_ex(int mode, int T, int B, int H, int ndir, float* gates, const uint16_t* w_hh_bf,
                                           const float* b_hh, const int64_t* lengths, float* out, uint16_t* out_bf, float* stash,
                                           float* h_final, uint16_t* gates_act_bf, slnlp_stream_t stream);
extern "C" int slnlp_rnn_layer_fwd_bf16(int mode, int T, int B, int H, int ndir, float* gates, const uint16_t* w_hh_bf,
                                        const float* b_hh, const int64_t* lengths, float* out, uint16_t* out_bf, float* stash,
                                        float* h_final, slnlp_stream_t stream) {
  return slnlp_rnn_layer_fwd_bf16_ex(mode, T, B, H, ndir, gates, w_hh_bf, b_hh, lengths, out, out_bf, stash, h_final, nullptr, stream);
}
extern "C" int slnlp_rnn_layer_fwd_bf16_ex(int mode, int T, int B, int H, int ndir, float* gates, const uint16_t* w_hh_bf,
                                           const float* b_hh, const int64_t* lengths, float* out, uint16_t* out_bf, float* stash,
                                           float* h_final, uint16_t* gates_act_bf, slnlp_stream_t stream) {
  using namespace slnlp;
  SLNLP_CHECK_ARG(mode == SLNLP_MODE_LSTM || mode == SLNLP_MODE_GRU, "rnn_layer_fwd_bf16: bad mode %d", mode);
  SLNLP_CHECK_ARG(T > 0 && B > 0 && H > 0 && (ndir == 1 || ndir == 2), "rnn_layer_fwd_bf16: bad shape");
  SLNLP_CHECK_ARG(gates && w_hh_bf && b_hh && out_bf && stash, "rnn_layer_fwd_bf16: null pointer");
  if (mode == SLNLP_MODE_LSTM && pair_step_enabled()) {
    const int rp = lstm_layer_fwd_pairstep(T, B, H, ndir, gates, w_hh_bf, b_hh, lengths, out, out_bf, stash, h_final, gates_act_bf,
                                           as_stream(stream));
    if (rp >= 0) return rp;
  }
  SLNLP_CHECK_ARG(!gates_act_bf, "rnn_layer_fwd_bf16: a bf16 gate stash needs the CTA-pair kernels: ask slnlp_rnn_bf16_pair_supported");
  SLNLP_CHECK_ARG(out, "rnn_layer_fwd_bf16: out = NULL (bf16 copy only) needs the CTA-pair kernels: ask slnlp_rnn_bf16_pair_supported");
  const int rc = rnn_layer_fwd_bfstep(mode, T, B, H, ndir, gates, w_hh_bf, b_hh, lengths, out, out_bf, stash, h_final, as_stream(stream));
  SLNLP_CHECK_ARG(rc >= 0, "rnn_layer_fwd_bf16: needs H a multiple of 64 and 16-byte aligned operands");
  return rc;
}

extern "C" int slnlp_rnn_layer_bwd_bf16(int mode, int T, int B, int H, int ndir, float* gates, uint16_t* dg_bf, float* stash,
                                        const float* out, const uint16_t* w_hhT_bf, const int64_t* lengths, const float* dout,
                                        const float* dh_final, const float* dc_final, float* carry, int write_f32,
                                        const uint32_t* dout_keep, float dout_scale, int gates_in_dg, slnlp_stream_t stream) {
  using namespace slnlp;
  SLNLP_CHECK_ARG(mode == SLNLP_MODE_LSTM, "rnn_layer_bwd_bf16: LSTM only (mode %d)", mode);
  SLNLP_CHECK_ARG(T > 0 && B > 0 && H > 0 && (ndir == 1 || ndir == 2), "rnn_layer_bwd_bf16: bad shape");
  SLNLP_CHECK_ARG(gates && dg_bf && stash && out && w_hhT_bf && carry, "rnn_layer_bwd_bf16: null pointer");
  if (pair_step_enabled()) {
    const int rp = lstm_layer_bwd_pairstep(T, B, H, ndir, gates, dg_bf, stash, w_hhT_bf, lengths, dout, dh_final, dc_final, carry, write_f32,
                                           dout_keep, dout_scale, gates_in_dg, as_stream(stream));
    if (rp >= 0) return rp;
  }
  SLNLP_CHECK_ARG(!dout_keep && !gates_in_dg, "rnn_layer_bwd_bf16: a dropout keep mask needs the CTA-pair kernels: ask slnlp_rnn_bf16_pair_supported");
  const int rc = rnn_layer_bwd_bfstep(mode, T, B, H, ndir, gates, dg_bf, stash, out, w_hhT_bf, lengths, dout, dh_final, dc_final,
                                      carry, as_stream(stream));
  SLNLP_CHECK_ARG(rc >= 0, "rnn_layer_bwd_bf16: needs H a multiple of 128 and 16-byte aligned operands");
  return rc;
}

// 1 where slnlp_rnn_layer_fwd/bwd_bf16 run the persistent CTA-pair kernels (the forms with out = NULL / a keep mask)
extern "C" int slnlp_rnn_bf16_pair_supported(int mode, int T, int B, int H, int ndir) {
  return (mode == SLNLP_MODE_LSTM && slnlp::pair_step_enabled() && T > 1 && B > 256 && ndir == 2 && (H == 256 || H == 512 || H == 1024) &&
          slnlp::encode_fn() != nullptr) ? 1 : 0;
}

extern "C" int slnlp_rnn_bf16_step_supported(int mode, int T, int B, int H, int ndir) {
  (void)ndir;
  // the per-step kernels are the family of batches beyond the persistent / cluster kernels' reach
  return (mode == SLNLP_MODE_LSTM && T > 1 && B > 256 && H % 128 == 0 && H >= 256 && H <= 4096 && slnlp::encode_fn() != nullptr) ? 1 : 0;
}
