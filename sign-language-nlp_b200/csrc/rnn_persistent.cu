// K4 on the 5th-gen tensor cores: persistent recurrent layer for H = 128.
//
// One CTA owns one direction and a slice of 4 (16 at large batch) sequences for ALL timesteps
// (the recurrence is independent across the batch, so no CTA ever waits for another).  Warp
// specialised: 16 epilogue warps (512 threads, one per (hidden unit, sequence column group)) and ONE
// MMA warp that does nothing but issue the step's tcgen05.mma batch as soon as the CTA barrier that
// publishes h_t releases.  (Round-2 per-phase %clock table of the previous single-role layout,
// profiles/r02_persist_phases.txt: the issuing thread spent 419 cycles of a 1,253-cycle step issuing
// the 32 MMAs and then another 317 issuing its own share of the deferred stores and prefetches while
// the other 15 warps already waited for it; the step loop was instruction-issue bound - ~155
// instructions per warp and step, a third of them 64-bit address arithmetic - not tensor bound: the
// accumulators were ready 80 cycles after the issuer finally looked.):
//   * W_hh (bf16) is loaded ONCE into TENSOR MEMORY (128 lanes x 64 packed columns per gate
//     tile) and stays resident for the whole layer as the A operand of every MMA; reading it
//     from shared memory instead costs 128 KB per step through a 128 B/clk port;
//   * each step is D[gate rows (M=128 per gate tile), batch (N=16)] = W_hh_g . h^T issued by
//     one elected thread as tcgen05.mma kind::f16 (A from TMEM, B = the h tile in shared
//     memory, bf16 x bf16 -> fp32 in TMEM); swap-AB: gate rows are MMA-M, the batch is MMA-N;
//   * the epilogue threads (TMEM lane = hidden unit) read the 4 (3) gate accumulators with
//     tcgen05.ld, add the hoisted x W_ih^T + b_ih, apply the gate nonlinearities and the cell
//     update in fp32, keep c / h in registers across timesteps, and store h (bf16) straight
//     into the next step's B tile - the only store on the step-to-step critical path;
//   * out / stash / activated gates go to HBM one iteration later and the next step's hoisted
//     projection is prefetched into a second register set, both while the next MMAs run.
// The BPTT twin runs dG . W_hh (K = G*H) the same way with W_hh^T resident in TMEM.
//
// Numerics: the recurrent product rounds h and W_hh (dG in BPTT) to bf16 with fp32
// accumulation - the 2e-2 path of north_star.  Everything else is fp32.
#include <stdlib.h>
#include "tc05.cuh"

namespace slnlp {

constexpr int PH = 128;   // hidden size handled by this kernel
constexpr int PN = 16;    // MMA N (the smallest N of an M = 128 instruction)
// PSEQ (template): sequences a CTA actually owns; columns PSEQ..PN-1 of the B tile stay zero.  The per-step
// gate math and stores scale with PSEQ, and at the reference's batch of 50 most SMs are idle, so small
// batches run 4 sequences per CTA (26 CTAs at B = 50) and large ones 16 (fewer W_hh loads, fewer waves).
// epilogue threads = 128 x NCG (TMEM lane quadrant = warp % 4, column group = warp / 4), each owning PC batch
// columns of one hidden unit: a CTA owns PSEQ = NCG x PC sequences.  The tensor-core part of a step costs the
// same whatever PSEQ is (the MMA N is 16 at least), the epilogue scales with it, and at the reference's
// batch of 50 most SMs are idle: the launcher takes the smallest PSEQ whose grid still fits the device in one
// wave (B = 50: one sequence per CTA, 100 CTAs).  One more warp issues the MMAs.
constexpr int NACC = 4;                 // BPTT: partial accumulators (independent MMA chains)
constexpr int A_COL0 = 128;    // TMEM: accumulators in columns [0, 128), the resident W_hh operand from 128 on
constexpr int TMEM_COLS = 512; // 64 + up to 256 operand columns -> the whole tensor memory of the SM

// ---- per-phase cycle table of the step loop (slnlp_debug_persist_config): the PROF instantiations read
// %clock at fixed points of every step and three threads of CTA (0,0) (the MMA warp's lane 0, thread 160
// and thread 511 of the epilogue warps) publish the per-phase sums here.  [fwd|bwd][thread][phase]
constexpr int NPHASE = 8;
__device__ uint32_t g_persist_prof[2][3][NPHASE];
__device__ __forceinline__ uint32_t clk32() {
  uint32_t c;
  asm volatile("mov.u32 %0, %%clock;" : "=r"(c));
  return c;
}
struct PhaseClock {
  uint32_t acc[NPHASE], last;
  __device__ __forceinline__ void start() {
#pragma unroll
    for (int i = 0; i < NPHASE; ++i) acc[i] = 0;
    last = clk32();
  }
  __device__ __forceinline__ void mark(int i) {
    const uint32_t now = clk32();
    acc[i] += now - last;
    last = now;
  }
  __device__ __forceinline__ void publish(int which, int ethr) {
    const int tid = threadIdx.x;
    const int slot = tid == ethr ? 0 : tid == (ethr > 160 ? 160 : 32) ? 1 : tid == ethr - 1 ? 2 : -1;
    if (slot < 0 || blockIdx.x != 0 || blockIdx.y != 0) return;
#pragma unroll
    for (int i = 0; i < NPHASE; ++i) g_persist_prof[which][slot][i] = acc[i];
  }
};
#define PROF_MARK(i) do { if (PROF) pclk.mark(i); } while (0)

// CTA-wide barrier over the epilogue warps AND the MMA warp (reached from two code paths: bar.sync with an
// explicit thread count instead of __syncthreads)
template <int NTHREADS>
__device__ __forceinline__ void cta_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory"); }

struct PersistFwd {
  int T, B, ndir;
  float* gates;          // [T,B,ndir,G,H]
  const float* w_hh;     // [ndir,G*H,H]
  const float* b_hh;     // [ndir,G*H]
  const int64_t* lengths;
  const float* h0;
  const float* c0;
  float* out;            // [T,B,ndir*H]
  float* stash;          // [T,B,ndir,H]
  float* h_final;
  int64_t hf_d, hf_b;    // h_final strides: direction, batch ([ndir,B,H]: B*H, H; concatenated [B,ndir*H]: H, ndir*H)
  float* out_drop;       // dropout(out) for the next layer's input, or null
  float p_drop;
  const uint64_t* rng;   // {seed, step} of the module's Philox stream
  uint32_t site;
  const float* mask;     // precomputed keep / scale factors (out's layout) instead of Philox in the loop, or null
};

// the keep / scale factor slnlp_dropout(site) applies to element e of a tensor: Philox block e / 4, lane e % 4
__device__ __forceinline__ float dropout_factor(uint64_t seed, uint64_t step, uint32_t site, int64_t e, float p) {
  float u[4];
  philox_uniform4(seed, step, site, (uint64_t)(e >> 2), u);
  const int l = (int)(e & 3);
  const float uu = l == 0 ? u[0] : l == 1 ? u[1] : l == 2 ? u[2] : u[3];
  return uu < 1.f - p ? 1.f / (1.f - p) : 0.f;
}

// FA: partial accumulators per gate tile (FA independent MMA chains, summed by the epilogue).
// (Issuing from four warps instead of one was measured: the issue sequence halves, 433 -> 196 cycles, but the
// accumulators complete at the same time - the tensor pipe retires an M128 x N16 x K16 MMA every ~18 cycles,
// ~27 when its B operand differs from the previous MMA's, so the k-major order below (four gate tiles per
// h slice) is the fast one and a per-gate staged epilogue, which needs gate-major issue, loses.)
template <int G, int NCG, int PC, int FA, bool PROF>
__global__ void __launch_bounds__(128 * NCG + 32, 1) rnn_persistent_fwd_kernel(PersistFwd p) {
  constexpr int PSEQ = NCG * PC, ETHR = 128 * NCG, NTHR = ETHR + 32;
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int H = PH;
  uint8_t* sH = smem_raw;                       // [PN x 128] bf16, canonical (B operand)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sH + PN * H * 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 4);

  const int tid = threadIdx.x, warp = tid >> 5, warp_u = warp_uniform();
  const bool mma_warp = warp_u >= ETHR / 32;
  const int d = blockIdx.y, b0 = blockIdx.x * PSEQ;
  const int T = p.T, B = p.B;
  const int j = tid & (H - 1);   // hidden unit = TMEM lane
  const int cg = (tid >> 7) & 3; // column group: batch columns cg*PC .. cg*PC+PC-1

  // ---- prologue that only touches the weights: it runs BEFORE griddepcontrol.wait, i.e. under the tail of
  // the hoisted-projection GEMM that precedes this kernel in the stream (weights were last written by the
  // previous training step's SGD kernel, complete long before that GEMM started)
  const float* W = p.w_hh + (int64_t)d * G * H * H;
  if (tid == 0) mbar_init(bar, 1);
  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_mine = tmem + ((uint32_t)((warp & 3) * 32) << 16) + cg * PC;
  // W_hh -> tensor memory, the A operand of every step: gate tile g = [128 rows (lanes) x 128 k]
  // bf16 = 64 packed 32-bit columns at column A_COL0 + 64 g.  Reading A from TMEM instead of
  // shared memory takes the 128 KB-per-step operand fetch off the 128 B/clk shared-memory port
  // (it bounded the step at ~1000 cycles).  Warp w fills lane quadrant w % 4 of gate w / 4.
  if (!mma_warp) {
#pragma unroll 1
    for (int g = cg; g < G; g += NCG) {
      const float* wrow = W + ((int64_t)g * H + j) * H;
#pragma unroll 1
      for (int i = 0; i < 4; ++i) {
        uint32_t pk[16];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(wrow + i * 32) + u);
          pk[2 * u] = pack2_bf16(v.x, v.y);
          pk[2 * u + 1] = pack2_bf16(v.z, v.w);
        }
        tmem_st16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + A_COL0 + g * 64 + i * 16, pk);
      }
    }
    tmem_wait_st();
  }
  const float* bh = p.b_hh + (int64_t)d * G * H;
  float bias[G];
#pragma unroll
  for (int g = 0; g < G; ++g) bias[g] = mma_warp ? 0.f : bh[g * H + j];
  // the h tile rows PSEQ..PN-1 are never written: zero them once
  for (int e = tid; e < PN * H / 8; e += NTHR) {
    const int n = e % PN;   // canonical tile: 16-byte core rows, row index = (e % (PN)) within a K-group
    if (n >= PSEQ) reinterpret_cast<uint4*>(sH)[e] = make_uint4(0, 0, 0, 0);
  }

  pdl_wait();   // ---- from here on: data of this step (hoisted projection, lengths, initial state)

  float hreg[PC], creg[PC];
  int len[PC];
  if (!mma_warp) {
#pragma unroll
    for (int c = 0; c < PC; ++c) {
      const int n = cg * PC + c, b = b0 + n;
      hreg[c] = (p.h0 && b < B) ? p.h0[((int64_t)d * B + b) * H + j] : 0.f;
      creg[c] = (p.c0 && b < B) ? p.c0[((int64_t)d * B + b) * H + j] : 0.f;
      len[c] = b < B ? (p.lengths ? (int)p.lengths[b] : T) : 0;
      *reinterpret_cast<__nv_bfloat16*>(sH + canon_off(n, j, PN)) = __float2bfloat16(hreg[c]);
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  constexpr uint32_t idesc = make_idesc(128, PN);
  const uint64_t descB0 = make_desc(smem_u32(sH), PN * 16, 128);
  const bool have_state0 = p.h0 != nullptr;
  PhaseClock pclk;
  if (PROF) pclk.start();

  if (mma_warp) {
    // ================= MMA warp: one elected lane issues the step's batch, then the whole warp waits at
    // the CTA barrier for the epilogue warps to publish h_t
    for (int step = 0; step < T; ++step) {
      const bool do_mma = step > 0 || have_state0;
      if (do_mma && elect_one()) {
        // G gate tiles x (H/16) k-steps, A = W_hh from tensor memory, B = the h tile in shared memory;
        // k-major: consecutive MMAs share the B slice and accumulate into different gate tiles
#pragma unroll
        for (int kk = 0; kk < H / 16; ++kk)
#pragma unroll
          for (int g = 0; g < G; ++g)
            umma_bf16_ts(tmem + (g * FA + (kk % FA)) * PN, tmem + A_COL0 + g * 64 + kk * 8,
                         descB0 + (uint64_t)((kk * 2 * (PN * 16)) >> 4), idesc, kk >= FA ? 1u : 0u);
        umma_commit(bar);
      }
      __syncwarp();
      PROF_MARK(0);   // MMA issue + commit
      tc_fence_before();
      cta_bar<NTHR>();
      tc_fence_after();
      PROF_MARK(6);   // waiting for the epilogue warps (h_t published)
    }
  } else {
    // ================= epilogue warps
    // running pointers at the CURRENT step's row of every array; a step adds +-(row stride)
    const int64_t gstride = (int64_t)B * p.ndir * G * H, ostride = (int64_t)B * p.ndir * H;
    const int64_t gdelta = d == 0 ? gstride : -gstride, odelta = d == 0 ? ostride : -ostride;
    const int t0 = d == 0 ? 0 : T - 1;
    float* gcur[PC];
    float* ocur[PC];
    float* scur[PC];
    float* fin[PC];
    uint32_t hoff[PC];
    bool valid[PC];
    int tfin[PC];
#pragma unroll
    for (int c = 0; c < PC; ++c) {
      const int n = cg * PC + c;
      valid[c] = b0 + n < B;
      const int b = valid[c] ? b0 + n : 0;
      gcur[c] = p.gates + ((int64_t)b * p.ndir + d) * G * H + j + t0 * gstride;
      ocur[c] = p.out + (int64_t)b * p.ndir * H + (int64_t)d * H + j + t0 * ostride;
      scur[c] = p.stash + ((int64_t)b * p.ndir + d) * H + j + t0 * ostride;
      fin[c] = p.h_final ? p.h_final + (int64_t)d * p.hf_d + (int64_t)b * p.hf_b + j : nullptr;
      hoff[c] = canon_off(n, j, PN);
      tfin[c] = d == 0 ? len[c] - 1 : 0;
    }
    // inter-layer dropout fused into the deferred store (one more store and one Philox block per element, in
    // the slack under the MMAs): out_drop has out's layout
    const bool drop = p.out_drop != nullptr;
    const uint64_t rseed = (drop && !p.mask) ? p.rng[0] : 0, rstep = (drop && !p.mask) ? p.rng[1] : 0;
    const int64_t ddelta = p.out_drop - p.out;
    // with precomputed factors, the factor of the step being computed is loaded with the next step's prefetch
    // and used one iteration later by the deferred store (a load right before the store would put an L2 round
    // trip on the step-to-step chain)
    float mk[PC];
#pragma unroll
    for (int c = 0; c < PC; ++c) mk[c] = 0.f;
    auto keep = [&](const float* o, int c) {
      return p.mask ? mk[c] : dropout_factor(rseed, rstep, p.site, o - p.out, p.p_drop);
    };
    // hoisted input projection, loaded one step ahead into a second register set.  The two sets swap
    // roles every step (the loop is unrolled by two): a register copy at the end of the step would make
    // every warp wait for its loads there, on the step-to-step chain.
    float xa[G][PC], xb[G][PC];
#pragma unroll
    for (int c = 0; c < PC; ++c)
#pragma unroll
      for (int g = 0; g < G; ++g) xa[g][c] = t0 < len[c] ? gcur[c][g * H] : 0.f;
    // results of a step are written to HBM one iteration later, while the next step's MMAs run
    float gout[G][PC], hv[PC], sv[PC];
    uint32_t phase = 0;
    auto one_step = [&](float (&xg)[G][PC], float (&xn)[G][PC], int step) {
      const int t = d == 0 ? step : T - 1 - step;
      const int t_prev = d == 0 ? t - 1 : t + 1, t_next = d == 0 ? t + 1 : t - 1;
      const bool do_mma = step > 0 || have_state0;
      // off the critical path (the tensor core is busy): previous step -> HBM, next step's x <- HBM
      if (step > 0) {
#pragma unroll
        for (int c = 0; c < PC; ++c) {
          if (!valid[c]) continue;
          float* gp = gcur[c] - gdelta;
          float* op = ocur[c] - odelta;
          float* sp = scur[c] - odelta;
          if (t_prev >= len[c]) {
            *op = 0.f;
            *sp = 0.f;
            if (drop) op[ddelta] = 0.f;
            continue;
          }
#pragma unroll
          for (int g = 0; g < G; ++g) gp[g * H] = gout[g][c];
          *sp = sv[c];
          *op = hv[c];
          if (drop) op[ddelta] = hv[c] * keep(op, c);
          if (fin[c] && t_prev == tfin[c]) *fin[c] = hv[c];
        }
      }
      if (step + 1 < T) {
#pragma unroll
        for (int c = 0; c < PC; ++c) {
          const float* gp = gcur[c] + gdelta;
          const bool act = t_next < len[c];
#pragma unroll
          for (int g = 0; g < G; ++g) xn[g][c] = act ? gp[g * H] : 0.f;
        }
      }
      if (drop && p.mask) {
#pragma unroll
        for (int c = 0; c < PC; ++c) mk[c] = (valid[c] && t < len[c]) ? p.mask[ocur[c] - p.out] : 0.f;
      }
      PROF_MARK(1);   // issue of the deferred stores and of the next step's loads
      // accumulators of gate tile g (FA partial sums at columns (g*FA + q)*PN) -> a[]
      auto read_gate = [&](int g, float (&a)[PC]) {
        uint32_t raw[FA][PC];
#pragma unroll
        for (int q = 0; q < FA; ++q) tmem_ldn_nowait<PC>(tmem_mine + (g * FA + q) * PN, raw[q]);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < PC; ++c) {
          float v = __uint_as_float(raw[0][c]);
#pragma unroll
          for (int q = 1; q < FA; ++q) v += __uint_as_float(raw[q][c]);
          a[c] = v;
        }
      };
      // x + b_hh does not wait for the tensor core (sigmoid(v) = 0.5 tanh(0.5 v) + 0.5: the halving too)
      float xb[G][PC];
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int c = 0; c < PC; ++c) xb[g][c] = (G == 3 && g == 2) ? xg[g][c] : xg[g][c] + bias[g];
      float acc[G][PC];
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int c = 0; c < PC; ++c) acc[g][c] = 0.f;
      if (do_mma) {
        mbar_wait(bar, phase);
        tc_fence_after();
        PROF_MARK(2);   // wait for the MMAs
#pragma unroll
        for (int g = 0; g < G; ++g) read_gate(g, acc[g]);
        PROF_MARK(3);   // tcgen05.ld of the accumulators
      }
#pragma unroll
      for (int c = 0; c < PC; ++c) {
        if (G == 4) {
          const float gi = sigmoid_fast(xb[0][c] + acc[0][c]);
          const float gf = sigmoid_fast(xb[1][c] + acc[1][c]);
          const float gg = tanh_fast(xb[2][c] + acc[2][c]);
          const float go = sigmoid_fast(xb[G - 1][c] + acc[G - 1][c]);
          const float cn = gf * creg[c] + gi * gg;
          hv[c] = go * tanh_fast(cn);
          sv[c] = cn;
          gout[0][c] = gi; gout[1][c] = gf; gout[2][c] = gg; gout[G - 1][c] = go;
        } else {
          const float hn = acc[2][c] + bias[2];
          const float gr = sigmoid_fast(xb[0][c] + acc[0][c]);
          const float gz = sigmoid_fast(xb[1][c] + acc[1][c]);
          const float gn = tanh_fast(xb[2][c] + gr * hn);
          hv[c] = (1.f - gz) * gn + gz * hreg[c];
          sv[c] = hn;
          gout[0][c] = gr; gout[1][c] = gz; gout[2][c] = gn;
        }
      }
      if (do_mma) phase ^= 1;
#pragma unroll
      for (int c = 0; c < PC; ++c) {
        // state update + next step's B tile: the only stores the next MMA waits for
        if (valid[c] && t < len[c]) {
          if (G == 4) creg[c] = sv[c];
          hreg[c] = hv[c];
          *reinterpret_cast<__nv_bfloat16*>(sH + hoff[c]) = __float2bfloat16(hv[c]);
        }
      }
      PROF_MARK(4);   // gate math + h (bf16) into the next step's B tile
      // h tile (generic-proxy stores) -> visible to the tensor core; accumulators free to overwrite
      fence_async_smem();
      tc_fence_before();
      PROF_MARK(5);   // proxy fence (MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC)
      cta_bar<NTHR>();
      if (PROF) {     // BAR.SYNC defers its blocking to the next consumer of barrier-protected state: make one
        asm volatile("" ::"r"(*reinterpret_cast<volatile uint32_t*>(tmem_slot)) : "memory");
      }
      PROF_MARK(6);   // CTA barrier (time until the slowest warp arrives)
#pragma unroll
      for (int c = 0; c < PC; ++c) {
        gcur[c] += gdelta;
        ocur[c] += odelta;
        scur[c] += odelta;
      }
    };
    for (int step = 0; step < T; step += 2) {
      one_step(xa, xb, step);
      if (step + 1 < T) one_step(xb, xa, step + 1);
    }
    // the last step's results
    {
      const int t_last = d == 0 ? T - 1 : 0;
#pragma unroll
      for (int c = 0; c < PC; ++c) {
        if (!valid[c]) continue;
        float* gp = gcur[c] - gdelta;
        float* op = ocur[c] - odelta;
        float* sp = scur[c] - odelta;
        if (t_last >= len[c]) {
          *op = 0.f;
          *sp = 0.f;
          if (drop) op[ddelta] = 0.f;
          continue;
        }
#pragma unroll
        for (int g = 0; g < G; ++g) gp[g * H] = gout[g][c];
        *sp = sv[c];
        *op = hv[c];
        if (drop) op[ddelta] = hv[c] * keep(op, c);
        if (fin[c] && t_last == tfin[c]) *fin[c] = hv[c];
      }
    }
  }
  if (PROF) pclk.publish(0, ETHR);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

// ---------------------------------------------------------------- BPTT twin
struct PersistBwd {
  int T, B, ndir;
  float* gates;          // in: activated gates; out: d pre-activations (x side)
  float* stash;          // LSTM: c_t; GRU: hn -> d hn
  const float* out;
  const float* w_hh;
  const int64_t* lengths;
  const float* h0;
  const float* c0;
  const float* dout;
  const float* dh_final;
  const float* dc_final;
  float* dh0;
  float* dc0;
  int64_t hf_d, hf_b;    // dh_final / dc_final strides (direction, batch)
  float p_drop;          // > 0: dout is the gradient of dropout(out): apply the forward's mask while reading it
  const uint64_t* rng;
  uint32_t site;
  const float* mask;     // precomputed factors (dout's layout) instead of Philox, or null
};

// NACC partial accumulators of the single [128 x PN] output tile (independent MMA chains).
template <int G, int NCG, int PC, bool PROF>
__global__ void __launch_bounds__(128 * NCG + 32, 1) rnn_persistent_bwd_kernel(PersistBwd p) {
  constexpr int NA = NACC;
  constexpr int PSEQ = NCG * PC, ETHR = 128 * NCG, NTHR = ETHR + 32;
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int H = PH, GH = G * PH;
  uint8_t* sD = smem_raw;                        // B' = dG: [PN x GH] bf16, canonical
  uint64_t* bar = reinterpret_cast<uint64_t*>(sD + PN * GH * 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, warp_u = warp_uniform();
  const bool mma_warp = warp_u >= ETHR / 32;
  const int d = blockIdx.y, b0 = blockIdx.x * PSEQ;
  const int T = p.T, B = p.B;
  const int k = tid & (H - 1);  // hidden unit = TMEM lane = output row of W_hh^T
  const int cg = (tid >> 7) & 3;

  // ---- weights-only prologue, before griddepcontrol.wait (see the forward kernel)
  const float* W = p.w_hh + (int64_t)d * GH * H;
  for (int e = tid; e < PN * GH / 8; e += NTHR) reinterpret_cast<uint4*>(sD)[e] = make_uint4(0, 0, 0, 0);
  if (tid == 0) mbar_init(bar, 1);
  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_mine = tmem + ((uint32_t)((warp & 3) * 32) << 16) + cg * PC;
  // A' = W_hh^T -> tensor memory: lane = hidden unit k, K index = gate row j' (GH of them) = GH/2 packed
  // columns from A_COL0.  Column group cg fills rows j' in [cg*GH/NCG, (cg+1)*GH/NCG): global reads are
  // coalesced across the warp (consecutive k).
  if (!mma_warp) {
    constexpr int JQ = GH / NCG;   // gate rows per column group -> JQ / 2 packed columns = JQ / 32 stores of 16
#pragma unroll 1
    for (int i = 0; i < JQ / 32; ++i) {
      uint32_t pk[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int jr = cg * JQ + i * 32 + 2 * u;
        pk[u] = pack2_bf16(__ldg(W + (int64_t)jr * H + k), __ldg(W + (int64_t)(jr + 1) * H + k));
      }
      tmem_st16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + A_COL0 + cg * (JQ / 2) + i * 16, pk);
    }
    tmem_wait_st();
  }

  pdl_wait();   // ---- from here on: activations and gradients of this step

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  constexpr uint32_t idesc = make_idesc(128, PN);
  const uint64_t descB0 = make_desc(smem_u32(sD), PN * 16, 128);
  const int nsteps = T + ((p.dh0 || p.dc0) ? 1 : 0);
  PhaseClock pclk;
  if (PROF) pclk.start();

  if (mma_warp) {
    for (int step = 0; step < nsteps; ++step) {
      if (step > 0 && elect_one()) {
        // A' = W_hh^T from tensor memory, B' = the dG tile; NA partial accumulators: consecutive
        // MMAs are independent, the epilogue adds them
#pragma unroll
        for (int kk = 0; kk < GH / 16; ++kk)
          umma_bf16_ts(tmem + (kk % NA) * PN, tmem + A_COL0 + kk * 8,
                       descB0 + (uint64_t)((kk * 2 * (PN * 16)) >> 4), idesc, kk >= NA ? 1u : 0u);
        umma_commit(bar);
      }
      __syncwarp();
      PROF_MARK(0);
      if (step == T) break;     // the final pseudo-step (state gradients) has no barrier
      tc_fence_before();
      cta_bar<NTHR>();
      tc_fence_after();
      PROF_MARK(6);
    }
  } else {
    int len[PC];
    float carry[PC];  // LSTM: dc carry; GRU: direct dh carry (dh * z)
    // running pointers at the CURRENT step's row; BPTT walks the time axis against the forward direction
    const int64_t gstride = (int64_t)B * p.ndir * GH, ostride = (int64_t)B * p.ndir * H;
    const int64_t gdelta = d == 0 ? -gstride : gstride, odelta = d == 0 ? -ostride : ostride;
    const int t0 = d == 0 ? T - 1 : 0;
    float* gcur[PC];
    float* scur[PC];
    const float* ocur[PC];
    const float* dcur[PC];
    int64_t cidx[PC], fidx[PC];
    uint32_t doff[PC];
    bool valid[PC];
#pragma unroll
    for (int c = 0; c < PC; ++c) {
      const int n = cg * PC + c;
      valid[c] = b0 + n < B;
      const int b = valid[c] ? b0 + n : 0;
      len[c] = valid[c] ? (p.lengths ? (int)p.lengths[b] : T) : 0;
      carry[c] = 0.f;
      gcur[c] = p.gates + ((int64_t)b * p.ndir + d) * GH + k + t0 * gstride;
      scur[c] = p.stash + ((int64_t)b * p.ndir + d) * H + k + t0 * ostride;
      ocur[c] = p.out + (int64_t)b * p.ndir * H + (int64_t)d * H + k + t0 * ostride;
      dcur[c] = p.dout ? p.dout + (int64_t)b * p.ndir * H + (int64_t)d * H + k + t0 * ostride : nullptr;
      cidx[c] = ((int64_t)d * B + b) * H + k;
      fidx[c] = (int64_t)d * p.hf_d + (int64_t)b * p.hf_b + k;
      doff[c] = canon_off(n, k, PN);
    }
    const bool undrop = p.p_drop > 0.f && p.dout != nullptr;
    const uint64_t rseed = (undrop && !p.mask) ? p.rng[0] : 0, rstep = (undrop && !p.mask) ? p.rng[1] : 0;
    // per-step operands, loaded one step ahead into a second register set (the two sets swap roles every
    // step: loop unrolled by two): activated gates, stash, predecessor state, dout
    struct StepIn {
      float g[G][PC], s[PC], pv[PC], d[PC];
    };
    StepIn sa, sb;
    // `rel` = rows ahead of the current step (0: this step, 1: the next one BPTT visits)
    auto load_step = [&](StepIn& in, int t, int rel) {
      // the state the forward step t started from sits one row AGAINST the BPTT walk: row t - 1 (d = 0)
      const int tp = d == 0 ? t - 1 : t + 1;
      const bool has_prev = tp >= 0 && tp < T;
#pragma unroll
      for (int c = 0; c < PC; ++c) {
        const bool act = t < len[c];
        const float* gp = gcur[c] + rel * gdelta;
        const float* sp = scur[c] + rel * odelta;
#pragma unroll
        for (int g = 0; g < G; ++g) in.g[g][c] = act ? gp[g * H] : 0.f;
        in.s[c] = act ? *sp : 0.f;
        float dv = 0.f;
        if (act && dcur[c]) {
          const float* dq = dcur[c] + rel * odelta;
          dv = *dq;
          if (undrop) dv *= p.mask ? p.mask[dq - p.dout] : dropout_factor(rseed, rstep, p.site, dq - p.dout, p.p_drop);
        }
        in.d[c] = dv;
        float pv = 0.f;
        if (act) {
          if (G == 4) pv = has_prev ? *(sp + odelta) : (p.c0 ? p.c0[cidx[c]] : 0.f);
          else pv = has_prev ? *(ocur[c] + (rel + 1) * odelta) : (p.h0 ? p.h0[cidx[c]] : 0.f);
        }
        in.pv[c] = pv;
      }
    };
    load_step(sa, t0, 0);
    // d(pre-activations) of a step go to HBM one iteration later, while the next step's MMAs run
    float dg[G][PC], dst[PC];
    auto store_prev = [&]() {
#pragma unroll
      for (int c = 0; c < PC; ++c) {
        if (!valid[c]) continue;
        float* gp = gcur[c] - gdelta;
#pragma unroll
        for (int g = 0; g < G; ++g) gp[g * H] = dg[g][c];
        if (G == 3) *(scur[c] - odelta) = dst[c];
      }
    };
    uint32_t phase = 0;
    bool pending = false;
    // returns false after the final (state-gradient only) pseudo-step
    auto one_step = [&](StepIn& cur, StepIn& nxt, int step) -> bool {
      const bool final_only = step == T;
      const int t = final_only ? (d == 0 ? -1 : T) : (d == 0 ? T - 1 - step : step);
      const bool do_mma = step > 0;
      // off the critical path: previous step's gradients -> HBM, next step's operands <- HBM
      if (pending) {
        store_prev();
        pending = false;
      }
      if (step + 1 < T) load_step(nxt, d == 0 ? t - 1 : t + 1, 1);
      PROF_MARK(1);
      // everything that does not depend on this step's W_hh^T dG product is computed while the MMAs run:
      // dG is LINEAR in the incoming dh, so only a handful of multiplies remain after the wait
      // (one column per thread - the small-batch shape; with four columns per thread the extra live
      // registers would spill, so there the coefficients are formed after the wait as before)
      constexpr bool PRE = PC == 1;
      float k0[PC], k1[PC], k2[PC], k3[PC], a1[PC], dhb[PC], dcb[PC];
      bool act[PC], inj[PC];
      auto coeff = [&](int c) {
        act[c] = t < len[c];
        inj[c] = d == 0 ? t == len[c] - 1 : t == 0;
        dhb[c] = cur.d[c] + ((inj[c] && p.dh_final && act[c]) ? p.dh_final[fidx[c]] : 0.f);
        if (G == 4) {
          const float gi = cur.g[0][c], gf = cur.g[1][c], gg = cur.g[2][c], go = cur.g[G - 1][c];
          const float tc = tanh_fast(cur.s[c]);
          a1[c] = go * (1.f - tc * tc);
          k0[c] = gg * gi * (1.f - gi);
          k1[c] = cur.pv[c] * gf * (1.f - gf);
          k2[c] = gi * (1.f - gg * gg);
          k3[c] = tc * go * (1.f - go);
          dcb[c] = inj[c] ? ((p.dc_final && act[c]) ? p.dc_final[fidx[c]] : 0.f) : carry[c];
        } else {
          const float gr = cur.g[0][c], gz = cur.g[1][c], gn = cur.g[2][c];
          a1[c] = (1.f - gz) * (1.f - gn * gn);
          k0[c] = cur.s[c] * gr * (1.f - gr);
          k1[c] = (cur.pv[c] - gn) * gz * (1.f - gz);
          k2[c] = gr;
          k3[c] = gz;
          dcb[c] = inj[c] ? 0.f : carry[c];
        }
      };
      if (PRE && !final_only) {
#pragma unroll
        for (int c = 0; c < PC; ++c) coeff(c);
      }
      float m[PC];
      if (do_mma) {
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        PROF_MARK(2);
        uint32_t raw[NA][PC];
#pragma unroll
        for (int a = 0; a < NA; ++a) tmem_ldn_nowait<PC>(tmem_mine + a * PN, raw[a]);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < PC; ++c) {
          float acc = __uint_as_float(raw[0][c]);
#pragma unroll
          for (int a = 1; a < NA; ++a) acc += __uint_as_float(raw[a][c]);
          m[c] = acc;
        }
        PROF_MARK(3);
      } else {
#pragma unroll
        for (int c = 0; c < PC; ++c) m[c] = 0.f;
      }
      if (final_only) {
#pragma unroll
        for (int c = 0; c < PC; ++c) {
          if (!valid[c]) continue;
          if (G == 4) {
            if (p.dh0) p.dh0[cidx[c]] = m[c];
            if (p.dc0) p.dc0[cidx[c]] = carry[c];
          } else if (p.dh0) {
            p.dh0[cidx[c]] = m[c] + carry[c];
          }
        }
        return false;
      }
#pragma unroll
      for (int c = 0; c < PC; ++c) {
        if (!PRE) coeff(c);
        dst[c] = 0.f;
        if (!act[c]) {
#pragma unroll
          for (int g = 0; g < G; ++g) dg[g][c] = 0.f;
        } else if (G == 4) {
          const float dh = dhb[c] + (inj[c] ? 0.f : m[c]);
          const float dc = dh * a1[c] + dcb[c];
          dg[0][c] = dc * k0[c];
          dg[1][c] = dc * k1[c];
          dg[2][c] = dc * k2[c];
          dg[G - 1][c] = dh * k3[c];
          carry[c] = dc * cur.g[1][c];
        } else {
          const float dh = dhb[c] + (inj[c] ? 0.f : m[c] + dcb[c]);
          const float da_n = dh * a1[c];
          dg[0][c] = da_n * k0[c];
          dg[1][c] = dh * k1[c];
          dg[2][c] = da_n;
          dst[c] = da_n * k2[c];
          carry[c] = dh * k3[c];
        }
        // h-side gradients of this step = next step's B operand: dG[n][g*H + k]; consecutive gates are
        // H/8 K-groups apart in the canonical tile.  The only stores the next MMA waits for.
        if (valid[c]) {
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const float hv = (G == 3 && g == 2) ? dst[c] : dg[g][c];
            *reinterpret_cast<__nv_bfloat16*>(sD + doff[c] + g * (H / 8) * (PN * 16)) = __float2bfloat16(hv);
          }
        }
      }
      PROF_MARK(4);
      fence_async_smem();
      tc_fence_before();
      PROF_MARK(5);
      cta_bar<NTHR>();
      if (PROF) {
        asm volatile("" ::"r"(*reinterpret_cast<volatile uint32_t*>(tmem_slot)) : "memory");
      }
      PROF_MARK(6);
      pending = true;
#pragma unroll
      for (int c = 0; c < PC; ++c) {
        gcur[c] += gdelta;
        scur[c] += odelta;
        ocur[c] += odelta;
        if (dcur[c]) dcur[c] += odelta;
      }
      return true;
    };
    for (int step = 0; step < nsteps; step += 2) {
      if (!one_step(sa, sb, step)) break;
      if (step + 1 < nsteps && !one_step(sb, sa, step + 1)) break;
    }
    if (pending) store_prev();
  }
  if (PROF) pclk.publish(1, ETHR);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

static size_t persist_fwd_smem(int) { return (size_t)PN * PH * 2 + 64; }
static size_t persist_bwd_smem(int G) { return (size_t)PN * G * PH * 2 + 64; }

static bool tc_shape_ok(int H, const float* w_hh) { return H == PH && ((uintptr_t)w_hh & 15) == 0; }

// sequences per CTA: the smallest of 1 / 2 / 4 whose grid fits the device in one wave, else 16
// ($SLNLP_PERSIST_PSEQ overrides)
static int seqs_per_cta(int B, int ndir) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("SLNLP_PERSIST_PSEQ");
    forced = e ? atoi(e) : 0;
  }
  if (forced == 1 || forced == 2 || forced == 4 || forced == 16) return forced;
  const int sms = sm_count() > 0 ? sm_count() : 148;
  for (int ps : {1, 2, 4})
    if (ceil_div(B, ps) * ndir <= sms) return ps;
  return 16;
}

// debug / tuning switches (slnlp_debug_persist_config): loop variant of the forward kernel and the
// instrumented (per-phase %clock) instantiations.  Plain ints: set before launching, not per stream.
static int g_persist_var = -1, g_persist_profile_on = 0;
static int persist_var() {
  if (g_persist_var < 0) {
    const char* e = getenv("SLNLP_PERSIST_VAR");
    g_persist_var = e ? atoi(e) : 1;
  }
  return g_persist_var;
}

template <int G, int NCG, int PC, int FA>
static void launch_persist_fwd2(const PersistFwd& p, cudaStream_t s) {
  const size_t sm = persist_fwd_smem(G);
  const dim3 grid(ceil_div(p.B, NCG * PC), p.ndir), block(128 * NCG + 32);
  if (g_persist_profile_on) launch_pdl(rnn_persistent_fwd_kernel<G, NCG, PC, FA, true>, grid, block, sm, s, p);
  else launch_pdl(rnn_persistent_fwd_kernel<G, NCG, PC, FA, false>, grid, block, sm, s, p);
}
template <int G>
static void launch_persist_fwd(const PersistFwd& p, cudaStream_t s) {
  const bool fa1 = persist_var() & 1;       // one accumulator per gate tile instead of two partial ones
  const int ps = seqs_per_cta(p.B, p.ndir);
#define SLNLP_FWD(NCG, PC) do { if (fa1) launch_persist_fwd2<G, NCG, PC, 1>(p, s); else launch_persist_fwd2<G, NCG, PC, 2>(p, s); } while (0)
  if (ps == 1) SLNLP_FWD(1, 1);
  else if (ps == 2) SLNLP_FWD(2, 1);
  else if (ps == 4) SLNLP_FWD(4, 1);
  else SLNLP_FWD(4, 4);
#undef SLNLP_FWD
}
template <int G, int NCG, int PC>
static void launch_persist_bwd2(const PersistBwd& p, cudaStream_t s) {
  const size_t sm = persist_bwd_smem(G);
  const dim3 grid(ceil_div(p.B, NCG * PC), p.ndir), block(128 * NCG + 32);
  if (g_persist_profile_on) launch_pdl(rnn_persistent_bwd_kernel<G, NCG, PC, true>, grid, block, sm, s, p);
  else launch_pdl(rnn_persistent_bwd_kernel<G, NCG, PC, false>, grid, block, sm, s, p);
}
template <int G>
static void launch_persist_bwd(const PersistBwd& p, cudaStream_t s) {
  const int ps = seqs_per_cta(p.B, p.ndir);
  if (ps == 1) launch_persist_bwd2<G, 1, 1>(p, s);
  else if (ps == 2) launch_persist_bwd2<G, 2, 1>(p, s);
  else if (ps == 4) launch_persist_bwd2<G, 4, 1>(p, s);
  else launch_persist_bwd2<G, 4, 4>(p, s);
}

int rnn_layer_fwd_tc(int mode, int T, int B, int H, int ndir, float* gates, const float* w_hh,
                     const float* b_hh, const int64_t* lengths, const float* h0, const float* c0,
                     float* out, float* stash, float* h_final, const slnlp_rnn_extras* ex, cudaStream_t s) {
  if (!tc_shape_ok(H, w_hh)) return -1;
  const bool cat = ex && ex->hfinal_cat;
  PersistFwd p{T, B, ndir, gates, w_hh, b_hh, lengths, h0, c0, out, stash, h_final,
               cat ? (int64_t)H : (int64_t)B * H, cat ? (int64_t)ndir * H : (int64_t)H,
               ex ? ex->out_drop : nullptr, ex ? ex->p_drop : 0.f, ex ? ex->rng : nullptr, ex ? ex->site : 0u,
               ex ? ex->mask : nullptr};
  if (mode == SLNLP_MODE_LSTM) launch_persist_fwd<4>(p, s); else launch_persist_fwd<3>(p, s);
  SLNLP_LAUNCH_OK("rnn_layer_fwd(tcgen05)");
  return 0;
}

int rnn_layer_bwd_tc(int mode, int T, int B, int H, int ndir, float* gates, float* stash, const float* out,
                     const float* w_hh, const int64_t* lengths, const float* h0, const float* c0,
                     const float* dout, const float* dh_final, const float* dc_final, float* dh0, float* dc0,
                     const slnlp_rnn_extras* ex, cudaStream_t s) {
  if (!tc_shape_ok(H, w_hh)) return -1;
  const bool cat = ex && ex->hfinal_cat;
  const bool undrop = ex && ex->dout_dropped && ex->p_drop > 0.f;
  PersistBwd p{T, B, ndir, gates, stash, out, w_hh, lengths, h0, c0, dout, dh_final, dc_final, dh0, dc0,
               cat ? (int64_t)H : (int64_t)B * H, cat ? (int64_t)ndir * H : (int64_t)H,
               undrop ? ex->p_drop : 0.f, undrop ? ex->rng : nullptr, undrop ? ex->site : 0u, undrop ? ex->mask : nullptr};
  if (mode == SLNLP_MODE_LSTM) launch_persist_bwd<4>(p, s); else launch_persist_bwd<3>(p, s);
  SLNLP_LAUNCH_OK("rnn_layer_bwd(tcgen05)");
  return 0;
}

int persist_debug_config(int variant, int profile, uint32_t* out48) {
  if (variant >= 0) g_persist_var = variant;
  if (profile >= 0) g_persist_profile_on = profile;
  if (out48) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpyFromSymbol(out48, g_persist_prof, sizeof(uint32_t) * 2 * 3 * NPHASE);
    if (e != cudaSuccess) return fail("persist profile read-back: %s", cudaGetErrorString(e));
  }
  return 0;
}

}  // namespace slnlp

extern "C" int slnlp_debug_persist_config(int variant, int profile, uint32_t* out48) {
  return slnlp::persist_debug_config(variant, profile, out48);
}
