// K4 on the 5th-gen tensor cores: persistent recurrent layer for H = 128.
//
// One CTA owns one direction and a slice of 4 (16 at large batch) sequences for ALL timesteps
// (the recurrence is independent across the batch, so no CTA ever waits for another):
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
#include "tc05.cuh"

namespace slnlp {

constexpr int PH = 128;   // hidden size handled by this kernel
constexpr int PN = 16;    // MMA N (the smallest N of an M = 128 instruction)
// PSEQ (template): sequences a CTA actually owns; columns PSEQ..PN-1 of the B tile stay zero.  The per-step
// gate math and stores scale with PSEQ, and at the reference's batch of 50 most SMs are idle, so small
// batches run 4 sequences per CTA (26 CTAs at B = 50) and large ones 16 (fewer W_hh loads, fewer waves).
constexpr int PTHREADS = 512;  // 16 warps: TMEM lane quadrant = warp % 4, column group = warp / 4
constexpr int NACC = 4;        // BPTT: partial accumulators (independent MMA chains)
constexpr int FACC = 2;        // forward: partial accumulators per gate tile (two independent MMA chains)
constexpr int A_COL0 = 128;    // TMEM: accumulators in columns [0, 128), the resident W_hh operand from 128 on
constexpr int TMEM_COLS = 512; // 64 + up to 256 operand columns -> the whole tensor memory of the SM

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
};

template <int G, int PSEQ>
__global__ void __launch_bounds__(PTHREADS, 1) rnn_persistent_fwd_kernel(PersistFwd p) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int PC = PSEQ / 4;   // batch columns per thread
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int H = PH;
  uint8_t* sH = smem_raw;                       // [PN x 128] bf16, canonical (B operand)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sH + PN * H * 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, warp_u = warp_uniform();
  const int d = blockIdx.y, b0 = blockIdx.x * PSEQ;
  const int T = p.T, B = p.B;
  const int j = tid & (H - 1);   // hidden unit = TMEM lane
  const int cg = tid >> 7;       // column group: batch columns cg*PC .. cg*PC+PC-1

  // ---- one-time setup: h tile = h0 or 0 (W_hh goes to tensor memory below)
  const float* W = p.w_hh + (int64_t)d * G * H * H;
  float hreg[PC], creg[PC];
  int len[PC];
#pragma unroll
  for (int c = 0; c < PC; ++c) {
    const int n = cg * PC + c, b = b0 + n;
    hreg[c] = (p.h0 && b < B) ? p.h0[((int64_t)d * B + b) * H + j] : 0.f;
    creg[c] = (p.c0 && b < B) ? p.c0[((int64_t)d * B + b) * H + j] : 0.f;
    len[c] = b < B ? (p.lengths ? (int)p.lengths[b] : T) : 0;
    *reinterpret_cast<__nv_bfloat16*>(sH + canon_off(n, j, PN)) = __float2bfloat16(hreg[c]);
  }
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
  if (cg < G) {
    const float* wrow = W + ((int64_t)cg * H + j) * H;
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
      uint32_t pk[16];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(wrow + i * 32) + u);
        pk[2 * u] = pack2_bf16(v.x, v.y);
        pk[2 * u + 1] = pack2_bf16(v.z, v.w);
      }
      tmem_st16(tmem + ((uint32_t)((warp & 3) * 32) << 16) + A_COL0 + cg * 64 + i * 16, pk);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  constexpr uint32_t idesc = make_idesc(128, PN);
  const uint64_t descB0 = make_desc(smem_u32(sH), PN * 16, 128);
  const bool have_state0 = p.h0 != nullptr;
  const float* bh = p.b_hh + (int64_t)d * G * H;
  float bias[G];
#pragma unroll
  for (int g = 0; g < G; ++g) bias[g] = bh[g * H + j];

  // per-column base pointers; a timestep only adds t * (row stride) to them
  const int64_t gstride = (int64_t)B * p.ndir * G * H, ostride = (int64_t)B * p.ndir * H;
  float* gbase[PC];
  float* obase[PC];
  float* sbase[PC];
  float* fin[PC];
  uint32_t hoff[PC];
  bool valid[PC];
#pragma unroll
  for (int c = 0; c < PC; ++c) {
    const int n = cg * PC + c;
    valid[c] = b0 + n < B;
    const int b = valid[c] ? b0 + n : 0;
    gbase[c] = p.gates + ((int64_t)b * p.ndir + d) * G * H + j;
    obase[c] = p.out + (int64_t)b * p.ndir * H + (int64_t)d * H + j;
    sbase[c] = p.stash + ((int64_t)b * p.ndir + d) * H + j;
    fin[c] = p.h_final ? p.h_final + ((int64_t)d * B + b) * H + j : nullptr;
    hoff[c] = canon_off(n, j, PN);
  }
  // the h tile rows PSEQ..PN-1 are never written: zero them once
  for (int e = tid; e < PN * H / 8; e += PTHREADS) {
    const int n = e % PN;   // canonical tile: 16-byte core rows, row index = (e % (PN)) within a K-group
    if (n >= PSEQ) reinterpret_cast<uint4*>(sH)[e] = make_uint4(0, 0, 0, 0);
  }
  fence_async_smem();
  __syncthreads();
  // hoisted input projection, loaded one step ahead into a second register set
  float xg[G][PC], xn[G][PC];
  auto load_x = [&](float (&x)[G][PC], int t) {
    const int64_t off = (int64_t)t * gstride;
#pragma unroll
    for (int c = 0; c < PC; ++c) {
      const bool act = t < len[c];
#pragma unroll
      for (int g = 0; g < G; ++g) x[g][c] = act ? gbase[c][off + g * H] : 0.f;
    }
  };
  load_x(xg, d == 0 ? 0 : T - 1);
  // results of a step are written to HBM one iteration later, while the next step's MMAs run
  float gout[G][PC], hv[PC], sv[PC];
  auto store_step = [&](int t) {
    const int64_t goff = (int64_t)t * gstride, ooff = (int64_t)t * ostride;
#pragma unroll
    for (int c = 0; c < PC; ++c) {
      if (!valid[c]) continue;
      if (t >= len[c]) {
        obase[c][ooff] = 0.f;
        sbase[c][ooff] = 0.f;
        continue;
      }
#pragma unroll
      for (int g = 0; g < G; ++g) gbase[c][goff + g * H] = gout[g][c];
      sbase[c][ooff] = sv[c];
      obase[c][ooff] = hv[c];
      if (fin[c] && (d == 0 ? t == len[c] - 1 : t == 0)) *fin[c] = hv[c];
    }
  };

  uint32_t phase = 0;
  int t_prev = 0;
  for (int step = 0; step < T; ++step) {
    const int t = d == 0 ? step : T - 1 - step;
    const bool do_mma = step > 0 || have_state0;
    if (do_mma && warp_u == 0 && elect_one()) {
      // G gate tiles x (H/16) k-steps, A = W_hh from tensor memory, B = the h tile in shared memory;
      // issued k-major so that consecutive MMAs accumulate into different gate tiles
#pragma unroll
      for (int kk = 0; kk < H / 16; ++kk)
#pragma unroll
        for (int g = 0; g < G; ++g)
          umma_bf16_ts(tmem + (g * FACC + (kk % FACC)) * PN, tmem + A_COL0 + g * 64 + kk * 8,
                       descB0 + (uint64_t)((kk * 2 * (PN * 16)) >> 4), idesc, kk >= FACC ? 1u : 0u);
      umma_commit(bar);
    }
    __syncwarp();
    // off the critical path (the tensor core is busy): previous step -> HBM, next step's x <- HBM
    if (step > 0) store_step(t_prev);
    if (step + 1 < T) load_x(xn, d == 0 ? step + 1 : T - 2 - step);
    float acc[G][PC];
    if (do_mma) {
      mbar_wait(bar, phase);
      phase ^= 1;
      tc_fence_after();
      uint32_t raw[G * FACC][PC];
#pragma unroll
      for (int g = 0; g < G * FACC; ++g) tmem_ldn_nowait<PC>(tmem_mine + g * PN, raw[g]);
      tmem_wait_ld();
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int c = 0; c < PC; ++c) {
          float a = 0.f;
#pragma unroll
          for (int q = 0; q < FACC; ++q) a += __uint_as_float(raw[g * FACC + q][c]);
          acc[g][c] = a;
        }
    } else {
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int c = 0; c < PC; ++c) acc[g][c] = 0.f;
    }
#pragma unroll
    for (int c = 0; c < PC; ++c) {
      if (G == 4) {
        const float gi = sigmoid_fast(xg[0][c] + acc[0][c] + bias[0]);
        const float gf = sigmoid_fast(xg[1][c] + acc[1][c] + bias[1]);
        const float gg = tanh_fast(xg[2][c] + acc[2][c] + bias[2]);
        const float go = sigmoid_fast(xg[G - 1][c] + acc[G - 1][c] + bias[G - 1]);
        const float cn = gf * creg[c] + gi * gg;
        hv[c] = go * tanh_fast(cn);
        sv[c] = cn;
        gout[0][c] = gi; gout[1][c] = gf; gout[2][c] = gg; gout[G - 1][c] = go;
      } else {
        const float hn = acc[2][c] + bias[2];
        const float gr = sigmoid_fast(xg[0][c] + acc[0][c] + bias[0]);
        const float gz = sigmoid_fast(xg[1][c] + acc[1][c] + bias[1]);
        const float gn = tanh_fast(xg[2][c] + gr * hn);
        hv[c] = (1.f - gz) * gn + gz * hreg[c];
        sv[c] = hn;
        gout[0][c] = gr; gout[1][c] = gz; gout[2][c] = gn;
      }
      // state update + next step's B tile: the only stores the next MMA waits for
      if (valid[c] && t < len[c]) {
        if (G == 4) creg[c] = sv[c];
        hreg[c] = hv[c];
        *reinterpret_cast<__nv_bfloat16*>(sH + hoff[c]) = __float2bfloat16(hv[c]);
      }
    }
    // h tile (generic-proxy stores) -> visible to the tensor core; accumulators free to overwrite
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    t_prev = t;
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int c = 0; c < PC; ++c) xg[g][c] = xn[g][c];
  }
  store_step(t_prev);
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
};

template <int G, int PSEQ>
__global__ void __launch_bounds__(PTHREADS, 1) rnn_persistent_bwd_kernel(PersistBwd p) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int PC = PSEQ / 4;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int H = PH, GH = G * PH;
  uint8_t* sD = smem_raw;                        // B' = dG: [PN x GH] bf16, canonical
  uint64_t* bar = reinterpret_cast<uint64_t*>(sD + PN * GH * 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, warp_u = warp_uniform();
  const int d = blockIdx.y, b0 = blockIdx.x * PSEQ;
  const int T = p.T, B = p.B;
  const int k = tid & (H - 1);  // hidden unit = TMEM lane = output row of W_hh^T
  const int cg = tid >> 7;

  const float* W = p.w_hh + (int64_t)d * GH * H;
  for (int e = tid; e < PN * GH / 8; e += PTHREADS) reinterpret_cast<uint4*>(sD)[e] = make_uint4(0, 0, 0, 0);
  if (tid == 0) mbar_init(bar, 1);
  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_mine = tmem + ((uint32_t)((warp & 3) * 32) << 16) + cg * PC;
  // A' = W_hh^T -> tensor memory: lane = hidden unit k, K index = gate row j' (GH of them) = GH/2 packed
  // columns from A_COL0.  Column group cg fills rows j' in [cg*GH/4, (cg+1)*GH/4): global reads are
  // coalesced across the warp (consecutive k).
  {
    constexpr int JQ = GH / 4;   // 32 G gate rows -> 16 G packed columns = G stores of 16 columns
#pragma unroll 1
    for (int i = 0; i < G; ++i) {
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  constexpr uint32_t idesc = make_idesc(128, PN);
  const uint64_t descB0 = make_desc(smem_u32(sD), PN * 16, 128);

  int len[PC];
  float carry[PC];  // LSTM: dc carry; GRU: direct dh carry (dh * z)
#pragma unroll
  for (int c = 0; c < PC; ++c) {
    const int b = b0 + cg * PC + c;
    len[c] = b < B ? (p.lengths ? (int)p.lengths[b] : T) : 0;
    carry[c] = 0.f;
  }

  // per-column base pointers; a timestep only adds t * (row stride) to them
  const int64_t gstride = (int64_t)B * p.ndir * GH, ostride = (int64_t)B * p.ndir * H;
  float* gbase[PC];
  float* sbase[PC];
  const float* obase[PC];
  const float* dbase[PC];
  int64_t cidx[PC];
  uint32_t doff[PC];
  bool valid[PC];
#pragma unroll
  for (int c = 0; c < PC; ++c) {
    const int n = cg * PC + c;
    valid[c] = b0 + n < B;
    const int b = valid[c] ? b0 + n : 0;
    gbase[c] = p.gates + ((int64_t)b * p.ndir + d) * GH + k;
    sbase[c] = p.stash + ((int64_t)b * p.ndir + d) * H + k;
    obase[c] = p.out + (int64_t)b * p.ndir * H + (int64_t)d * H + k;
    dbase[c] = p.dout ? p.dout + (int64_t)b * p.ndir * H + (int64_t)d * H + k : nullptr;
    cidx[c] = ((int64_t)d * B + b) * H + k;
    doff[c] = canon_off(n, k, PN);
  }
  // per-step operands, loaded one step ahead into a second register set: activated gates, stash,
  // predecessor state, dout
  struct StepIn {
    float g[G][PC], s[PC], pv[PC], d[PC];
  };
  StepIn cur, nxt;
  auto load_step = [&](StepIn& in, int t) {
    const int tp = d == 0 ? t - 1 : t + 1;
    const bool has_prev = tp >= 0 && tp < T;
    const int64_t goff = (int64_t)t * gstride, ooff = (int64_t)t * ostride, poff = (int64_t)tp * ostride;
#pragma unroll
    for (int c = 0; c < PC; ++c) {
      const bool act = t < len[c];
#pragma unroll
      for (int g = 0; g < G; ++g) in.g[g][c] = act ? gbase[c][goff + g * H] : 0.f;
      in.s[c] = act ? sbase[c][ooff] : 0.f;
      in.d[c] = (act && dbase[c]) ? dbase[c][ooff] : 0.f;
      float pv = 0.f;
      if (act) {
        if (G == 4) pv = has_prev ? sbase[c][poff] : (p.c0 ? p.c0[cidx[c]] : 0.f);
        else pv = has_prev ? obase[c][poff] : (p.h0 ? p.h0[cidx[c]] : 0.f);
      }
      in.pv[c] = pv;
    }
  };
  load_step(cur, d == 0 ? T - 1 : 0);
  // d(pre-activations) of a step go to HBM one iteration later, while the next step's MMAs run
  float dg[G][PC], dst[PC];
  auto store_step = [&](int t) {
    const int64_t goff = (int64_t)t * gstride, ooff = (int64_t)t * ostride;
#pragma unroll
    for (int c = 0; c < PC; ++c) {
      if (!valid[c]) continue;
#pragma unroll
      for (int g = 0; g < G; ++g) gbase[c][goff + g * H] = dg[g][c];
      if (G == 3) sbase[c][ooff] = dst[c];
    }
  };

  uint32_t phase = 0;
  const int nsteps = T + ((p.dh0 || p.dc0) ? 1 : 0);
  int t_prev = 0;
  bool pending = false;
  for (int step = 0; step < nsteps; ++step) {
    const bool final_only = step == T;
    const int t = final_only ? (d == 0 ? -1 : T) : (d == 0 ? T - 1 - step : step);
    const bool do_mma = step > 0;
    if (do_mma && warp_u == 0 && elect_one()) {
      // A' = W_hh^T from tensor memory, B' = the dG tile; NACC partial accumulators: consecutive
      // MMAs are independent, the epilogue adds them
#pragma unroll
      for (int kk = 0; kk < GH / 16; ++kk)
        umma_bf16_ts(tmem + (kk % NACC) * PN, tmem + A_COL0 + kk * 8,
                     descB0 + (uint64_t)((kk * 2 * (PN * 16)) >> 4), idesc, kk >= NACC ? 1u : 0u);
      umma_commit(bar);
    }
    __syncwarp();
    // off the critical path: previous step's gradients -> HBM, next step's operands <- HBM
    if (pending) {
      store_step(t_prev);
      pending = false;
    }
    if (step + 1 < T) load_step(nxt, d == 0 ? T - 2 - step : step + 1);
    float m[PC];
    if (do_mma) {
      mbar_wait(bar, phase);
      phase ^= 1;
      tc_fence_after();
      uint32_t raw[NACC][PC];
#pragma unroll
      for (int a = 0; a < NACC; ++a) tmem_ldn_nowait<PC>(tmem_mine + a * PN, raw[a]);
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < PC; ++c) {
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < NACC; ++a) acc += __uint_as_float(raw[a][c]);
        m[c] = acc;
      }
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
      break;
    }
#pragma unroll
    for (int c = 0; c < PC; ++c) {
      dst[c] = 0.f;
      if (t >= len[c]) {
#pragma unroll
        for (int g = 0; g < G; ++g) dg[g][c] = 0.f;
      } else {
        const bool inject = d == 0 ? t == len[c] - 1 : t == 0;
        float dh = cur.d[c];
        if (G == 4) {
          float dc_in;
          if (inject) {
            dh += p.dh_final ? p.dh_final[cidx[c]] : 0.f;
            dc_in = p.dc_final ? p.dc_final[cidx[c]] : 0.f;
          } else {
            dh += m[c];
            dc_in = carry[c];
          }
          const float gi = cur.g[0][c], gf = cur.g[1][c], gg = cur.g[2][c], go = cur.g[G - 1][c];
          const float tc = tanh_fast(cur.s[c]);
          const float dc = dh * go * (1.f - tc * tc) + dc_in;
          dg[0][c] = dc * gg * gi * (1.f - gi);
          dg[1][c] = dc * cur.pv[c] * gf * (1.f - gf);
          dg[2][c] = dc * gi * (1.f - gg * gg);
          dg[G - 1][c] = dh * tc * go * (1.f - go);
          carry[c] = dc * gf;
        } else {
          if (inject) dh += p.dh_final ? p.dh_final[cidx[c]] : 0.f;
          else dh += m[c] + carry[c];
          const float gr = cur.g[0][c], gz = cur.g[1][c], gn = cur.g[2][c];
          const float da_n = dh * (1.f - gz) * (1.f - gn * gn);
          dg[0][c] = da_n * cur.s[c] * gr * (1.f - gr);
          dg[1][c] = dh * (cur.pv[c] - gn) * gz * (1.f - gz);
          dg[2][c] = da_n;
          dst[c] = da_n * gr;
          carry[c] = dh * gz;
        }
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
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    t_prev = t;
    pending = true;
    cur = nxt;
  }
  if (pending) store_step(t_prev);
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

static size_t persist_fwd_smem(int) { return (size_t)PN * PH * 2 + 64; }
static size_t persist_bwd_smem(int G) { return (size_t)PN * G * PH * 2 + 64; }

static bool tc_shape_ok(int H, const float* w_hh) { return H == PH && ((uintptr_t)w_hh & 15) == 0; }

static int seqs_per_cta(int B, int ndir) { return ceil_div(B, 4) * ndir <= (sm_count() > 0 ? sm_count() : 148) ? 4 : 16; }

template <int G, int PSEQ>
static void launch_persist_fwd(const PersistFwd& p, cudaStream_t s) {
  const size_t sm = persist_fwd_smem(G);
  launch_pdl(rnn_persistent_fwd_kernel<G, PSEQ>, dim3(dim3(ceil_div(p.B, PSEQ), p.ndir)), dim3(PTHREADS), sm, s, p);
}
template <int G, int PSEQ>
static void launch_persist_bwd(const PersistBwd& p, cudaStream_t s) {
  const size_t sm = persist_bwd_smem(G);
  launch_pdl(rnn_persistent_bwd_kernel<G, PSEQ>, dim3(dim3(ceil_div(p.B, PSEQ), p.ndir)), dim3(PTHREADS), sm, s, p);
}

int rnn_layer_fwd_tc(int mode, int T, int B, int H, int ndir, float* gates, const float* w_hh,
                     const float* b_hh, const int64_t* lengths, const float* h0, const float* c0,
                     float* out, float* stash, float* h_final, cudaStream_t s) {
  if (!tc_shape_ok(H, w_hh)) return -1;
  PersistFwd p{T, B, ndir, gates, w_hh, b_hh, lengths, h0, c0, out, stash, h_final};
  const bool small = seqs_per_cta(B, ndir) == 4;
  if (mode == SLNLP_MODE_LSTM) {
    if (small) launch_persist_fwd<4, 4>(p, s); else launch_persist_fwd<4, 16>(p, s);
  } else {
    if (small) launch_persist_fwd<3, 4>(p, s); else launch_persist_fwd<3, 16>(p, s);
  }
  SLNLP_LAUNCH_OK("rnn_layer_fwd(tcgen05)");
  return 0;
}

int rnn_layer_bwd_tc(int mode, int T, int B, int H, int ndir, float* gates, float* stash, const float* out,
                     const float* w_hh, const int64_t* lengths, const float* h0, const float* c0,
                     const float* dout, const float* dh_final, const float* dc_final, float* dh0, float* dc0,
                     cudaStream_t s) {
  if (!tc_shape_ok(H, w_hh)) return -1;
  PersistBwd p{T, B, ndir, gates, stash, out, w_hh, lengths, h0, c0, dout, dh_final, dc_final, dh0, dc0};
  const bool small = seqs_per_cta(B, ndir) == 4;
  if (mode == SLNLP_MODE_LSTM) {
    if (small) launch_persist_bwd<4, 4>(p, s); else launch_persist_bwd<4, 16>(p, s);
  } else {
    if (small) launch_persist_bwd<3, 4>(p, s); else launch_persist_bwd<3, 16>(p, s);
  }
  SLNLP_LAUNCH_OK("rnn_layer_bwd(tcgen05)");
  return 0;
}

}  // namespace slnlp
