// K4 at data-parallel batch sizes (BASELINE.json configs[3]: LSTM, batch 4096, H 512): one launch per timestep, as a
// PERSISTENT CTA-PAIR kernel in the SWAPPED orientation - hidden units are the MMA's M (TMEM lanes), sequences its N
// (TMEM columns) - with the LSTM cell as the epilogue.
//
// Why: the per-step kernels of rnn_step_tc.cu put a SEQUENCE on a TMEM lane, so an epilogue thread walks one row of
// gates / c / h: a warp-level access touches 32 different 16 KB-strided rows = 32 LSU wavefronts per instruction, and
// the step is bound by that (measured 99 us forward / 164 us backward per step at B 4096, tf32 or bf16 operands
// alike; a first pair kernel in the same orientation was slower still, profiles/r02_cfg4_tables.txt), on top of
// ~10 us of fixed latency in each of its seven waves of one-tile CTAs.  With a hidden UNIT on the lane, the 32 lanes
// of a warp read and write 128 contiguous bytes of one sequence's row: one wavefront per instruction, no staging.
//
//   forward : 74 resident clusters of two CTAs walk tiles of 256 units x 128 sequences; per tile FOUR accumulators
//             (one per gate, 4 x 128 = all 512 TMEM columns), A = the W_hh rows of gate g for the CTA's 128 units
//             (bf16, TMA), B = h_{t-1} of the tile's sequences (the bf16 copy of `out` the previous launch wrote;
//             CTA r stages 64 of them), tcgen05.mma.cta_group::2 kind::f16 M256 N128 K16; 16 epilogue warps per CTA
//             (thread = unit x 32 sequences): hoisted projection + c_{t-1} of the first chunk in flight before the
//             accumulator is waited for, gates + cell + length freeze, activated gates / c_t / h_t (fp32) and h_t
//             (bf16) written coalesced;
//   backward: tiles of 256 units x 128 sequences, K = 4H: A = W_hh^T rows (bf16 [H][4H]), B = dG_{t+1} (bf16, written
//             by the previous launch), accumulators double-buffered (2 x 128 columns) so that the pair's MMAs run on
//             the next tile under the cell backward of this one; d(pre-activations) leave as bf16 (every consumer -
//             the next step, the dX / dW GEMMs, the bias column sums - reads that copy; the fp32 in-place write is
//             optional: it was 27 % of the step's HBM bytes).
#include "pair.cuh"

namespace slnlp {

constexpr int RP_STAGES_F = 3, RP_STAGES_B = 8;
constexpr int RP_AF = 4 * 128 * 64 * 2;     // forward A: 4 gates x 128 units x 64 k, bf16 (64 KB)
constexpr int RP_BS = 64 * 64 * 2;          // B: 64 sequences x 64 k (8 KB)
constexpr int RP_AB = 128 * 64 * 2;         // backward A: 128 units x 64 j (16 KB)
constexpr int RP_EG = 4;                    // epilogue warp groups (4 warps each; group e owns sequences [32 e, 32 e + 32))
constexpr int RP_THREADS = 64 + 128 * RP_EG;
__host__ __device__ constexpr size_t rp_smem(int stages, int stage_bytes) { return (size_t)stages * stage_bytes + (2 * stages + 4) * 8 + 16 + 1024; }

__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}

// barriers + TMEM of a pair kernel (the protocol of gemm_pair_kernel); ASTAGE / BSTAGE = bytes per ring stage
struct PairRing {
  uint8_t *sA, *sB;
  uint64_t *full, *empty, *tfull, *tempty;
  uint32_t tmem;
};
template <int STAGES, int ASTAGE, int BSTAGE, int TCOLS>
__device__ __forceinline__ PairRing pair_setup(uint8_t* smem_dyn, int warp, const CUtensorMap* m0, const CUtensorMap* m1) {
  PairRing r;
  uint8_t* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  r.sA = base;
  r.sB = base + STAGES * ASTAGE;
  r.full = reinterpret_cast<uint64_t*>(r.sB + STAGES * BSTAGE);
  r.empty = r.full + STAGES;
  r.tfull = r.empty + STAGES;
  r.tempty = r.tfull + 2;
  uint32_t* slot = reinterpret_cast<uint32_t*>(r.tempty + 2);
  if (warp == 0 && elect_one()) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(m1) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&r.full[s], 1);
      mbar_init(&r.empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&r.tfull[a], 1);
      mbar_init(&r.tempty[a], 8 * RP_EG);
    }
  }
  if (warp == 1) tmem_alloc_pair(slot, TCOLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  r.tmem = *slot;
  return r;
}

struct PairFwd {
  int T, B, H, ndir, step, tiles_u, tiles_s, tiles;
  float* gates;
  const float* b_hh;
  const int64_t* lengths;
  float* out;
  __nv_bfloat16* out_bf;
  float* stash;
  float* h_final;
  __nv_bfloat16* gact_bf;   // optional: activated gates stashed as bf16 [T][B][ndir*4H] instead of in place in `gates`
};

// mapW: w_hh_bf as [ndir][4H][H], box {64, 128, 1}; mapH: out_bf as [T][B][ndir*H], box {64, 64, 1}
// SEQ = sequences per tile: 128 (one accumulator set = all 512 TMEM columns: the MMAs of a tile wait for the previous
// tile's epilogue) or 64 (two sets: MMAs and epilogue overlap, W_hh tiles are streamed twice as often)
template <int HT, int SEQ>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(RP_THREADS, 1)
    lstm_step_fwd_pair_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapH, PairFwd p) {
  extern __shared__ uint8_t smem_dyn[];
  constexpr int G = 4, NACC = (512 / (G * SEQ)) > 2 ? 2 : 512 / (G * SEQ), BS = SEQ * 64;   // accumulator sets (at most two); bytes of a CTA's B tile (SEQ/2 rows)
  constexpr int SPT = SEQ / RP_EG, NCH = SPT / 4;            // sequences per epilogue thread, chunks of four
  const int warp = warp_uniform(), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int H = p.H, B = p.B, T = p.T;
  const bool has_prev = p.step > 0;                  // zero initial state: the recurrent product of step 0 is exactly 0
  const int nk = has_prev ? H / 64 : 0;
  const int per_dir = p.tiles_u * p.tiles_s;
  PairRing r = pair_setup<RP_STAGES_F, RP_AF, BS, 512>(smem_dyn, warp, &mapW, &mapH);
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (elect_one() && nk > 0) {
      uint32_t it = 0;
      for (int tile = cluster_id; tile < p.tiles; tile += nclusters) {
        const int d = tile / per_dir, rem = tile - d * per_dir;
        const int ub = rem / p.tiles_s, sb = rem - ub * p.tiles_s;
        const int t = d == 0 ? p.step : T - 1 - p.step, tp = d == 0 ? t - 1 : t + 1;
        const int u0 = ub * 256 + (int)rank * 128, s0 = sb * SEQ + (int)rank * (SEQ / 2);
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const uint32_t s = it % RP_STAGES_F, round = it / RP_STAGES_F;
          if (round > 0) mbar_wait(&r.empty[s], (round - 1) & 1u);
          if (rank == 0) mbar_expect_tx(&r.full[s], 2 * (RP_AF + BS));
          const uint32_t bar = map_to_cta(smem_u32(&r.full[s]), 0);
#pragma unroll
          for (int g = 0; g < G; ++g) tma_load_3d_pair(r.sA + s * RP_AF + g * 16384, &mapW, bar, kb * 64, g * H + u0, d);
          tma_load_3d_pair(r.sB + s * BS, &mapH, bar, d * H + kb * 64, s0, tp);
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && nk > 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, SEQ, 0, 0);
      const uint64_t dA = make_desc_sw128(smem_u32(r.sA), 16, 1024, 2), dB = make_desc_sw128(smem_u32(r.sB), 16, 1024, 2);
      uint32_t it = 0, tl = 0;
      for (int tile = cluster_id; tile < p.tiles; tile += nclusters, ++tl) {
        const uint32_t as = tl % NACC, use = tl / NACC;
        if (use > 0) mbar_wait(&r.tempty[as], (use - 1) & 1u);
        tc_fence_after();
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const uint32_t s = it % RP_STAGES_F, round = it / RP_STAGES_F;
          mbar_wait(&r.full[s], round & 1u);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
              for (int g = 0; g < G; ++g)
                umma_bf16_pair(r.tmem + as * (G * SEQ) + g * SEQ, dA + (uint64_t)((s * RP_AF + g * 16384 + kk * 32) >> 4),
                               dB + (uint64_t)((s * BS + kk * 32) >> 4), idesc, (kb > 0 || kk > 0) ? 1u : 0u);
            umma_commit_pair(&r.empty[s]);
            if (kb == nk - 1) umma_commit_pair(&r.tfull[as]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ---- epilogue: lane = hidden unit, columns = sequences.  H is a template parameter: every address inside a tile is
    // a per-tile 64-bit base + a compile-time offset (the first version spent 280 instructions per element, most of
    // them 64-bit index arithmetic: 45 % issue utilisation, neither HBM- nor tensor-bound).  Sequences go in chunks
    // of four, double-buffered: the loads of chunk i+1 are in flight under the math and stores of chunk i.
    constexpr int GH = G * HT, SG = 2 * GH, SH = 2 * HT;       // floats per sequence row of gates / of stash, out
    const int q = warp & 3, e = (warp - 2) >> 2;
    const uint32_t tempty_leader = map_to_cta(smem_u32(&r.tempty[0]), 0);
    uint32_t tl = 0;
    for (int tile = cluster_id; tile < p.tiles; tile += nclusters, ++tl) {
      const int d = tile / per_dir, rem = tile - d * per_dir;
      const int ub = rem / p.tiles_s, sb = rem - ub * p.tiles_s;
      const int t = d == 0 ? p.step : T - 1 - p.step, tp = d == 0 ? t - 1 : t + 1;
      const int unit = ub * 256 + (int)rank * 128 + q * 32 + lane;      // this thread's hidden unit (TMEM lane)
      const uint32_t as = tl % NACC, use = tl / NACC;
      const int b0 = sb * SEQ + e * SPT;                                // first of its SPT sequences (TMEM columns)
      const int nvalid = min(SPT, B - b0);
      const int64_t rb = ((int64_t)t * B + b0) * 2 + d;
      float* gp = p.gates + rb * GH + unit;
      __nv_bfloat16* gab = p.gact_bf ? p.gact_bf + rb * GH + unit : nullptr;
      float* sp = p.stash + rb * HT + unit;
      const float* pp = sp + (int64_t)(tp - t) * B * SH;                // c_{t-1} of the same sequences
      float* op = p.out ? p.out + rb * HT + unit : nullptr;      // NULL: nobody reads the fp32 copy of this layer's output
      __nv_bfloat16* obp = p.out_bf + rb * HT + unit;
      float* hfp = p.h_final ? p.h_final + ((int64_t)d * B + b0) * HT + unit : nullptr;
      const int64_t* lp = p.lengths ? p.lengths + b0 : nullptr;
      float bias[G];
#pragma unroll
      for (int g = 0; g < G; ++g) bias[g] = p.b_hh[d * GH + g * HT + unit];
      float xg[2][G][4], pv[2][4];
      int len[2][4];       // -1: sequence out of range
      auto fetch = [&](int ch, int buf) {
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const int j = ch * 4 + x;
          const bool valid = j < nvalid;
          len[buf][x] = valid ? (lp ? (int)lp[j] : T) : -1;
        }
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const int j = ch * 4 + x;
          const bool act = t < len[buf][x];
#pragma unroll
          for (int g = 0; g < G; ++g) xg[buf][g][x] = act ? gp[j * SG + g * HT] : 0.f;
          pv[buf][x] = (act && has_prev) ? pp[j * SH] : 0.f;
        }
      };
      fetch(0, 0);         // in flight before the accumulator is waited for
      if (nk > 0) {
        mbar_wait(&r.tfull[as], use & 1u);
        tc_fence_after();
      }
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int buf = ch & 1;
        if (ch < NCH - 1) fetch(ch + 1, buf ^ 1);
        float acc[G][4];
        if (nk > 0) {
          uint32_t raw[G][4];
#pragma unroll
          for (int g = 0; g < G; ++g) tmem_ld4_nowait(r.tmem + ((uint32_t)(q * 32) << 16) + as * (G * SEQ) + g * SEQ + e * SPT + ch * 4, raw[g]);
          tmem_wait_ld();
#pragma unroll
          for (int g = 0; g < G; ++g)
#pragma unroll
            for (int x = 0; x < 4; ++x) acc[g][x] = __uint_as_float(raw[g][x]);
          if (ch == NCH - 1) {      // this warp's part of the accumulators is in registers: hand them back to the MMA thread
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_leader + as * 8);
          }
        } else {
#pragma unroll
          for (int g = 0; g < G; ++g)
#pragma unroll
            for (int x = 0; x < 4; ++x) acc[g][x] = 0.f;
        }
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const int j = ch * 4 + x;
          const int ln = len[buf][x];
          if (ln < 0) continue;
          if (t >= ln) {
            if (op) op[j * SH] = 0.f;
            sp[j * SH] = 0.f;
            obp[j * SH] = __float2bfloat16_rn(0.f);
            continue;
          }
          const float gi = sigmoid_fast(xg[buf][0][x] + acc[0][x] + bias[0]);
          const float gf = sigmoid_fast(xg[buf][1][x] + acc[1][x] + bias[1]);
          const float gg = tanh_fast(xg[buf][2][x] + acc[2][x] + bias[2]);
          const float go = sigmoid_fast(xg[buf][3][x] + acc[3][x] + bias[3]);
          const float cc = gf * pv[buf][x] + gi * gg;
          const float hv = go * tanh_fast(cc);
          if (gab) {     // the BPTT stash as bf16 (8 instead of 16 bytes per element, and again when it is read back)
            gab[j * SG] = __float2bfloat16_rn(gi); gab[j * SG + HT] = __float2bfloat16_rn(gf);
            gab[j * SG + 2 * HT] = __float2bfloat16_rn(gg); gab[j * SG + 3 * HT] = __float2bfloat16_rn(go);
          } else {
            gp[j * SG] = gi; gp[j * SG + HT] = gf; gp[j * SG + 2 * HT] = gg; gp[j * SG + 3 * HT] = go;
          }
          sp[j * SH] = cc;
          if (op) op[j * SH] = hv;
          obp[j * SH] = __float2bfloat16_rn(hv);
          if (hfp && (d == 0 ? t == ln - 1 : t == 0)) hfp[j * HT] = hv;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(r.tmem, 512);
}

struct PairBwd {
  int T, B, H, ndir, step, tiles_u, tiles_s, tiles, write_f32;
  float* gates;
  __nv_bfloat16* dg_bf;
  const float* stash;
  const int64_t* lengths;
  const float *dout, *dh_final, *dc_final;
  float* carry;
  const uint32_t* dout_keep;   // optional keep mask of the dropout between this layer and the next (1 bit per element of dout)
  float dout_scale;            // 1 / (1 - p)
  int gates_in_dg;             // the activated gates were stashed as bf16 IN dg_bf (forward's gact_bf): read them there
};

// mapW: w_hhT_bf as [ndir][H][4H], box {64, 128, 1}; mapG: dg_bf as [T][B][ndir*4H], box {64, 64, 1}
// SEQ = sequences per tile (128 / 64 / 32): the launcher takes the largest that still gives every cluster a tile
template <int HT, int SEQ>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(RP_THREADS, 1)
    lstm_step_bwd_pair_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapG, PairBwd p) {
  extern __shared__ uint8_t smem_dyn[];
  constexpr int G = 4, BS = SEQ * 64, SPT = SEQ / RP_EG, NCH = SPT / 4;   // B tile bytes; sequences per thread; chunks of four
  const int warp = warp_uniform(), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int H = p.H, B = p.B, T = p.T, GH = G * p.H;
  const bool has_next = p.step > 0;                 // the step processed just before this one exists
  const int nk = has_next ? GH / 64 : 0;
  const int per_dir = p.tiles_u * p.tiles_s;
  PairRing r = pair_setup<RP_STAGES_B, RP_AB, BS, (2 * SEQ < 32 ? 32 : 2 * SEQ)>(smem_dyn, warp, &mapW, &mapG);
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (elect_one() && nk > 0) {
      uint32_t it = 0;
      for (int tile = cluster_id; tile < p.tiles; tile += nclusters) {
        const int d = tile / per_dir, rem = tile - d * per_dir;
        const int ub = rem / p.tiles_s, sb = rem - ub * p.tiles_s;
        const int t = d == 0 ? T - 1 - p.step : p.step, tn = d == 0 ? t + 1 : t - 1;
        const int u0 = ub * 256 + (int)rank * 128, s0 = sb * SEQ + (int)rank * (SEQ / 2);
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const uint32_t s = it % RP_STAGES_B, round = it / RP_STAGES_B;
          if (round > 0) mbar_wait(&r.empty[s], (round - 1) & 1u);
          if (rank == 0) mbar_expect_tx(&r.full[s], 2 * (RP_AB + BS));
          const uint32_t bar = map_to_cta(smem_u32(&r.full[s]), 0);
          tma_load_3d_pair(r.sA + s * RP_AB, &mapW, bar, kb * 64, u0, d);
          tma_load_3d_pair(r.sB + s * BS, &mapG, bar, d * GH + kb * 64, s0, tn);
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && nk > 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, SEQ, 0, 0);
      const uint64_t dA = make_desc_sw128(smem_u32(r.sA), 16, 1024, 2), dB = make_desc_sw128(smem_u32(r.sB), 16, 1024, 2);
      uint32_t it = 0, tl = 0;
      for (int tile = cluster_id; tile < p.tiles; tile += nclusters, ++tl) {
        const uint32_t as = tl & 1u, use = tl >> 1;
        if (use > 0) mbar_wait(&r.tempty[as], (use - 1) & 1u);
        tc_fence_after();
        const uint32_t acc = r.tmem + as * SEQ;
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const uint32_t s = it % RP_STAGES_B, round = it / RP_STAGES_B;
          mbar_wait(&r.full[s], round & 1u);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16_pair(acc, dA + (uint64_t)((s * RP_AB + kk * 32) >> 4), dB + (uint64_t)((s * BS + kk * 32) >> 4), idesc,
                             (kb > 0 || kk > 0) ? 1u : 0u);
            umma_commit_pair(&r.empty[s]);
            if (kb == nk - 1) umma_commit_pair(&r.tfull[as]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ---- epilogue (see the forward kernel): lane = hidden unit, sequences in double-buffered chunks of four,
    // compile-time strides
    constexpr int GH = G * HT, SG = 2 * GH, SH = 2 * HT;
    const int q = warp & 3, e = (warp - 2) >> 2;
    const uint32_t tempty_leader = map_to_cta(smem_u32(&r.tempty[0]), 0);
    uint32_t tl = 0;
    for (int tile = cluster_id; tile < p.tiles; tile += nclusters, ++tl) {
      const int d = tile / per_dir, rem = tile - d * per_dir;
      const int ub = rem / p.tiles_s, sb = rem - ub * p.tiles_s;
      const int t = d == 0 ? T - 1 - p.step : p.step;
      const int tp = d == 0 ? t - 1 : t + 1;   // forward-time predecessor
      const bool has_prev = tp >= 0 && tp < T;
      const uint32_t as = tl & 1u, use = tl >> 1;
      const int unit = ub * 256 + (int)rank * 128 + q * 32 + lane;
      const int b0 = sb * SEQ + e * SPT;
      const int nvalid = min(SPT, B - b0);
      const int64_t rb = ((int64_t)t * B + b0) * 2 + d;
      float* gp = p.gates + rb * GH + unit;
      __nv_bfloat16* gbp = p.dg_bf + rb * GH + unit;
      const float* sp = p.stash + rb * HT + unit;
      const float* pp = sp + (int64_t)(tp - t) * B * SH;
      const float* dop = p.dout ? p.dout + rb * HT + unit : nullptr;
      // dout is the gradient of the DROPPED output when a keep mask comes along: 32 lanes = 32 consecutive units = one word
      const uint32_t* kp = (p.dout_keep && dop) ? p.dout_keep + ((rb * HT + (unit & ~31)) >> 5) : nullptr;
      const int64_t ci0 = ((int64_t)d * B + b0) * HT + unit;
      float* cp = p.carry + ci0;
      const int64_t* lp = p.lengths ? p.lengths + b0 : nullptr;
      float sv[2][4], pv[2][4], dh[2][4], cr[2][4];
      uint32_t gv[2][G][4], kw[2][4];
      int len[2][4];
      const float dscale = kp ? p.dout_scale : 1.f;
      auto fetch = [&](int ch, int buf) {
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const int j = ch * 4 + x;
          len[buf][x] = j < nvalid ? (lp ? (int)lp[j] : T) : -1;
        }
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const int j = ch * 4 + x;
          const bool live = t < len[buf][x];
#pragma unroll
          for (int g = 0; g < G; ++g) {      // raw bits (fp32, or bf16 zero-extended by the load): loads only (see kw)
            if (p.gates_in_dg) gv[buf][g][x] = live ? (uint32_t)reinterpret_cast<const unsigned short*>(gbp)[j * SG + g * HT] : 0u;
            else gv[buf][g][x] = live ? __float_as_uint(gp[j * SG + g * HT]) : 0u;
          }
          sv[buf][x] = live ? sp[j * SH] : 0.f;
          pv[buf][x] = (live && has_prev) ? pp[j * SH] : 0.f;
          dh[buf][x] = (live && dop) ? dop[j * SH] : 0.f;
          kw[buf][x] = (kp && live) ? kp[j * (SH / 32)] : 0xFFFFFFFFu;      // only LOADS here: arithmetic on a loaded value
                                                                            // would wait for it and end the prefetch
          cr[buf][x] = live ? cp[j * HT] : 0.f;
        }
      };
      fetch(0, 0);
      if (nk > 0) {
        mbar_wait(&r.tfull[as], use & 1u);
        tc_fence_after();
      }
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int buf = ch & 1;
        if (ch < NCH - 1) fetch(ch + 1, buf ^ 1);
        float m[4];
        if (nk > 0) {
          uint32_t raw[4];
          tmem_ld4_nowait(r.tmem + ((uint32_t)(q * 32) << 16) + as * SEQ + e * SPT + ch * 4, raw);
          tmem_wait_ld();
#pragma unroll
          for (int x = 0; x < 4; ++x) m[x] = __uint_as_float(raw[x]);
          if (ch == NCH - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_leader + as * 8);
          }
        } else {
#pragma unroll
          for (int x = 0; x < 4; ++x) m[x] = 0.f;
        }
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const int j = ch * 4 + x;
          const int ln = len[buf][x];
          if (ln < 0) continue;
          float og[G];
          if (t >= ln) {   // zero gradients for the hoisted dW / dx GEMMs, carry untouched
#pragma unroll
            for (int g = 0; g < G; ++g) og[g] = 0.f;
          } else {
            const bool inject = d == 0 ? t == ln - 1 : t == 0;
            float dhx = ((kw[buf][x] >> lane) & 1u) ? dh[buf][x] * dscale : 0.f, cin = cr[buf][x];
            if (inject) {          // final-state gradients enter here instead of the recurrent ones (warp-uniform: one sequence)
              if (p.dh_final) dhx += p.dh_final[ci0 + j * HT];
              cin = p.dc_final ? p.dc_final[ci0 + j * HT] : 0.f;
            } else {
              dhx += m[x];
            }
            const int sh = p.gates_in_dg ? 16 : 0;      // bf16 bits -> the fp32 with the same value
            const float gi = __uint_as_float(gv[buf][0][x] << sh), gf = __uint_as_float(gv[buf][1][x] << sh);
            const float gg = __uint_as_float(gv[buf][2][x] << sh), go = __uint_as_float(gv[buf][3][x] << sh);
            const float tc = tanh_fast(sv[buf][x]);
            const float dc = dhx * go * (1.f - tc * tc) + cin;
            og[0] = dc * gg * gi * (1.f - gi);
            og[1] = dc * pv[buf][x] * gf * (1.f - gf);
            og[2] = dc * gi * (1.f - gg * gg);
            og[3] = dhx * tc * go * (1.f - go);
            cp[j * HT] = dc * gf;
          }
#pragma unroll
          for (int g = 0; g < G; ++g) {
            gbp[j * SG + g * HT] = __float2bfloat16_rn(og[g]);
            if (p.write_f32) gp[j * SG + g * HT] = og[g];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(r.tmem, (2 * SEQ < 32 ? 32 : 2 * SEQ));
}

static int pair_seq_env() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SLNLP_PAIR_SEQ");
    const int x = e ? atoi(e) : 0;
    v = (x == 128 || x == 64 || x == 32) ? x : 0;
  }
  return v;
}

// hidden sizes with an instantiation (H is a template parameter of the epilogues); two directions only
static bool pair_step_ok(int T, int B, int H, int ndir) {
  return T > 1 && B >= 1 && ndir == 2 && (H == 256 || H == 512 || H == 1024) && encode_fn() != nullptr;
}

int lstm_layer_fwd_pairstep(int T, int B, int H, int ndir, float* gates, const uint16_t* w_hh_bf, const float* b_hh,
                            const int64_t* lengths, float* out, uint16_t* out_bf, float* stash, float* h_final, uint16_t* gact_bf,
                            cudaStream_t s) {
  if (!pair_step_ok(T, B, H, ndir)) return -1;
  if (((uintptr_t)w_hh_bf | (uintptr_t)out_bf) & 15) return -1;
  CUtensorMap mapW, mapH;
  if (!tensor_map3_bf16(w_hh_bf, H, (uint64_t)4 * H, ndir, H, (uint64_t)4 * H * H, 128, &mapW)) return -1;
  // sequences per tile: 64 (two accumulator sets: MMAs under the epilogue), 32 when 64 would leave clusters without a
  // tile (small per-rank batches of the data-parallel config); $SLNLP_PAIR_SEQ=128|64|32 forces one
  const int sms = sm_count() > 0 ? sm_count() : 148;
  int seq = pair_seq_env();
  if (!seq) seq = (ndir * (H / 256) * ceil_div(B, 64) >= sms / 2) ? 64 : 32;
  if (!tensor_map3_bf16(out_bf, (uint64_t)ndir * H, B, T, (uint64_t)ndir * H, (uint64_t)B * ndir * H, seq / 2, &mapH)) return -1;
  PairFwd p{T, B, H, ndir, 0, H / 256, ceil_div(B, seq), 0, gates, b_hh, lengths, out, reinterpret_cast<__nv_bfloat16*>(out_bf),
            stash, h_final, reinterpret_cast<__nv_bfloat16*>(gact_bf)};
  p.tiles = ndir * p.tiles_u * p.tiles_s;
  dim3 grid(2 * std::min(p.tiles, sms / 2));
  constexpr size_t sm = rp_smem(RP_STAGES_F, RP_AF + RP_BS);
  auto run = [&](auto kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    for (int step = 0; step < T; ++step) {
      p.step = step;
      launch_pdl(kernel, grid, dim3(RP_THREADS), sm, s, mapW, mapH, p);
    }
  };
#define SLNLP_FWD(S)                                            \
  do {                                                          \
    if (H == 256) run(lstm_step_fwd_pair_kernel<256, S>);       \
    else if (H == 512) run(lstm_step_fwd_pair_kernel<512, S>);  \
    else run(lstm_step_fwd_pair_kernel<1024, S>);               \
  } while (0)
  if (seq == 128) SLNLP_FWD(128);
  else if (seq == 64) SLNLP_FWD(64);
  else SLNLP_FWD(32);
#undef SLNLP_FWD
  note_launches(T);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("rnn_layer_fwd(pair step): launch failed: %s", cudaGetErrorString(e));
  return 0;
}

int lstm_layer_bwd_pairstep(int T, int B, int H, int ndir, float* gates, uint16_t* dg_bf, const float* stash,
                            const uint16_t* w_hhT_bf, const int64_t* lengths, const float* dout, const float* dh_final,
                            const float* dc_final, float* carry, int write_f32, const uint32_t* dout_keep, float dout_scale,
                            int gates_in_dg, cudaStream_t s) {
  if (!pair_step_ok(T, B, H, ndir)) return -1;
  if (((uintptr_t)w_hhT_bf | (uintptr_t)dg_bf) & 15) return -1;
  CUtensorMap mapW, mapG;
  if (!tensor_map3_bf16(w_hhT_bf, (uint64_t)4 * H, H, ndir, (uint64_t)4 * H, (uint64_t)4 * H * H, 128, &mapW)) return -1;
  const int sms = sm_count() > 0 ? sm_count() : 148;
  int seq = pair_seq_env();
  if (!seq) {
    seq = 32;
    for (int cand : {128, 64})
      if (ndir * (H / 256) * ceil_div(B, cand) >= sms / 2) {
        seq = cand;
        break;
      }
  }
  if (!tensor_map3_bf16(dg_bf, (uint64_t)ndir * 4 * H, B, T, (uint64_t)ndir * 4 * H, (uint64_t)B * ndir * 4 * H, seq / 2, &mapG)) return -1;
  PairBwd p{T, B, H, ndir, 0, H / 256, ceil_div(B, seq), 0, write_f32, gates, reinterpret_cast<__nv_bfloat16*>(dg_bf), stash,
            lengths, dout, dh_final, dc_final, carry, dout_keep, dout_scale, gates_in_dg};
  p.tiles = ndir * p.tiles_u * p.tiles_s;
  dim3 grid(2 * std::min(p.tiles, sms / 2));
  constexpr size_t sm = rp_smem(RP_STAGES_B, RP_AB + RP_BS);
  auto run = [&](auto kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    for (int step = 0; step < T; ++step) {
      p.step = step;
      launch_pdl(kernel, grid, dim3(RP_THREADS), sm, s, mapW, mapG, p);
    }
  };
#define SLNLP_BWD(S)                                            \
  do {                                                          \
    if (H == 256) run(lstm_step_bwd_pair_kernel<256, S>);       \
    else if (H == 512) run(lstm_step_bwd_pair_kernel<512, S>);  \
    else run(lstm_step_bwd_pair_kernel<1024, S>);               \
  } while (0)
  if (seq == 128) SLNLP_BWD(128);
  else if (seq == 64) SLNLP_BWD(64);
  else SLNLP_BWD(32);
#undef SLNLP_BWD
  note_launches(T);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("rnn_layer_bwd(pair step): launch failed: %s", cudaGetErrorString(e));
  return 0;
}

}  // namespace slnlp
