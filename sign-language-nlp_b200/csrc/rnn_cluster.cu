// K4 for H = 256 / 512: persistent recurrent layer on a THREAD-BLOCK CLUSTER.
//
// W_hh of these sizes (bf16: 384 KB ... 2 MB) does not fit one SM, so a cluster of C = H/32 CTAs
// (8 or 16) owns one direction and one slice of sequences for ALL timesteps, and CTA c keeps the
// gate rows of hidden units [32c, 32c+32) - a [128 (gate,unit) x H] bf16 slice - resident in its
// TENSOR MEMORY as the A operand of every step's tcgen05.mma (kind::f16, fp32 accumulate):
//
//   forward : acc[(g,u), n] = sum_k W_hh[gH+u, k] h_{t-1}[n, k]; the epilogue regroups the four
//             gates of a unit through shared memory, applies the cell update, and BROADCASTS the
//             CTA's 32 new h values (bf16) into the h tile of every CTA of the cluster through
//             distributed shared memory (st.shared::cluster) - the B operand of the next step;
//   backward: the same W_hh slice, transposed in TMEM, turns the CTA's own d(gate) rows into a
//             PARTIAL dh_{t-1}[k, n] for all H units; partials are scattered to the owning CTAs
//             through distributed shared memory and summed there (an all-to-all reduce per step).
//
// One hardware cluster barrier (barrier.cluster arrive.release / wait.acquire) per step orders the
// exchange; h tiles and receive buffers are double-buffered by step parity so that a fast CTA can
// never overwrite what a slow one still reads.  No global-memory round trip and no kernel launch
// sits on the step-to-step critical path; results go to HBM off that path, one iteration later.
//
// Numerics: h, W_hh and dG are rounded to bf16 for the recurrent product (fp32 accumulation):
// the 2e-2 path of north_star.  Encoder layers only (no initial state, no dh0/dc0).
#include <cuda.h>

#include "tc05.cuh"

namespace slnlp {

constexpr int CU = 32;        // hidden units per CTA
// MMA N (template NT): 16, or 32 when a cluster owns 32 sequences
constexpr int CT = 128;       // threads: TMEM lane = thread
constexpr int CACC = 4;       // forward: partial accumulators (independent MMA chains)
constexpr int CA_COL0 = 128;  // TMEM: accumulators in [0, 128), resident operand from 128 (<= 256 columns)

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster16(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

struct ClFwd {
  int T, B, H, ndir;
  float* gates;
  const float* w_hh;
  const float* b_hh;
  const int64_t* lengths;
  float* out;
  float* stash;
  float* h_final;
};

// grid (C, ceil(B/NSEQ), ndir), cluster (C,1,1), block 128.
template <int G, int NSEQ>
__global__ void __launch_bounds__(CT, 1) rnn_cluster_fwd_kernel(ClFwd p) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int PCB = NSEQ / 4;   // sequences per thread in the (unit, sequence) phase
  constexpr int CN = NSEQ <= 16 ? 16 : 32;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int H = p.H, B = p.B, T = p.T;
  const int C = H / CU;
  uint8_t* sB = smem_raw;                                          // 2 x [16 x H] bf16, canonical K-major
  const int tile_bytes = CN * H * 2;
  float* raw = reinterpret_cast<float*>(smem_raw + 2 * tile_bytes);  // [NSEQ][4][32]
  __nv_bfloat16* hst = reinterpret_cast<__nv_bfloat16*>(raw + NSEQ * 4 * 32);  // [NSEQ][32]
  uint64_t* bar = reinterpret_cast<uint64_t*>(hst + NSEQ * 32);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, warp_u = warp_uniform();
  const uint32_t c = cluster_rank();
  const int d = blockIdx.z, b0 = blockIdx.y * NSEQ, u0 = (int)c * CU;
  const float* W = p.w_hh + (int64_t)d * G * H * H;

  for (int e = tid; e < 2 * tile_bytes / 16; e += CT) reinterpret_cast<uint4*>(sB)[e] = make_uint4(0, 0, 0, 0);
  if (tid == 0) mbar_init(bar, 1);
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  // A operand: row r = (gate warp, unit lane) of this CTA's slice = W_hh[gate*H + u0 + unit][0..H)
  {
    const bool real = warp < G;
    const float* wrow = W + ((int64_t)(real ? warp : 0) * H + u0 + lane) * H;
    for (int i = 0; i < H / 32; ++i) {
      uint32_t pk[16];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float4 v = real ? __ldg(reinterpret_cast<const float4*>(wrow + i * 32) + u) : make_float4(0.f, 0.f, 0.f, 0.f);
        pk[2 * u] = pack2_bf16(v.x, v.y);
        pk[2 * u + 1] = pack2_bf16(v.z, v.w);
      }
      tmem_st16(lane_base + CA_COL0 + i * 16, pk);
    }
    tmem_wait_st();
  }
  const float bias_r = warp < G ? p.b_hh[(int64_t)d * G * H + warp * H + u0 + lane] : 0.f;
  tc_fence_before();
  cluster_sync_all();          // every CTA's h tiles are zeroed before anyone writes into them
  tc_fence_after();

  // (unit, sequence) role of this thread: unit = lane, sequences n = warp*PCB + i
  int len[PCB];
  bool valid[PCB];
  float hreg[PCB], creg[PCB];
  float *gbase[PCB], *obase[PCB], *sbase[PCB], *fin[PCB];
#pragma unroll
  for (int i = 0; i < PCB; ++i) {
    const int n = warp * PCB + i, b = b0 + n;
    valid[i] = b < B;
    const int bb = valid[i] ? b : 0;
    len[i] = valid[i] ? (p.lengths ? (int)p.lengths[b] : T) : 0;
    hreg[i] = creg[i] = 0.f;
    gbase[i] = p.gates + ((int64_t)bb * p.ndir + d) * G * H + u0 + lane;
    obase[i] = p.out + (int64_t)bb * p.ndir * H + (int64_t)d * H + u0 + lane;
    sbase[i] = p.stash + ((int64_t)bb * p.ndir + d) * H + u0 + lane;
    fin[i] = p.h_final ? p.h_final + ((int64_t)d * B + bb) * H + u0 + lane : nullptr;
  }
  const int64_t gstride = (int64_t)B * p.ndir * G * H, ostride = (int64_t)B * p.ndir * H;
  float xg[PCB][G], xn[PCB][G];
  auto load_x = [&](float (&x)[PCB][G], int t) {
    const int64_t off = (int64_t)t * gstride;
#pragma unroll
    for (int i = 0; i < PCB; ++i) {
      const bool act = t < len[i];
#pragma unroll
      for (int g = 0; g < G; ++g) x[i][g] = act ? gbase[i][off + g * H] : 0.f;
    }
  };
  load_x(xg, d == 0 ? 0 : T - 1);
  float gout[PCB][G], hv[PCB], sv[PCB];
  auto store_step = [&](int t) {
    const int64_t goff = (int64_t)t * gstride, ooff = (int64_t)t * ostride;
#pragma unroll
    for (int i = 0; i < PCB; ++i) {
      if (!valid[i]) continue;
      if (t >= len[i]) {
        obase[i][ooff] = 0.f;
        sbase[i][ooff] = 0.f;
        continue;
      }
#pragma unroll
      for (int g = 0; g < G; ++g) gbase[i][goff + g * H] = gout[i][g];
      sbase[i][ooff] = sv[i];
      obase[i][ooff] = hv[i];
      if (fin[i] && (d == 0 ? t == len[i] - 1 : t == 0)) *fin[i] = hv[i];
    }
  };

  constexpr uint32_t idesc = make_idesc(128, CN);
  const uint32_t sB_addr = smem_u32(sB);
  uint32_t phase = 0;
  int t_prev = 0;
  for (int step = 0; step < T; ++step) {
    const int t = d == 0 ? step : T - 1 - step;
    const int in = (step & 1) ^ 1, outb = step & 1;   // step reads tile `in` (h_{t-1}), everyone writes h_t into `outb`
    const bool do_mma = step > 0;
    if (do_mma && warp_u == 0 && elect_one()) {
      const uint64_t descB = make_desc(sB_addr + in * tile_bytes, CN * 16, 128);
      for (int kk = 0; kk < H / 16; ++kk)
        umma_bf16_ts(tmem + (kk % CACC) * CN, tmem + CA_COL0 + kk * 8, descB + (uint64_t)((kk * 2 * (CN * 16)) >> 4), idesc,
                     kk >= CACC ? 1u : 0u);
      umma_commit(bar);
    }
    __syncwarp();
    if (step > 0) store_step(t_prev);
    if (step + 1 < T) load_x(xn, d == 0 ? step + 1 : T - 2 - step);
    // phase A: thread = accumulator row (gate = warp, unit = lane): partial sums + b_hh -> raw[n][gate][unit]
    {
      float a[NSEQ];
#pragma unroll
      for (int n = 0; n < NSEQ; ++n) a[n] = bias_r;
      if (do_mma) {
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
#pragma unroll
        for (int q = 0; q < CACC; ++q) {
          uint32_t r4[NSEQ];
#pragma unroll
          for (int n = 0; n < NSEQ; n += 4) tmem_ld4_nowait(lane_base + q * CN + n, reinterpret_cast<uint32_t(&)[4]>(r4[n]));
          tmem_wait_ld();
#pragma unroll
          for (int n = 0; n < NSEQ; ++n) a[n] += __uint_as_float(r4[n]);
        }
      }
#pragma unroll
      for (int n = 0; n < NSEQ; ++n) raw[(n * 4 + warp) * 32 + lane] = a[n];
    }
    __syncthreads();
    // phase B: thread = (unit = lane, sequences warp*PCB + i): gates, cell update, new h
#pragma unroll
    for (int i = 0; i < PCB; ++i) {
      const int n = warp * PCB + i;
      const float* rw = raw + n * 4 * 32 + lane;
      if (G == 4) {
        const float gi = sigmoid_fast(xg[i][0] + rw[0]);
        const float gf = sigmoid_fast(xg[i][1] + rw[32]);
        const float gg = tanh_fast(xg[i][2] + rw[64]);
        const float go = sigmoid_fast(xg[i][G - 1] + rw[(G - 1) * 32]);
        const float cn = gf * creg[i] + gi * gg;
        hv[i] = go * tanh_fast(cn);
        sv[i] = cn;
        gout[i][0] = gi; gout[i][1] = gf; gout[i][2] = gg; gout[i][G - 1] = go;
      } else {
        const float hn = rw[64];
        const float gr = sigmoid_fast(xg[i][0] + rw[0]);
        const float gz = sigmoid_fast(xg[i][1] + rw[32]);
        const float gn = tanh_fast(xg[i][2] + gr * hn);
        hv[i] = (1.f - gz) * gn + gz * hreg[i];
        sv[i] = hn;
        gout[i][0] = gr; gout[i][1] = gz; gout[i][2] = gn;
      }
      if (valid[i] && t < len[i]) {
        if (G == 4) creg[i] = sv[i];
        hreg[i] = hv[i];
      }
      hst[n * 32 + lane] = __float2bfloat16(hreg[i]);   // frozen sequences re-send their old h
    }
    __syncwarp();
    // broadcast this warp's PCB x 32 new h values to the h tile `outb` of every CTA of the cluster:
    // 16-byte chunks (8 consecutive k of one sequence) at canonical offset (k/8)*(CN*16) + n*16
    for (int item = lane; item < PCB * 4 * C; item += 32) {
      const int pr = item % C, j = (item / C) & 3, i = item / (4 * C);
      const int n = warp * PCB + i;
      const uint4 v = *reinterpret_cast<const uint4*>(hst + n * 32 + j * 8);
      const uint32_t local = sB_addr + outb * tile_bytes + (uint32_t)((4 * (int)c + j) * (CN * 16) + n * 16);
      st_cluster16(map_to_cta(local, (uint32_t)pr), v);
    }
    fence_proxy_async_all();
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    fence_proxy_async_all();
    t_prev = t;
#pragma unroll
    for (int i = 0; i < PCB; ++i)
#pragma unroll
      for (int g = 0; g < G; ++g) xg[i][g] = xn[i][g];
  }
  store_step(t_prev);
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

struct ClBwd {
  int T, B, H, ndir;
  float* gates;
  float* stash;
  const float* out;
  const float* w_hh;
  const int64_t* lengths;
  const float* dout;
  const float* dh_final;
  const float* dc_final;
};

// grid (C, ceil(B/NSEQ), ndir), cluster (C,1,1), block 128.
template <int G, int NSEQ>
__global__ void __launch_bounds__(CT, 1) rnn_cluster_bwd_kernel(ClBwd p) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int PCB = NSEQ / 4;
  constexpr int CN = NSEQ <= 16 ? 16 : 32;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int H = p.H, B = p.B, T = p.T;
  const int C = H / CU, MT = H / 128;     // CTAs per cluster, M tiles of the partial dh product
  uint8_t* sD = smem_raw;                 // B' tile: this CTA's d(gate rows) [16 x 128] bf16, canonical
  // partials for my units: [2][C][32][NSEQ + 4].  The +4 keeps rows 16-byte aligned and spreads the
  // per-unit rows over the banks (stride NSEQ alone puts every lane of a warp on the same bank)
  constexpr int RLD = NSEQ + 4;
  float* recv = reinterpret_cast<float*>(smem_raw + CN * 128 * 2);
  const int recv_half = C * 32 * RLD;
  uint64_t* bar = reinterpret_cast<uint64_t*>(recv + 2 * recv_half);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, warp_u = warp_uniform();
  const uint32_t c = cluster_rank();
  const int d = blockIdx.z, b0 = blockIdx.y * NSEQ, u0 = (int)c * CU;
  const float* W = p.w_hh + (int64_t)d * G * H * H;

  for (int e = tid; e < CN * 128 * 2 / 16; e += CT) reinterpret_cast<uint4*>(sD)[e] = make_uint4(0, 0, 0, 0);
  if (tid == 0) mbar_init(bar, 1);
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
  // A' tile i: lane = hidden unit k = 128 i + tid, K index j = local gate row (gate j/32, unit j%32)
  for (int i = 0; i < MT; ++i) {
    const float* wcol = W + 128 * i + tid;
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
      uint32_t pk[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int j = q * 32 + 2 * u, g = j >> 5;    // j and j+1 are in the same gate
        float lo = 0.f, hi = 0.f;
        if (g < G) {
          lo = __ldg(wcol + ((int64_t)g * H + u0 + (j & 31)) * H);
          hi = __ldg(wcol + ((int64_t)g * H + u0 + (j & 31) + 1) * H);
        }
        pk[u] = pack2_bf16(lo, hi);
      }
      tmem_st16(lane_base + CA_COL0 + i * 64 + q * 16, pk);
    }
  }
  tmem_wait_st();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();

  int len[PCB];
  bool valid[PCB];
  float carry[PCB];
  float *gbase[PCB], *sbase[PCB];
  const float *obase[PCB], *dbase[PCB];
  int64_t cidx[PCB];
#pragma unroll
  for (int i = 0; i < PCB; ++i) {
    const int n = warp * PCB + i, b = b0 + n;
    valid[i] = b < B;
    const int bb = valid[i] ? b : 0;
    len[i] = valid[i] ? (p.lengths ? (int)p.lengths[b] : T) : 0;
    carry[i] = 0.f;
    gbase[i] = p.gates + ((int64_t)bb * p.ndir + d) * G * H + u0 + lane;
    sbase[i] = p.stash + ((int64_t)bb * p.ndir + d) * H + u0 + lane;
    obase[i] = p.out + (int64_t)bb * p.ndir * H + (int64_t)d * H + u0 + lane;
    dbase[i] = p.dout ? p.dout + (int64_t)bb * p.ndir * H + (int64_t)d * H + u0 + lane : nullptr;
    cidx[i] = ((int64_t)d * B + bb) * H + u0 + lane;
  }
  const int64_t gstride = (int64_t)B * p.ndir * G * H, ostride = (int64_t)B * p.ndir * H;
  struct StepIn {
    float g[PCB][G], s[PCB], pv[PCB], dd[PCB];
  };
  StepIn cur, nxt;
  auto load_step = [&](StepIn& in, int t) {
    const int tp = d == 0 ? t - 1 : t + 1;
    const bool has_prev = tp >= 0 && tp < T;
    const int64_t goff = (int64_t)t * gstride, ooff = (int64_t)t * ostride, poff = (int64_t)tp * ostride;
#pragma unroll
    for (int i = 0; i < PCB; ++i) {
      const bool act = t < len[i];
#pragma unroll
      for (int g = 0; g < G; ++g) in.g[i][g] = act ? gbase[i][goff + g * H] : 0.f;
      in.s[i] = act ? sbase[i][ooff] : 0.f;
      in.dd[i] = (act && dbase[i]) ? dbase[i][ooff] : 0.f;
      float pv = 0.f;
      if (act && has_prev) pv = G == 4 ? sbase[i][poff] : obase[i][poff];
      in.pv[i] = pv;
    }
  };
  load_step(cur, d == 0 ? T - 1 : 0);
  float dg[PCB][G], dst[PCB];
  auto store_step = [&](int t) {
    const int64_t goff = (int64_t)t * gstride, ooff = (int64_t)t * ostride;
#pragma unroll
    for (int i = 0; i < PCB; ++i) {
      if (!valid[i]) continue;
#pragma unroll
      for (int g = 0; g < G; ++g) gbase[i][goff + g * H] = dg[i][g];
      if (G == 3) sbase[i][ooff] = dst[i];
    }
  };

  constexpr uint32_t idesc = make_idesc(128, CN);
  const uint64_t descD = make_desc(smem_u32(sD), CN * 16, 128);
  const uint32_t recv_addr = smem_u32(recv);
  uint32_t phase = 0;
  for (int step = 0; step < T; ++step) {
    const int t = d == 0 ? T - 1 - step : step;
    // 1. recurrent gradient of this step = sum over the cluster of the partials scattered last iteration
    float m[PCB];
#pragma unroll
    for (int i = 0; i < PCB; ++i) m[i] = 0.f;
    if (step > 0) {
      const float* rb = recv + ((step - 1) & 1) * recv_half + lane * RLD + warp * PCB;
      for (int src = 0; src < C; ++src) {
        if (PCB % 4 == 0) {
#pragma unroll
          for (int i = 0; i < PCB; i += 4) {
            const float4 v = *reinterpret_cast<const float4*>(rb + src * 32 * RLD + i);
            m[i] += v.x; m[i + 1] += v.y; m[i + 2] += v.z; m[i + 3] += v.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < PCB; i += 2) {
            const float2 v = *reinterpret_cast<const float2*>(rb + src * 32 * RLD + i);
            m[i] += v.x; m[i + 1] += v.y;
          }
        }
      }
    }
    // 2. cell backward for (unit = lane, sequences warp*PCB + i); d(gate rows) -> the B' tile
#pragma unroll
    for (int i = 0; i < PCB; ++i) {
      const int n = warp * PCB + i;
      dst[i] = 0.f;
      if (t >= len[i]) {
#pragma unroll
        for (int g = 0; g < G; ++g) dg[i][g] = 0.f;
      } else {
        const bool inject = d == 0 ? t == len[i] - 1 : t == 0;
        float dh = cur.dd[i];
        if (G == 4) {
          float dc_in;
          if (inject) {
            dh += p.dh_final ? p.dh_final[cidx[i]] : 0.f;
            dc_in = p.dc_final ? p.dc_final[cidx[i]] : 0.f;
          } else {
            dh += m[i];
            dc_in = carry[i];
          }
          const float gi = cur.g[i][0], gf = cur.g[i][1], gg = cur.g[i][2], go = cur.g[i][G - 1];
          const float tc = tanh_fast(cur.s[i]);
          const float dc = dh * go * (1.f - tc * tc) + dc_in;
          dg[i][0] = dc * gg * gi * (1.f - gi);
          dg[i][1] = dc * cur.pv[i] * gf * (1.f - gf);
          dg[i][2] = dc * gi * (1.f - gg * gg);
          dg[i][G - 1] = dh * tc * go * (1.f - go);
          carry[i] = dc * gf;
        } else {
          if (inject) dh += p.dh_final ? p.dh_final[cidx[i]] : 0.f;
          else dh += m[i] + carry[i];
          const float gr = cur.g[i][0], gz = cur.g[i][1], gn = cur.g[i][2];
          const float da_n = dh * (1.f - gz) * (1.f - gn * gn);
          dg[i][0] = da_n * cur.s[i] * gr * (1.f - gr);
          dg[i][1] = dh * (cur.pv[i] - gn) * gz * (1.f - gz);
          dg[i][2] = da_n;
          dst[i] = da_n * gr;
          carry[i] = dh * gz;
        }
      }
      if (valid[i]) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float hvv = (G == 3 && g == 2) ? dst[i] : dg[i][g];
          const int j = g * 32 + lane;
          *reinterpret_cast<__nv_bfloat16*>(sD + (j >> 3) * (CN * 16) + n * 16 + (j & 7) * 2) = __float2bfloat16(hvv);
        }
      }
    }
    if (step + 1 == T) break;   // dh0 is not needed for encoder layers: no product after the last step
    // 3. partial dh_{t-1}[k, n] for ALL units k from this CTA's gate rows
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp_u == 0 && elect_one()) {
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16_ts(tmem + i * CN, tmem + CA_COL0 + i * 64 + kk * 8, descD + (uint64_t)((kk * 2 * (CN * 16)) >> 4), idesc,
                       kk > 0 ? 1u : 0u);
      umma_commit(bar);
    }
    __syncwarp();
    store_step(t);
    load_step(nxt, d == 0 ? T - 2 - step : step + 1);
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    // 4. scatter: row k = 128 i + tid belongs to CTA k/32 = 4 i + warp, unit slot = lane
    for (int i = 0; i < MT; ++i) {
      uint32_t r4[NSEQ];
#pragma unroll
      for (int n = 0; n < NSEQ; n += 4) tmem_ld4_nowait(lane_base + i * CN + n, reinterpret_cast<uint32_t(&)[4]>(r4[n]));
      tmem_wait_ld();
      const uint32_t peer = (uint32_t)(4 * i + warp);
      const uint32_t local = recv_addr + (uint32_t)(((step & 1) * recv_half + ((int)c * 32 + lane) * RLD) * 4);
      const uint32_t remote = map_to_cta(local, peer);
#pragma unroll
      for (int n = 0; n < NSEQ; n += 4) st_cluster16(remote + n * 4, make_uint4(r4[n], r4[n + 1], r4[n + 2], r4[n + 3]));
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    cur = nxt;
  }
  store_step(d == 0 ? 0 : T - 1);
  // nobody may leave while a peer can still write into its shared memory
  cluster_sync_all();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <typename K>
static cudaError_t launch_cluster(K kernel, dim3 grid, int C, size_t smem, cudaStream_t s, const void* arg_struct, size_t) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(CT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  void* args[] = {const_cast<void*>(arg_struct)};
  return cudaLaunchKernelExC(&cfg, reinterpret_cast<const void*>(kernel), args);
}

static bool cluster_shape_ok(int H, int B, const void* w_hh) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("SLNLP_CLUSTER_RNN");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  return enabled && (H == 256 || H == 512) && B >= 1 && B <= 256 && ((uintptr_t)w_hh & 15) == 0;
}

// sequences per cluster: as few as keeps every cluster of the layer resident at once (16-CTA clusters
// fit one per GPC, 8-CTA clusters two: a second wave would double the layer time)
extern "C" int slnlp_max_active_clusters(int H, int nseq);
static int cluster_nseq(int H, int B) {
  // measured on B200: 15 clusters of 8 CTAs, 7 clusters of 16 CTAs (cudaOccupancyMaxActiveClusters)
  static int cap256 = 0, cap512 = 0;
  int& cap = H == 256 ? cap256 : cap512;
  if (cap == 0) {
    cap = slnlp_max_active_clusters(H, 16);
    if (cap <= 0) cap = H == 256 ? 14 : 6;
  }
  for (int nseq : {8, 16, 24, 32})
    if (ceil_div(B, nseq) * 2 <= cap) return nseq;
  return 32;
}

template <typename K>
static int prep_cluster_kernel(K kernel, int C, size_t smem) {
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 1;
  if (C > 8 && cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return 1;
  return 0;
}

// returns -1 when the shape is not covered (caller falls back to the per-step kernels)
int rnn_layer_fwd_cluster(int mode, int T, int B, int H, int ndir, float* gates, const float* w_hh, const float* b_hh,
                          const int64_t* lengths, const float* h0, const float* c0, float* out, float* stash,
                          float* h_final, cudaStream_t s) {
  if (h0 || c0 || !cluster_shape_ok(H, B, w_hh)) return -1;
  const int C = H / CU, nseq = cluster_nseq(H, B), cn = nseq <= 16 ? 16 : 32;
  ClFwd p{T, B, H, ndir, gates, w_hh, b_hh, lengths, out, stash, h_final};
  dim3 grid(C, ceil_div(B, nseq), ndir);
  const size_t sm = 2 * (size_t)cn * H * 2 + (size_t)nseq * 4 * 32 * 4 + (size_t)nseq * 32 * 2 + 64;
  cudaError_t e;
#define SLNLP_GO(GG, NS)                                                            \
  do {                                                                              \
    if (prep_cluster_kernel(rnn_cluster_fwd_kernel<GG, NS>, C, sm)) return -1;      \
    e = launch_cluster(rnn_cluster_fwd_kernel<GG, NS>, grid, C, sm, s, &p, sizeof(p)); \
  } while (0)
  if (mode == SLNLP_MODE_LSTM) {
    if (nseq == 8) SLNLP_GO(4, 8); else if (nseq == 16) SLNLP_GO(4, 16); else if (nseq == 24) SLNLP_GO(4, 24); else SLNLP_GO(4, 32);
  } else {
    if (nseq == 8) SLNLP_GO(3, 8); else if (nseq == 16) SLNLP_GO(3, 16); else if (nseq == 24) SLNLP_GO(3, 24); else SLNLP_GO(3, 32);
  }
#undef SLNLP_GO
  if (e != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  note_launches(1);
  return 0;
}

int rnn_layer_bwd_cluster(int mode, int T, int B, int H, int ndir, float* gates, float* stash, const float* out,
                          const float* w_hh, const int64_t* lengths, const float* h0, const float* c0,
                          const float* dout, const float* dh_final, const float* dc_final, float* dh0, float* dc0,
                          cudaStream_t s) {
  if (h0 || c0 || dh0 || dc0 || !cluster_shape_ok(H, B, w_hh)) return -1;
  const int C = H / CU, nseq = cluster_nseq(H, B), cn = nseq <= 16 ? 16 : 32;
  ClBwd p{T, B, H, ndir, gates, stash, out, w_hh, lengths, dout, dh_final, dc_final};
  dim3 grid(C, ceil_div(B, nseq), ndir);
  const size_t sm = (size_t)cn * 128 * 2 + 2 * (size_t)C * 32 * (nseq + 4) * 4 + 64;
  cudaError_t e;
#define SLNLP_GO(GG, NS)                                                            \
  do {                                                                              \
    if (prep_cluster_kernel(rnn_cluster_bwd_kernel<GG, NS>, C, sm)) return -1;      \
    e = launch_cluster(rnn_cluster_bwd_kernel<GG, NS>, grid, C, sm, s, &p, sizeof(p)); \
  } while (0)
  if (mode == SLNLP_MODE_LSTM) {
    if (nseq == 8) SLNLP_GO(4, 8); else if (nseq == 16) SLNLP_GO(4, 16); else if (nseq == 24) SLNLP_GO(4, 24); else SLNLP_GO(4, 32);
  } else {
    if (nseq == 8) SLNLP_GO(3, 8); else if (nseq == 16) SLNLP_GO(3, 16); else if (nseq == 24) SLNLP_GO(3, 24); else SLNLP_GO(3, 32);
  }
#undef SLNLP_GO
  if (e != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  note_launches(1);
  return 0;
}

}  // namespace slnlp

// how many clusters of the forward kernel the device can hold at once (diagnostic; used by
// profiles/ scripts and by cluster_nseq's constants).  H = 256 -> 8-CTA clusters, 512 -> 16-CTA.
extern "C" int slnlp_max_active_clusters(int H, int nseq) {
  using namespace slnlp;
  const int C = H / CU, cn = nseq <= 16 ? 16 : 32;
  const size_t sm = 2 * (size_t)cn * H * 2 + (size_t)nseq * 4 * 32 * 4 + (size_t)nseq * 32 * 2 + 64;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C, 64, 2);
  cfg.blockDim = dim3(CT);
  cfg.dynamicSmemBytes = sm;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = -1;
  auto q = [&](auto kernel) {
    if (prep_cluster_kernel(kernel, C, sm)) return;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) {
      cudaGetLastError();
      n = -1;
    }
  };
  if (nseq == 8) q(rnn_cluster_fwd_kernel<4, 8>);
  else if (nseq == 16) q(rnn_cluster_fwd_kernel<4, 16>);
  else q(rnn_cluster_fwd_kernel<4, 32>);
  return n;
}

