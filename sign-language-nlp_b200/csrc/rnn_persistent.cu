// K4 on the 5th-gen tensor cores: persistent recurrent layer for H = 128.
//
// One CTA owns one direction and one slice of 16 sequences for ALL timesteps (the
// recurrence is independent across the batch, so no CTA ever waits for another):
//   * W_hh (bf16) is loaded ONCE into shared memory in the canonical K-major
//     no-swizzle UMMA layout and stays resident for the whole layer;
//   * each step is D[gate rows (M=128 per gate tile), batch (N=16)] = W_hh_g . h^T
//     issued by one thread as tcgen05.mma kind::f16 (bf16 x bf16 -> fp32 in TMEM);
//     swap-AB: the gate rows are MMA-M, the batch is MMA-N;
//   * the epilogue threads (TMEM lane = hidden unit) read the 4 (3) gate accumulators
//     with tcgen05.ld, add the hoisted x W_ih^T + b_ih, apply the gate nonlinearities
//     and the cell update in fp32, keep c / h in registers across timesteps, write
//     out / stash for BPTT, and store h (bf16) straight into the next step's B tile.
// The BPTT twin runs dG . W_hh (K = G*H) the same way with W_hh^T resident.
//
// Numerics: the recurrent product rounds h and W_hh (dG in BPTT) to bf16 with fp32
// accumulation - the 2e-2 path of north_star.  Everything else is fp32.
#include <cuda_bf16.h>
#include "common.cuh"

namespace slnlp {

constexpr int PH = 128;   // hidden size handled by this kernel
constexpr int PN = 16;    // sequences per CTA (MMA N)
constexpr int PTHREADS = 128;

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor):
// [0,14) start>>4, [16,30) leading (K-direction core-matrix) byte offset>>4,
// [32,46) stride (M/N-direction core-matrix) byte offset>>4, [46,48) version = 1.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 (bit 4), a=b=bf16 (bits 7, 10),
// both operands K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// canonical K-major no-swizzle placement of element (row, k) of an operand with R rows:
// 8x8 core matrices (8 rows x 16 bytes), row-groups contiguous (SBO = 128 B), K-groups
// R/8*128 B apart (LBO).
__device__ __forceinline__ uint32_t canon_off(int row, int k, int R) {
  return (uint32_t)((k >> 3) * (R * 16) + (row >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2);
}

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

template <int G>
__global__ void __launch_bounds__(PTHREADS, 1) rnn_persistent_fwd_kernel(PersistFwd p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int H = PH;
  uint8_t* sW = smem_raw;                       // G tiles of [128 x 128] bf16, canonical
  uint8_t* sH = smem_raw + G * H * H * 2;       // [PN x 128] bf16, canonical (B operand)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sH + PN * H * 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int d = blockIdx.y, b0 = blockIdx.x * PN;
  const int T = p.T, B = p.B;
  const int j = tid;  // hidden unit = TMEM lane

  // ---- one-time setup: W_hh -> bf16 canonical tiles; h tile = h0 or 0
  const float* W = p.w_hh + (int64_t)d * G * H * H;
  for (int e = tid; e < G * H * (H / 8); e += PTHREADS) {
    const int r = e % (G * H), k8 = e / (G * H);  // consecutive threads -> consecutive rows (conflict-free stores)
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(W + (int64_t)r * H + k8 * 8));
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(W + (int64_t)r * H + k8 * 8) + 1);
    __nv_bfloat162 q0 = __floats2bfloat162_rn(v0.x, v0.y), q1 = __floats2bfloat162_rn(v0.z, v0.w);
    __nv_bfloat162 q2 = __floats2bfloat162_rn(v1.x, v1.y), q3 = __floats2bfloat162_rn(v1.z, v1.w);
    uint4 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&q0); pk.y = *reinterpret_cast<uint32_t*>(&q1);
    pk.z = *reinterpret_cast<uint32_t*>(&q2); pk.w = *reinterpret_cast<uint32_t*>(&q3);
    const int g = r / H, rr = r % H;
    *reinterpret_cast<uint4*>(sW + g * (H * H * 2) + canon_off(rr, k8 * 8, H)) = pk;
  }
  float hreg[PN], creg[PN];
#pragma unroll
  for (int n = 0; n < PN; ++n) {
    const int b = b0 + n;
    hreg[n] = (p.h0 && b < B) ? p.h0[((int64_t)d * B + b) * H + j] : 0.f;
    creg[n] = (p.c0 && b < B) ? p.c0[((int64_t)d * B + b) * H + j] : 0.f;
    *reinterpret_cast<__nv_bfloat16*>(sH + canon_off(n, j, PN)) = __float2bfloat16(hreg[n]);
  }
  if (tid == 0) mbar_init(bar, 1);
  if (warp == 0) tmem_alloc(tmem_slot, 64);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_lane = tmem + ((uint32_t)(warp * 32) << 16);
  constexpr uint32_t idesc = make_idesc(128, PN);
  const uint32_t sW_addr = smem_u32(sW), sH_addr = smem_u32(sH);
  const bool have_state0 = p.h0 != nullptr;

  int len[PN];
#pragma unroll
  for (int n = 0; n < PN; ++n) {
    const int b = b0 + n;
    len[n] = b < B ? (p.lengths ? (int)p.lengths[b] : T) : 0;
  }
  const float* bh = p.b_hh + (int64_t)d * G * H;
  float bias[G];
#pragma unroll
  for (int g = 0; g < G; ++g) bias[g] = bh[g * H + j];

  uint32_t phase = 0;
  for (int step = 0; step < T; ++step) {
    const int t = d == 0 ? step : T - 1 - step;
    const bool do_mma = step > 0 || have_state0;
    if (do_mma && tid == 0) {
      // G gate tiles x (H/16) k-steps; A K-step = 2 K-groups of H*16 bytes, B K-step = 2 K-groups of PN*16 bytes
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int kk = 0; kk < H / 16; ++kk) {
          const uint64_t da = make_desc(sW_addr + g * (H * H * 2) + kk * 2 * (H * 16), H * 16, 128);
          const uint64_t db = make_desc(sH_addr + kk * 2 * (PN * 16), PN * 16, 128);
          umma_bf16(tmem + g * PN, da, db, idesc, kk > 0 ? 1u : 0u);
        }
      umma_commit(bar);
    }
    __syncwarp();
    // hoisted input projection for this step (overlaps the MMAs): xp[t, b, d, g, j]
    float xg[G][PN];
#pragma unroll
    for (int n = 0; n < PN; ++n) {
      const int b = b0 + n;
      const bool act = t < len[n];
      const float* gt = p.gates + ((((int64_t)t * B + (act ? b : 0)) * p.ndir + d) * G) * H + j;
#pragma unroll
      for (int g = 0; g < G; ++g) xg[g][n] = act ? gt[g * H] : 0.f;
    }
    float acc[G][PN];
    if (do_mma) {
      mbar_wait(bar, phase);
      phase ^= 1;
      tc_fence_after();
#pragma unroll
      for (int g = 0; g < G; ++g) tmem_ld16(tmem_lane + g * PN, acc[g]);
    } else {
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int n = 0; n < PN; ++n) acc[g][n] = 0.f;
    }
#pragma unroll
    for (int n = 0; n < PN; ++n) {
      const int b = b0 + n;
      if (b >= B) continue;
      const int64_t row = (int64_t)t * B + b;
      float* o = p.out + row * p.ndir * H + (int64_t)d * H + j;
      float* st = p.stash + (row * p.ndir + d) * H + j;
      if (t >= len[n]) {
        *o = 0.f;
        *st = 0.f;
        continue;
      }
      float* gt = p.gates + ((row * p.ndir + d) * G) * H + j;
      float h;
      if (G == 4) {
        const float gi = sigmoidf_(xg[0][n] + acc[0][n] + bias[0]);
        const float gf = sigmoidf_(xg[1][n] + acc[1][n] + bias[1]);
        const float gg = tanhf(xg[2][n] + acc[2][n] + bias[2]);
        const float go = sigmoidf_(xg[G - 1][n] + acc[G - 1][n] + bias[G - 1]);
        const float c = gf * creg[n] + gi * gg;
        h = go * tanhf(c);
        creg[n] = c;
        gt[0] = gi; gt[H] = gf; gt[2 * H] = gg; gt[3 * H] = go;
        *st = c;
      } else {
        const float hn = acc[2][n] + bias[2];
        const float gr = sigmoidf_(xg[0][n] + acc[0][n] + bias[0]);
        const float gz = sigmoidf_(xg[1][n] + acc[1][n] + bias[1]);
        const float gn = tanhf(xg[2][n] + gr * hn);
        h = (1.f - gz) * gn + gz * hreg[n];
        gt[0] = gr; gt[H] = gz; gt[2 * H] = gn;
        *st = hn;
      }
      hreg[n] = h;
      *o = h;
      *reinterpret_cast<__nv_bfloat16*>(sH + canon_off(n, j, PN)) = __float2bfloat16(h);
      if (p.h_final && (d == 0 ? t == len[n] - 1 : t == 0)) p.h_final[((int64_t)d * B + b) * H + j] = h;
    }
    // h tile (generic-proxy stores) -> visible to the tensor core; accumulators free to overwrite
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc(tmem, 64);
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

template <int G>
__global__ void __launch_bounds__(PTHREADS, 1) rnn_persistent_bwd_kernel(PersistBwd p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int H = PH, GH = G * PH;
  uint8_t* sW = smem_raw;                        // A' = W_hh^T: [128 (k) x GH (j)] bf16, canonical
  uint8_t* sD = smem_raw + GH * H * 2;           // B' = dG: [PN x GH] bf16, canonical
  uint64_t* bar = reinterpret_cast<uint64_t*>(sD + PN * GH * 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int d = blockIdx.y, b0 = blockIdx.x * PN;
  const int T = p.T, B = p.B;
  const int k = tid;  // hidden unit = TMEM lane = output row of W_hh^T

  const float* W = p.w_hh + (int64_t)d * GH * H;
  // A'(m = kcol, kk = jrow) = W_hh[jrow][kcol]: thread = column kcol, 8 consecutive rows -> one 16-byte store
  // (global reads coalesced across the warp, shared stores 16 bytes apart: conflict-free)
#pragma unroll 2
  for (int jg = 0; jg < GH / 8; ++jg) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(W + (int64_t)(jg * 8 + u) * H + k);
    __nv_bfloat162 q0 = __floats2bfloat162_rn(v[0], v[1]), q1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 q2 = __floats2bfloat162_rn(v[4], v[5]), q3 = __floats2bfloat162_rn(v[6], v[7]);
    uint4 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&q0); pk.y = *reinterpret_cast<uint32_t*>(&q1);
    pk.z = *reinterpret_cast<uint32_t*>(&q2); pk.w = *reinterpret_cast<uint32_t*>(&q3);
    *reinterpret_cast<uint4*>(sW + canon_off(k, jg * 8, H)) = pk;
  }
  for (int e = tid; e < PN * GH / 8; e += PTHREADS) reinterpret_cast<uint4*>(sD)[e] = make_uint4(0, 0, 0, 0);
  if (tid == 0) mbar_init(bar, 1);
  if (warp == 0) tmem_alloc(tmem_slot, 32);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_lane = tmem + ((uint32_t)(warp * 32) << 16);
  constexpr uint32_t idesc = make_idesc(128, PN);
  const uint32_t sW_addr = smem_u32(sW), sD_addr = smem_u32(sD);

  int len[PN];
  float carry[PN];  // LSTM: dc carry; GRU: direct dh carry (dh * z)
#pragma unroll
  for (int n = 0; n < PN; ++n) {
    const int b = b0 + n;
    len[n] = b < B ? (p.lengths ? (int)p.lengths[b] : T) : 0;
    carry[n] = 0.f;
  }

  uint32_t phase = 0;
  const int nsteps = T + ((p.dh0 || p.dc0) ? 1 : 0);
  for (int step = 0; step < nsteps; ++step) {
    const bool final_only = step == T;
    const int t = final_only ? (d == 0 ? -1 : T) : (d == 0 ? T - 1 - step : step);
    const bool do_mma = step > 0;
    if (do_mma && tid == 0) {
#pragma unroll 4
      for (int kk = 0; kk < GH / 16; ++kk) {
        const uint64_t da = make_desc(sW_addr + kk * 2 * (H * 16), H * 16, 128);
        const uint64_t db = make_desc(sD_addr + kk * 2 * (PN * 16), PN * 16, 128);
        umma_bf16(tmem, da, db, idesc, kk > 0 ? 1u : 0u);
      }
      umma_commit(bar);
    }
    __syncwarp();
    float m[PN];
    if (do_mma) {
      mbar_wait(bar, phase);
      phase ^= 1;
      tc_fence_after();
      tmem_ld16(tmem_lane, m);
    } else {
#pragma unroll
      for (int n = 0; n < PN; ++n) m[n] = 0.f;
    }
#pragma unroll
    for (int n = 0; n < PN; ++n) {
      const int b = b0 + n;
      if (b >= B) continue;
      const int64_t cidx = ((int64_t)d * B + b) * H + k;
      if (final_only) {
        if (G == 4) {
          if (p.dh0) p.dh0[cidx] = m[n];
          if (p.dc0) p.dc0[cidx] = carry[n];
        } else if (p.dh0) {
          p.dh0[cidx] = m[n] + carry[n];
        }
        continue;
      }
      const int64_t row = ((int64_t)t * B + b) * p.ndir + d;
      float* gt = p.gates + row * GH + k;
      float* st = p.stash + row * H + k;
      float dg[G];
      float dstash = 0.f;
      if (t >= len[n]) {
#pragma unroll
        for (int g = 0; g < G; ++g) dg[g] = 0.f;
      } else {
        const bool inject = d == 0 ? t == len[n] - 1 : t == 0;
        float dh = p.dout ? p.dout[((int64_t)t * B + b) * p.ndir * H + (int64_t)d * H + k] : 0.f;
        const int tp = d == 0 ? t - 1 : t + 1;
        const bool has_prev = tp >= 0 && tp < T;
        if (G == 4) {
          float dc_in;
          if (inject) {
            dh += p.dh_final ? p.dh_final[cidx] : 0.f;
            dc_in = p.dc_final ? p.dc_final[cidx] : 0.f;
          } else {
            dh += m[n];
            dc_in = carry[n];
          }
          const float gi = gt[0], gf = gt[H], gg = gt[2 * H], go = gt[3 * H];
          const float c = *st;
          const float cp = has_prev ? p.stash[(((int64_t)tp * B + b) * p.ndir + d) * H + k]
                                    : (p.c0 ? p.c0[cidx] : 0.f);
          const float tc = tanhf(c);
          const float dc = dh * go * (1.f - tc * tc) + dc_in;
          dg[0] = dc * gg * gi * (1.f - gi);
          dg[1] = dc * cp * gf * (1.f - gf);
          dg[2] = dc * gi * (1.f - gg * gg);
          dg[G - 1] = dh * tc * go * (1.f - go);
          carry[n] = dc * gf;
        } else {
          if (inject) dh += p.dh_final ? p.dh_final[cidx] : 0.f;
          else dh += m[n] + carry[n];
          const float gr = gt[0], gz = gt[H], gn = gt[2 * H];
          const float hn = *st;
          const float hp = has_prev ? p.out[((int64_t)tp * B + b) * p.ndir * H + (int64_t)d * H + k]
                                    : (p.h0 ? p.h0[cidx] : 0.f);
          const float da_n = dh * (1.f - gz) * (1.f - gn * gn);
          dg[0] = da_n * hn * gr * (1.f - gr);
          dg[1] = dh * (hp - gn) * gz * (1.f - gz);
          dg[2] = da_n;
          dstash = da_n * gr;
          carry[n] = dh * gz;
        }
      }
#pragma unroll
      for (int g = 0; g < G; ++g) gt[g * H] = dg[g];
      if (G == 3) *st = dstash;
      // h-side gradients of this step = next step's B operand: dG[n][g*H + k]
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float hv = (G == 3 && g == 2) ? dstash : dg[g];
        *reinterpret_cast<__nv_bfloat16*>(sD + canon_off(n, g * H + k, PN)) = __float2bfloat16(hv);
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc(tmem, 32);
}

static size_t persist_fwd_smem(int G) { return (size_t)G * PH * PH * 2 + PN * PH * 2 + 64; }
static size_t persist_bwd_smem(int G) { return (size_t)G * PH * PH * 2 + (size_t)PN * G * PH * 2 + 64; }

static bool tc_shape_ok(int H, const float* w_hh) { return H == PH && ((uintptr_t)w_hh & 15) == 0; }

int rnn_layer_fwd_tc(int mode, int T, int B, int H, int ndir, float* gates, const float* w_hh,
                     const float* b_hh, const int64_t* lengths, const float* h0, const float* c0,
                     float* out, float* stash, float* h_final, cudaStream_t s) {
  if (!tc_shape_ok(H, w_hh)) return -1;
  PersistFwd p{T, B, ndir, gates, w_hh, b_hh, lengths, h0, c0, out, stash, h_final};
  dim3 grid(ceil_div(B, PN), ndir);
  if (mode == SLNLP_MODE_LSTM) {
    const size_t sm = persist_fwd_smem(4);
    cudaFuncSetAttribute(rnn_persistent_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    rnn_persistent_fwd_kernel<4><<<grid, PTHREADS, sm, s>>>(p);
  } else {
    const size_t sm = persist_fwd_smem(3);
    cudaFuncSetAttribute(rnn_persistent_fwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    rnn_persistent_fwd_kernel<3><<<grid, PTHREADS, sm, s>>>(p);
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
  dim3 grid(ceil_div(B, PN), ndir);
  if (mode == SLNLP_MODE_LSTM) {
    const size_t sm = persist_bwd_smem(4);
    cudaFuncSetAttribute(rnn_persistent_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    rnn_persistent_bwd_kernel<4><<<grid, PTHREADS, sm, s>>>(p);
  } else {
    const size_t sm = persist_bwd_smem(3);
    cudaFuncSetAttribute(rnn_persistent_bwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    rnn_persistent_bwd_kernel<3><<<grid, PTHREADS, sm, s>>>(p);
  }
  SLNLP_LAUNCH_OK("rnn_layer_bwd(tcgen05)");
  return 0;
}

}  // namespace slnlp
