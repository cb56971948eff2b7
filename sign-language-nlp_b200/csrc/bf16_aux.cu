// bf16-side companions of the large-batch tensor-core path (gemm_pair.cu, rnn_step_tc.cu / rnn_step_pair.cu keep their
// operands as bf16 copies): the inter-layer dropout that emits the next layer's bf16 input directly, and the bias
// column sums over the bf16 d(pre-activations).
#include "tc05.cuh"

namespace slnlp {

// y (bf16) = dropout(x) with exactly the mask slnlp_dropout(site) draws (same Philox counter per group of 4).
// x is fp32 (x) or bf16 (xb); bits != NULL also leaves the keep mask as one bit per element (element i -> bit i % 32 of
// word i / 32; needs n % 128 == 0), which the backward kernels apply to the gradient instead of a dropout pass.
__global__ void __launch_bounds__(256) dropout_bf16_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ xb,
                                                           __nv_bfloat16* __restrict__ y, uint32_t* __restrict__ bits, int64_t n, float p,
                                                           const uint64_t* __restrict__ rng, uint32_t site) {
  pdl_wait();
  pdl_launch_dependents();
  const uint64_t seed = rng[0], step = rng[1];
  const float keep = 1.f - p, inv = 1.f / (1.f - p);
  const int lane = threadIdx.x & 31;
  // whole warps stay in the loop together (the mask words are assembled with shuffles)
  for (int64_t q0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) - lane; q0 * 4 < n; q0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = q0 + lane;
    uint32_t c[4] = {(uint32_t)q, (uint32_t)(q >> 32) ^ (site * 0x9E3779B9u), (uint32_t)step, (uint32_t)(step >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    uint32_t nib = 0;
    if (q * 4 + 3 < n) {
      float in[4];
      if (xb) {
        const uint2 v = reinterpret_cast<const uint2*>(xb)[q];
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v.x), b = *reinterpret_cast<const __nv_bfloat162*>(&v.y);
        in[0] = __low2float(a); in[1] = __high2float(a); in[2] = __low2float(b); in[3] = __high2float(b);
      } else {
        const float4 v = reinterpret_cast<const float4*>(x)[q];
        in[0] = v.x; in[1] = v.y; in[2] = v.z; in[3] = v.w;
      }
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool k = ((float)(c[j] >> 8) * (1.0f / 16777216.0f)) < keep;
        o[j] = k ? in[j] * inv : 0.f;
        nib |= (k ? 1u : 0u) << j;
      }
      uint2 pk;
      pk.x = pack2_bf16(o[0], o[1]);
      pk.y = pack2_bf16(o[2], o[3]);
      reinterpret_cast<uint2*>(y)[q] = pk;
    } else {
      for (int j = 0; j < 4; ++j) {
        const int64_t i = q * 4 + j;
        if (i < n) {
          const float xi = xb ? __bfloat162float(xb[i]) : x[i];
          y[i] = __float2bfloat16_rn(((float)(c[j] >> 8) * (1.0f / 16777216.0f)) < keep ? xi * inv : 0.f);
        }
      }
    }
    if (bits) {      // n % 128 == 0: the warp's 128 elements are four full words
      uint32_t w = nib << (4 * (lane & 7));
      w |= __shfl_xor_sync(0xffffffffu, w, 1);
      w |= __shfl_xor_sync(0xffffffffu, w, 2);
      w |= __shfl_xor_sync(0xffffffffu, w, 4);
      if ((lane & 7) == 0 && q * 4 < n) bits[q >> 3] = w;
    }
  }
}

// out[c] (+)= sum_r A[r*lda + c], A bf16: 256 columns per CTA (a warp reads 512 contiguous bytes of a row, 16 per lane,
// four rows in flight per thread), rows split over gridDim.y chunks that add with red.global.add when there are several
__global__ void __launch_bounds__(1024) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ A, int rows, int cols, int64_t lda,
                                                           float* __restrict__ out, float beta, int chunk) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ float sm_cs[];          // [32][257]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + lane * 8;
  const int r_end = min(rows, (int)(blockIdx.y + 1) * chunk);
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  auto add8 = [&](uint4 v) {
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u[j]);
      s[2 * j] += __low2float(a);
      s[2 * j + 1] += __high2float(a);
    }
  };
  if (c < cols) {      // cols % 8 == 0: a group of 8 is inside or outside
    int r = blockIdx.y * chunk + w;
    for (; r + 96 < r_end; r += 128) {
      const uint4 v0 = *reinterpret_cast<const uint4*>(A + (int64_t)r * lda + c), v1 = *reinterpret_cast<const uint4*>(A + (int64_t)(r + 32) * lda + c);
      const uint4 v2 = *reinterpret_cast<const uint4*>(A + (int64_t)(r + 64) * lda + c), v3 = *reinterpret_cast<const uint4*>(A + (int64_t)(r + 96) * lda + c);
      add8(v0); add8(v1); add8(v2); add8(v3);
    }
    for (; r < r_end; r += 32) add8(*reinterpret_cast<const uint4*>(A + (int64_t)r * lda + c));
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sm_cs[w * 257 + lane * 8 + j] = s[j];
  __syncthreads();
  if (threadIdx.x < 256) {
    const int cc = blockIdx.x * 256 + threadIdx.x;
    if (cc < cols) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) t += sm_cs[i * 257 + threadIdx.x];
      if (gridDim.y > 1) atomicAdd(out + cc, t);
      else out[cc] = (beta == 0.f ? 0.f : beta * out[cc]) + t;
    }
  }
}

}  // namespace slnlp

using namespace slnlp;

static int dropout_bf16_launch(const float* x, const uint16_t* xb, uint16_t* y, uint32_t* bits, int64_t n, float p, const uint64_t* rng,
                               uint32_t site, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG((x || xb) && y && rng && n >= 0 && p >= 0.f && p < 1.f, "dropout_bf16: bad arguments");
  SLNLP_CHECK_ARG((((uintptr_t)x & 15) | ((uintptr_t)xb & 7) | ((uintptr_t)y & 7) | ((uintptr_t)bits & 3)) == 0,
                  "dropout_bf16: x needs 16-byte, the bf16 operands 8-byte alignment");
  SLNLP_CHECK_ARG(!bits || n % 128 == 0, "dropout_bf16: the bit mask needs n a multiple of 128");
  if (n == 0) return 0;
  const int64_t groups = (n + 3) / 4;
  const int64_t cap = (int64_t)(sm_count() > 0 ? sm_count() : 148) * 8;
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((groups + 255) / 256, cap));
  launch_pdl(dropout_bf16_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), x, reinterpret_cast<const __nv_bfloat16*>(xb),
             reinterpret_cast<__nv_bfloat16*>(y), bits, n, p, rng, site);
  SLNLP_LAUNCH_OK("dropout_bf16");
  return 0;
}
extern "C" int slnlp_dropout_bf16(const float* x, uint16_t* y, int64_t n, float p, const uint64_t* rng, uint32_t site,
                                  slnlp_stream_t stream) {
  return dropout_bf16_launch(x, nullptr, y, nullptr, n, p, rng, site, stream);
}
extern "C" int slnlp_dropout_bf16_masked(const uint16_t* xb, uint16_t* y, uint32_t* keep_bits, int64_t n, float p, const uint64_t* rng,
                                         uint32_t site, slnlp_stream_t stream) {
  return dropout_bf16_launch(nullptr, xb, y, keep_bits, n, p, rng, site, stream);
}

extern "C" int slnlp_colsum_bf16(const uint16_t* A, int rows, int cols, int64_t lda, float* out, float beta, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(A && out && rows >= 0 && cols > 0 && lda >= cols, "colsum_bf16: bad arguments");
  SLNLP_CHECK_ARG(cols % 8 == 0 && lda % 8 == 0 && ((uintptr_t)A & 15) == 0, "colsum_bf16: cols, lda multiples of 8, 16-byte aligned A");
  static int use_atomic = -1;
  if (use_atomic < 0) {
    const char* e = getenv("SLNLP_SPLITK_ATOMIC");
    use_atomic = (e && e[0] == '0') ? 0 : 1;
  }
  const int cb = ceil_div(cols, 256);
  int splits = 1;
  if (use_atomic && beta == 1.f && rows >= 512) {
    splits = (sm_count() > 0 ? sm_count() : 148) / cb;      // one wave of 1024-thread CTAs (one per SM)
    const int max_splits = rows / 128;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  const int chunk = ceil_div(rows > 0 ? rows : 1, splits);
  splits = ceil_div(rows > 0 ? rows : 1, chunk);
  constexpr size_t smem = 32 * 257 * sizeof(float);
  launch_pdl(colsum_bf16_kernel, dim3(cb, splits), dim3(1024), smem, as_stream(stream), reinterpret_cast<const __nv_bfloat16*>(A), rows, cols,
             lda, out, beta, chunk);
  SLNLP_LAUNCH_OK("colsum_bf16");
  return 0;
}
