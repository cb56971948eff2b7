// bf16-side companions of the large-batch tensor-core path (gemm_pair.cu, rnn_step_tc.cu / rnn_step_pair.cu keep their
// operands as bf16 copies): the inter-layer dropout that emits the next layer's bf16 input directly, and the bias
// column sums over the bf16 d(pre-activations).
#include "tc05.cuh"

namespace slnlp {

// y (bf16) = dropout(x) with exactly the mask slnlp_dropout(site) draws (same Philox counter per group of 4)
__global__ void __launch_bounds__(256) dropout_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t n, float p,
                                                           const uint64_t* __restrict__ rng, uint32_t site) {
  pdl_wait();
  pdl_launch_dependents();
  const uint64_t seed = rng[0], step = rng[1];
  const float keep = 1.f - p, inv = 1.f / (1.f - p);
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < n; q += (int64_t)gridDim.x * blockDim.x) {
    uint32_t c[4] = {(uint32_t)q, (uint32_t)(q >> 32) ^ (site * 0x9E3779B9u), (uint32_t)step, (uint32_t)(step >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    if (q * 4 + 3 < n) {
      const float4 v = reinterpret_cast<const float4*>(x)[q];
      const float in[4] = {v.x, v.y, v.z, v.w};
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = ((float)(c[j] >> 8) * (1.0f / 16777216.0f)) < keep ? in[j] * inv : 0.f;
      uint2 pk;
      pk.x = pack2_bf16(o[0], o[1]);
      pk.y = pack2_bf16(o[2], o[3]);
      reinterpret_cast<uint2*>(y)[q] = pk;
    } else {
      for (int j = 0; j < 4; ++j) {
        const int64_t i = q * 4 + j;
        if (i < n) y[i] = __float2bfloat16_rn(((float)(c[j] >> 8) * (1.0f / 16777216.0f)) < keep ? x[i] * inv : 0.f);
      }
    }
  }
}

// out[c] (+)= sum_r A[r*lda + c], A bf16: 128 columns per CTA (a warp reads 256 contiguous bytes of a row), rows
// split over gridDim.y chunks that add with red.global.add when there are several
__global__ void __launch_bounds__(1024) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ A, int rows, int cols, int64_t lda,
                                                           float* __restrict__ out, float beta, int chunk) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float sm[32][129];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + lane * 4;
  const int r_end = min(rows, (int)(blockIdx.y + 1) * chunk);
  float s[4] = {0.f, 0.f, 0.f, 0.f}, u[4] = {0.f, 0.f, 0.f, 0.f};
  auto add4 = [](float (&acc)[4], uint2 v) {
    const __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&v.x), b = *reinterpret_cast<__nv_bfloat162*>(&v.y);
    acc[0] += __low2float(a); acc[1] += __high2float(a); acc[2] += __low2float(b); acc[3] += __high2float(b);
  };
  if (c < cols) {      // cols % 4 == 0: a group of 4 is inside or outside
    int r = blockIdx.y * chunk + w;
    for (; r + 96 < r_end; r += 128) {
      const uint2 v0 = *reinterpret_cast<const uint2*>(A + (int64_t)r * lda + c), v1 = *reinterpret_cast<const uint2*>(A + (int64_t)(r + 32) * lda + c);
      const uint2 v2 = *reinterpret_cast<const uint2*>(A + (int64_t)(r + 64) * lda + c), v3 = *reinterpret_cast<const uint2*>(A + (int64_t)(r + 96) * lda + c);
      add4(s, v0); add4(u, v1); add4(s, v2); add4(u, v3);
    }
    for (; r < r_end; r += 32) add4(s, *reinterpret_cast<const uint2*>(A + (int64_t)r * lda + c));
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) sm[w][lane * 4 + j] = s[j] + u[j];
  __syncthreads();
  if (threadIdx.x < 128) {
    const int cc = blockIdx.x * 128 + threadIdx.x;
    if (cc < cols) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) t += sm[i][threadIdx.x];
      if (gridDim.y > 1) atomicAdd(out + cc, t);
      else out[cc] = (beta == 0.f ? 0.f : beta * out[cc]) + t;
    }
  }
}

}  // namespace slnlp

using namespace slnlp;

extern "C" int slnlp_dropout_bf16(const float* x, uint16_t* y, int64_t n, float p, const uint64_t* rng, uint32_t site,
                                  slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(x && y && rng && n >= 0 && p >= 0.f && p < 1.f, "dropout_bf16: bad arguments");
  SLNLP_CHECK_ARG((((uintptr_t)x & 15) | ((uintptr_t)y & 7)) == 0, "dropout_bf16: x needs 16-byte, y 8-byte alignment");
  if (n == 0) return 0;
  const int64_t groups = (n + 3) / 4;
  const int64_t cap = (int64_t)(sm_count() > 0 ? sm_count() : 148) * 8;
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((groups + 255) / 256, cap));
  launch_pdl(dropout_bf16_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), x, reinterpret_cast<__nv_bfloat16*>(y), n, p, rng, site);
  SLNLP_LAUNCH_OK("dropout_bf16");
  return 0;
}

extern "C" int slnlp_colsum_bf16(const uint16_t* A, int rows, int cols, int64_t lda, float* out, float beta, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(A && out && rows >= 0 && cols > 0 && lda >= cols, "colsum_bf16: bad arguments");
  SLNLP_CHECK_ARG(cols % 4 == 0 && lda % 4 == 0 && ((uintptr_t)A & 7) == 0, "colsum_bf16: cols, lda multiples of 4, 8-byte aligned A");
  static int use_atomic = -1;
  if (use_atomic < 0) {
    const char* e = getenv("SLNLP_SPLITK_ATOMIC");
    use_atomic = (e && e[0] == '0') ? 0 : 1;
  }
  const int cb = ceil_div(cols, 128);
  int splits = 1;
  if (use_atomic && beta == 1.f && rows >= 512) {
    splits = ceil_div(2 * (sm_count() > 0 ? sm_count() : 148), cb);
    const int max_splits = rows / 128;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  const int chunk = ceil_div(rows > 0 ? rows : 1, splits);
  splits = ceil_div(rows > 0 ? rows : 1, chunk);
  launch_pdl(colsum_bf16_kernel, dim3(cb, splits), dim3(1024), 0, as_stream(stream), reinterpret_cast<const __nv_bfloat16*>(A), rows, cols,
             lda, out, beta, chunk);
  SLNLP_LAUNCH_OK("colsum_bf16");
  return 0;
}
