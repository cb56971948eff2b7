// Error plumbing + the HBM-bound kernels of the hot path:
//   K1 embedding gather/scatter, K10 log-softmax + CE-on-logp, K11/K12 grad-norm +
//   clip + SGD-momentum, dropout (own Philox), pad fill, small elementwise glue.
// Reference call sites are cited in include/slnlp_b200.h next to each entry point.
#include <stdlib.h>

#include "common.cuh"
#include <atomic>

namespace slnlp {
int64_t launches();

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return 1;
}
static std::atomic<int64_t> g_launches{0};
void note_launches(int64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int64_t launches() { return g_launches.load(std::memory_order_relaxed); }
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SLNLP_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
    cached = p.multiProcessorCount;
    cached_dev = dev;
  }
  return cached;
}

// ------------------------------------------------------------------ embedding
struct FieldSpec {
  int nf;
  int etot;
  int64_t off[SLNLP_MAX_FIELDS];
  int w[SLNLP_MAX_FIELDS];
  int col[SLNLP_MAX_FIELDS];
  int64_t rows[SLNLP_MAX_FIELDS];
};

// one warp per output row; float4 path when every field width is a multiple of 4
template <bool VEC4>
__global__ void __launch_bounds__(256) embed_fwd_kernel(const float* __restrict__ table,
                                                        const int64_t* __restrict__ idx,
                                                        float* __restrict__ out, int B, int T,
                                                        FieldSpec fs, int time_major, float scale,
                                                        const float* __restrict__ pe) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (int64_t)B * T) return;
  const int b = (int)(row / T), t = (int)(row % T);
  const int64_t orow = time_major ? (int64_t)t * B + b : row;
  float* o = out + orow * fs.etot;
  const float* per = pe ? pe + (int64_t)t * fs.etot : nullptr;
  for (int f = 0; f < fs.nf; ++f) {
    const int64_t id = idx[row * fs.nf + f];
    const bool ok = id >= 0 && id < fs.rows[f];
    const float* src = table + fs.off[f] + (ok ? id : 0) * fs.w[f];
    const int c0 = fs.col[f];
    if (VEC4) {
      for (int e = lane * 4; e < fs.w[f]; e += 128) {
        float4 v = ok ? __ldg(reinterpret_cast<const float4*>(src + e))
                      : make_float4(NAN, NAN, NAN, NAN);
        v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
        if (per) {
          const float4 p = __ldg(reinterpret_cast<const float4*>(per + c0 + e));
          v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
        }
        *reinterpret_cast<float4*>(o + c0 + e) = v;
      }
    } else {
      for (int e = lane; e < fs.w[f]; e += 32) {
        float v = ok ? __ldg(src + e) * scale : NAN;
        if (per) v += __ldg(per + c0 + e);
        o[c0 + e] = v;
      }
    }
  }
}

__global__ void __launch_bounds__(256) embed_bwd_kernel(float* __restrict__ dtable,
                                                        const int64_t* __restrict__ idx,
                                                        const float* __restrict__ dout, int B, int T,
                                                        FieldSpec fs, int time_major, float scale,
                                                        int64_t padding_idx) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (int64_t)B * T) return;
  const int b = (int)(row / T), t = (int)(row % T);
  const int64_t orow = time_major ? (int64_t)t * B + b : row;
  const float* g = dout + orow * fs.etot;
  for (int f = 0; f < fs.nf; ++f) {
    const int64_t id = idx[row * fs.nf + f];
    if (id == padding_idx || id < 0 || id >= fs.rows[f]) continue;
    float* dst = dtable + fs.off[f] + id * fs.w[f];
    for (int e = lane; e < fs.w[f]; e += 32) atomicAdd(dst + e, scale * g[fs.col[f] + e]);
  }
}

static int make_fields(FieldSpec& fs, int F, const int64_t* off, const int* w, const int64_t* rows) {
  SLNLP_CHECK_ARG(F >= 1 && F <= SLNLP_MAX_FIELDS, "embed: F=%d out of range", F);
  fs.nf = F;
  fs.etot = 0;
  for (int f = 0; f < F; ++f) {
    SLNLP_CHECK_ARG(w[f] > 0 && rows[f] > 0, "embed: bad field %d", f);
    fs.off[f] = off[f];
    fs.w[f] = w[f];
    fs.rows[f] = rows[f];
    fs.col[f] = fs.etot;
    fs.etot += w[f];
  }
  return 0;
}

// ------------------------------------------------------------------ log-softmax / CE
__global__ void __launch_bounds__(256) log_softmax_fwd_kernel(const float* __restrict__ x,
                                                              float* __restrict__ y, int V) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float red[33];
  const float* xr = x + (int64_t)blockIdx.x * V;
  float* yr = y + (int64_t)blockIdx.x * V;
  float m = -INFINITY;
  for (int v = threadIdx.x; v < V; v += blockDim.x) m = fmaxf(m, xr[v]);
  m = block_max(m, red);
  float s = 0.f;
  for (int v = threadIdx.x; v < V; v += blockDim.x) s += expf(xr[v] - m);
  s = block_sum(s, red);
  const float lse = m + logf(s);
  for (int v = threadIdx.x; v < V; v += blockDim.x) yr[v] = xr[v] - lse;
}

__global__ void __launch_bounds__(256) log_softmax_bwd_kernel(const float* __restrict__ dy,
                                                              const float* __restrict__ y,
                                                              float* __restrict__ dx, int V, int ld) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float red[33];
  const int64_t o = (int64_t)blockIdx.x * V, od = (int64_t)blockIdx.x * ld;
  float s = 0.f;
  for (int v = threadIdx.x; v < V; v += blockDim.x) s += dy[o + v];
  s = block_sum(s, red);
  for (int v = threadIdx.x; v < V; v += blockDim.x) dx[od + v] = dy[o + v] - expf(y[o + v]) * s;
}

// row_ws: [0,B) row loss, [B,2B) valid flag, [2B,3B) second logsumexp
__global__ void __launch_bounds__(256) ce_rows_kernel(const float* __restrict__ logp,
                                                      const int64_t* __restrict__ y,
                                                      int64_t ignore, int B, int V,
                                                      float* __restrict__ row_ws) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float red[33];
  const int b = blockIdx.x;
  const float* r = logp + (int64_t)b * V;
  float m = -INFINITY;
  for (int v = threadIdx.x; v < V; v += blockDim.x) m = fmaxf(m, r[v]);
  m = block_max(m, red);
  float s = 0.f;
  for (int v = threadIdx.x; v < V; v += blockDim.x) s += expf(r[v] - m);
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    const float lse = m + logf(s);
    const int64_t yb = y[b];
    const bool valid = yb != ignore && yb >= 0 && yb < V;
    row_ws[b] = valid ? -(r[yb] - lse) : 0.f;
    row_ws[B + b] = valid ? 1.f : 0.f;
    row_ws[2 * B + b] = lse;
  }
}
// K10 fused: generator log_softmax (bkp:75-76) + CrossEntropyLoss(ignore_index) on the log-probs
// (config/*.yaml:36) + d loss / d logits in ONE pass per row.  The row stays in registers/L1:
// logits -> logp (written: the module's output) -> second log-sum-exp -> row loss, and the
// gradient (softmax(logp) - onehot) * valid / n_valid, where n_valid is recounted by every CTA
// from the B labels (cheaper than a grid-wide dependency).  The mean loss itself is off the
// critical path and is reduced by ce_reduce_kernel afterwards.
__global__ void __launch_bounds__(256) logsoftmax_ce_fused_kernel(const float* __restrict__ logits,
                                                                  const int64_t* __restrict__ y, int64_t ignore,
                                                                  int B, int V, float* __restrict__ logp,
                                                                  float* __restrict__ row_ws,
                                                                  float* __restrict__ dlogits, int ld) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float red[33];
  const int b = blockIdx.x;
  const float* xr = logits + (int64_t)b * V;
  float* lr = logp + (int64_t)b * V;
  float m = -INFINITY;
  for (int v = threadIdx.x; v < V; v += blockDim.x) m = fmaxf(m, xr[v]);
  m = block_max(m, red);
  float s = 0.f;
  for (int v = threadIdx.x; v < V; v += blockDim.x) s += expf(xr[v] - m);
  s = block_sum(s, red);
  const float lse1 = m + logf(s);
  // second log-softmax, on the log-probs (what CrossEntropyLoss applies to the module output)
  const float m2 = m - lse1;                      // max of logp
  float s2 = 0.f;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    const float lp = xr[v] - lse1;
    lr[v] = lp;
    s2 += expf(lp - m2);
  }
  s2 = block_sum(s2, red);
  const float lse2 = m2 + logf(s2);
  float cnt = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const int64_t yi = y[i];
    cnt += (yi != ignore && yi >= 0 && yi < V) ? 1.f : 0.f;
  }
  cnt = block_sum(cnt, red);
  const int64_t yb = y[b];
  const bool valid = yb != ignore && yb >= 0 && yb < V;
  if (threadIdx.x == 0) {
    row_ws[b] = valid ? -((xr[yb] - lse1) - lse2) : 0.f;
    row_ws[B + b] = valid ? 1.f : 0.f;
    row_ws[2 * B + b] = lse2;
  }
  if (dlogits) {
    const float inv = valid ? 1.f / cnt : 0.f;
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
      const float g = expf((xr[v] - lse1) - lse2) - (v == yb ? 1.f : 0.f);
      dlogits[(int64_t)b * ld + v] = valid ? g * inv : 0.f;
    }
  }
}
__global__ void __launch_bounds__(256) ce_reduce_kernel(const float* __restrict__ row_ws, int B,
                                                        float* __restrict__ loss_out) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float red[33];
  float s = 0.f, c = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    s += row_ws[b];
    c += row_ws[B + b];
  }
  s = block_sum(s, red);
  c = block_sum(c, red);
  if (threadIdx.x == 0) {
    loss_out[0] = s / c;  // 0/0 = NaN when every target is ignored, as torch
    loss_out[1] = c;
  }
}
__global__ void __launch_bounds__(256) ce_grad_kernel(const float* __restrict__ logp,
                                                      const int64_t* __restrict__ y, int B, int V,
                                                      const float* __restrict__ row_ws,
                                                      const float* __restrict__ loss_out,
                                                      float* __restrict__ dlogits, int ld) {
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.x;
  const float valid = row_ws[B + b], lse = row_ws[2 * B + b];
  const float inv = valid / loss_out[1];
  const int64_t yb = y[b];
  const int64_t o = (int64_t)b * V;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    // d/dlogp = softmax(logp) - onehot; through log_softmax(logits) the rowsum term is
    // (1 - 1) = 0, so the same expression is d/dlogits.
    float g = expf(logp[o + v] - lse) - (v == yb ? 1.f : 0.f);
    dlogits[(int64_t)b * ld + v] = valid != 0.f ? g * inv : 0.f;
  }
}

// ------------------------------------------------------------------ grad norm + clip + SGD
constexpr int kSumsqBlocks = 1024;
// One launch: every block leaves its partial sum, the LAST block to finish (ticket counter behind the
// partials, reset for the next call) adds them in double and writes the norm.
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, int64_t n,
                                                            float* __restrict__ partials, unsigned* __restrict__ ticket,
                                                            float* __restrict__ norm_out) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float red[33];
  __shared__ double redd[256];
  __shared__ bool last;
  float s = 0.f;
  const int64_t n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(g4 + i);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0)
    for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) s += g[i] * g[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = s;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  double t = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += (double)__ldcg(partials + i);
  redd[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) redd[threadIdx.x] += redd[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    norm_out[0] = (float)sqrt(redd[0]);
    *ticket = 0u;
  }
}
__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                  float* __restrict__ buf, int64_t n,
                                                  const float* __restrict__ hyper,
                                                  const float* __restrict__ norm, float grad_scale, float* zero_g) {
  pdl_wait();
  pdl_launch_dependents();
  const float lr = hyper[0], mom = hyper[1], max_norm = hyper[2];
  const bool first = hyper[3] != 0.f;
  float coef = grad_scale;
  if (max_norm > 0.f && norm) coef *= fminf(1.f, max_norm / (norm[0] * grad_scale + 1e-6f));
  const int64_t n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* b4 = reinterpret_cast<float4*>(buf);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float4 gv = __ldg(g4 + i);
    float4 bv = b4[i], pv = p4[i];
    bv.x = first ? gv.x * coef : mom * bv.x + gv.x * coef;
    bv.y = first ? gv.y * coef : mom * bv.y + gv.y * coef;
    bv.z = first ? gv.z * coef : mom * bv.z + gv.z * coef;
    bv.w = first ? gv.w * coef : mom * bv.w + gv.w * coef;
    pv.x -= lr * bv.x; pv.y -= lr * bv.y; pv.z -= lr * bv.z; pv.w -= lr * bv.w;
    b4[i] = bv;
    p4[i] = pv;
    if (zero_g) reinterpret_cast<float4*>(zero_g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);   // consumed: ready for the next step
  }
  if (blockIdx.x == 0)
    for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      const float gv = g[i] * coef;
      const float bv = first ? gv : mom * buf[i] + gv;
      buf[i] = bv;
      p[i] -= lr * bv;
      if (zero_g) zero_g[i] = 0.f;
    }
}

// ------------------------------------------------------------------ dropout (Philox4x32-10)
__global__ void __launch_bounds__(256) dropout_kernel(const float* __restrict__ x,
                                                      float* __restrict__ y, int64_t n, float p,
                                                      const uint64_t* __restrict__ rng, uint32_t site) {
  pdl_wait();
  pdl_launch_dependents();
  const uint64_t seed = rng[0], step = rng[1];
  const float keep = 1.f - p, inv = 1.f / (1.f - p);
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < n;
       q += (int64_t)gridDim.x * blockDim.x) {
    uint32_t c[4] = {(uint32_t)q, (uint32_t)(q >> 32) ^ (site * 0x9E3779B9u), (uint32_t)step,
                     (uint32_t)(step >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t i = q * 4 + j;
      if (i < n) {
        const float u = (float)(c[j] >> 8) * (1.0f / 16777216.0f);
        y[i] = u < keep ? x[i] * inv : 0.f;
      }
    }
  }
}
// the keep / scale factors slnlp_dropout(site) applies, as a tensor: generated once per step off the critical
// path, consumed by kernels that cannot afford ten Philox rounds in their loop (the persistent recurrent kernels)
__global__ void __launch_bounds__(256) dropout_mask_kernel(float* __restrict__ y, int64_t n, float p,
                                                           const uint64_t* __restrict__ rng, uint32_t site) {
  pdl_wait();
  pdl_launch_dependents();
  const uint64_t seed = rng[0], step = rng[1];
  const float keep = 1.f - p, inv = 1.f / (1.f - p);
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < n;
       q += (int64_t)gridDim.x * blockDim.x) {
    uint32_t c[4] = {(uint32_t)q, (uint32_t)(q >> 32) ^ (site * 0x9E3779B9u), (uint32_t)step,
                     (uint32_t)(step >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t i = q * 4 + j;
      if (i < n) y[i] = (float)(c[j] >> 8) * (1.0f / 16777216.0f) < keep ? inv : 0.f;
    }
  }
}
__global__ void rng_advance_kernel(uint64_t* rng) {
  pdl_wait();
  pdl_launch_dependents(); rng[1] += 1; }

// ------------------------------------------------------------------ small glue
__global__ void __launch_bounds__(256) pad_fill_kernel(float* __restrict__ x,
                                                       const int64_t* __restrict__ lengths, int T,
                                                       int B, int W, float value) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x;  // t*B + b
  const int t = row / B, b = row % B;
  if (t < lengths[b]) return;
  float* r = x + (int64_t)row * W;
  for (int e = threadIdx.x; e < W; e += blockDim.x) r[e] = value;
}
// out-of-place form: dst = src with the padded rows replaced by `value` (the BPTT keeps reading the
// zero-padded original, so nothing has to be un-filled later)
__global__ void __launch_bounds__(256) pad_fill_copy_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                            const int64_t* __restrict__ lengths, int T, int B, int W,
                                                            float value) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x;  // t*B + b
  const int t = row / B, b = row % B;
  const bool pad = t >= lengths[b];
  const float* s = src + (int64_t)row * W;
  float* r = dst + (int64_t)row * W;
  if ((W & 3) == 0 && ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0)) {
    const float4 v4 = make_float4(value, value, value, value);
    for (int e = threadIdx.x; e < W / 4; e += blockDim.x)
      reinterpret_cast<float4*>(r)[e] = pad ? v4 : reinterpret_cast<const float4*>(s)[e];
  } else {
    for (int e = threadIdx.x; e < W; e += blockDim.x) r[e] = pad ? value : s[e];
  }
}
__global__ void __launch_bounds__(256) concat_dirs_kernel(const float* __restrict__ src,
                                                          float* __restrict__ dst, int B, int H,
                                                          int ndir, int inverse) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t n = (int64_t)ndir * B * H;
  if (n < (1ll << 31) && (H & 3) == 0 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
    // 32-bit index arithmetic, float4 rows (64-bit div / mod per element made this 260 us at B 4096 x H 512)
    const uint32_t h4 = (uint32_t)H >> 2, n4 = (uint32_t)(n >> 2);
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
      const uint32_t row = i / h4, k = i - row * h4;      // row = d * B + b
      const uint32_t d = row / (uint32_t)B, b = row - d * (uint32_t)B;
      const uint32_t j = (b * (uint32_t)ndir + d) * h4 + k;
      if (inverse) d4[i] = s4[j]; else d4[j] = s4[i];
    }
    return;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % H);
    const int b = (int)((i / H) % B);
    const int d = (int)(i / ((int64_t)H * B));
    const int64_t j = ((int64_t)b * ndir + d) * H + k;  // [B, ndir*H]
    if (inverse) dst[i] = src[j]; else dst[j] = src[i];
  }
}
__global__ void __launch_bounds__(256) tanh_fwd_kernel(float* x, int64_t n) {
  pdl_wait();
  pdl_launch_dependents();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    x[i] = tanhf(x[i]);
}
__global__ void __launch_bounds__(256) tanh_bwd_kernel(float* dy, const float* __restrict__ y, int64_t n) {
  pdl_wait();
  pdl_launch_dependents();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    dy[i] *= 1.f - y[i] * y[i];
}
__global__ void __launch_bounds__(256) relu_fwd_kernel(float* x, int64_t n) {
  pdl_wait();
  pdl_launch_dependents();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    x[i] = fmaxf(x[i], 0.f);
}
__global__ void __launch_bounds__(256) relu_bwd_kernel(float* dy, const float* __restrict__ y, int64_t n) {
  pdl_wait();
  pdl_launch_dependents();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    dy[i] = y[i] > 0.f ? dy[i] : 0.f;
}
// y = dropout(relu(x), p) in place with the mask slnlp_dropout(site) draws (Philox block = 4 consecutive elements): the
// FFN's activation + dropout as ONE pass (nn.TransformerEncoderLayer: linear2(dropout(relu(linear1(x)))))
__global__ void __launch_bounds__(256) relu_dropout_fwd_kernel(float* __restrict__ x, int64_t n, float p,
                                                               const uint64_t* __restrict__ rng, uint32_t site) {
  pdl_wait();
  pdl_launch_dependents();
  const uint64_t seed = rng[0], step = rng[1];
  const float keep = 1.f - p, inv = 1.f / (1.f - p);
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < n; q += (int64_t)gridDim.x * blockDim.x) {
    float u[4];
    philox_uniform4(seed, step, site, (uint64_t)q, u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t i = q * 4 + j;
      if (i < n) x[i] = u[j] < keep ? fmaxf(x[i], 0.f) * inv : 0.f;
    }
  }
}
// d x of the same: y > 0 exactly where the unit was active AND kept, so the backward needs no random numbers
__global__ void __launch_bounds__(256) relu_scaled_bwd_kernel(float* dy, const float* __restrict__ y, float scale, int64_t n) {
  pdl_wait();
  pdl_launch_dependents();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    dy[i] = y[i] > 0.f ? dy[i] * scale : 0.f;
}
__global__ void __launch_bounds__(256) axpy_kernel(float* y, const float* __restrict__ x, float a, int64_t n) {
  pdl_wait();
  pdl_launch_dependents();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    y[i] += a * x[i];
}
__global__ void __launch_bounds__(256) dec_input_fwd_kernel(const float* __restrict__ row,
                                                            const float* __restrict__ src,
                                                            float* __restrict__ dst, int B, int E, int W) {
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.x;
  for (int e = threadIdx.x; e < E + W; e += blockDim.x)
    dst[(int64_t)b * (E + W) + e] = e < E ? row[e] : src[(int64_t)b * W + e - E];
}
// grid = ceil((E+W)/256) blocks; column sums over b for the first E columns (deterministic)
__global__ void __launch_bounds__(256) dec_input_bwd_kernel(const float* __restrict__ ddst,
                                                            float* __restrict__ drow,
                                                            float* __restrict__ dsrc, int B, int E, int W) {
  pdl_wait();
  pdl_launch_dependents();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E + W) return;
  if (e < E) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += ddst[(int64_t)b * (E + W) + e];
    drow[e] += s;
  } else {
    for (int b = 0; b < B; ++b) dsrc[(int64_t)b * W + e - E] = ddst[(int64_t)b * (E + W) + e];
  }
}
// out[c] = beta*out[c] + sum_r A[r,c]; 32 columns per CTA, 32 warps stride the rows of the CTA's
// row chunk, fixed combination order inside a CTA.  gridDim.y > 1 (beta == 1 only): every row chunk
// adds its partial sum with red.global.add, so that a tall matrix is read by the whole GPU instead
// of cols/32 SMs.
__global__ void __launch_bounds__(1024) colsum_kernel(const float* __restrict__ A, int rows, int cols,
                                                      int lda, float* __restrict__ out, float beta, int chunk) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float sm[32][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int r_end = min(rows, (int)(blockIdx.y + 1) * chunk);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (c < cols) {
    int r = blockIdx.y * chunk + w;
    for (; r + 96 < r_end; r += 128) {
      s0 += A[(int64_t)r * lda + c];
      s1 += A[(int64_t)(r + 32) * lda + c];
      s2 += A[(int64_t)(r + 64) * lda + c];
      s3 += A[(int64_t)(r + 96) * lda + c];
    }
    for (; r < r_end; r += 32) s0 += A[(int64_t)r * lda + c];
  }
  sm[w][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (w == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += sm[i][lane];
    if (gridDim.y > 1) atomicAdd(out + c, t);
    else out[c] = (beta == 0.f ? 0.f : beta * out[c]) + t;
  }
}

static inline int ew_grid(int64_t n) {
  int64_t g = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace slnlp

using namespace slnlp;

extern "C" {

int slnlp_abi_version(void) { return SLNLP_ABI_VERSION; }
const char* slnlp_last_error_string(void) { return err_buf(); }
int64_t slnlp_launch_count(void) { return slnlp::launches(); }
void* slnlp_stream_create(void) {
  cudaStream_t st = nullptr;
  if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) {
    slnlp::fail("slnlp_stream_create: %s", cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  return st;
}
void* slnlp_stream_create_priority(int high) {
  int lo = 0, hi = 0;      // numerically lower = higher priority
  cudaStream_t st = nullptr;
  if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess ||
      cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, high ? hi : lo) != cudaSuccess) {
    slnlp::fail("slnlp_stream_create_priority: %s", cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  return st;
}
int slnlp_stream_destroy(void* stream) {
  const cudaError_t e = cudaStreamDestroy(reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return slnlp::fail("slnlp_stream_destroy: %s", cudaGetErrorString(e));
  return 0;
}
int slnlp_device_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return n;
}

int slnlp_embed_gather_fwd(const float* table, const int64_t* idx, float* out, int B, int T, int F,
                           const int64_t* field_off, const int* field_w, const int64_t* field_rows,
                           int time_major, float scale, const float* pe, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(table && idx && out && B > 0 && T > 0, "embed_gather_fwd: bad arguments");
  FieldSpec fs;
  if (make_fields(fs, F, field_off, field_w, field_rows)) return 1;
  bool vec = (fs.etot % 4 == 0) && ((uintptr_t)table % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
             (!pe || (uintptr_t)pe % 16 == 0);
  for (int f = 0; f < F; ++f) vec = vec && fs.w[f] % 4 == 0 && fs.off[f] % 4 == 0;
  const int grid = ceil_div((int64_t)B * T, 8);
  if (vec)
    launch_pdl(embed_fwd_kernel<true>, dim3(grid), dim3(256), 0, as_stream(stream), table, idx, out, B, T, fs, time_major, scale, pe);
  else
    launch_pdl(embed_fwd_kernel<false>, dim3(grid), dim3(256), 0, as_stream(stream), table, idx, out, B, T, fs, time_major, scale, pe);
  SLNLP_LAUNCH_OK("embed_gather_fwd");
  return 0;
}

int slnlp_embed_gather_bwd(float* dtable, const int64_t* idx, const float* dout, int B, int T, int F,
                           const int64_t* field_off, const int* field_w, const int64_t* field_rows,
                           int time_major, float scale, int64_t padding_idx, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(dtable && idx && dout && B > 0 && T > 0, "embed_gather_bwd: bad arguments");
  FieldSpec fs;
  if (make_fields(fs, F, field_off, field_w, field_rows)) return 1;
  launch_pdl(embed_bwd_kernel, dim3(ceil_div((int64_t)B * T, 8)), dim3(256), 0, as_stream(stream), 
      dtable, idx, dout, B, T, fs, time_major, scale, padding_idx);
  SLNLP_LAUNCH_OK("embed_gather_bwd");
  return 0;
}

int slnlp_colsum_f32(const float* A, int rows, int cols, int lda, float* out, float beta,
                     slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(A && out && rows >= 0 && cols > 0 && lda >= cols, "colsum: bad arguments");
  // tall gradient reductions (out += colsum): split the rows over ~2 CTAs per SM.  The fp32 order of
  // the chunk sums then varies (~1e-7 relative); SLNLP_SPLITK_ATOMIC=0 keeps the single-pass form.
  static int use_atomic = -1;
  if (use_atomic < 0) {
    const char* e = getenv("SLNLP_SPLITK_ATOMIC");
    use_atomic = (e && e[0] == '0') ? 0 : 1;
  }
  const int cb = ceil_div(cols, 32);
  int splits = 1;
  if (use_atomic && beta == 1.f && rows >= 512) {
    splits = ceil_div(2 * sm_count(), cb);
    const int max_splits = rows / 128;   // at least 4 rows per warp
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  const int chunk = ceil_div(rows > 0 ? rows : 1, splits);
  splits = ceil_div(rows > 0 ? rows : 1, chunk);
  launch_pdl(colsum_kernel, dim3(cb, splits), dim3(1024), 0, as_stream(stream), A, rows, cols, lda, out, beta, chunk);
  SLNLP_LAUNCH_OK("colsum");
  return 0;
}

int slnlp_pad_fill(float* x, const int64_t* lengths, int T, int B, int W, float value,
                   slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(x && lengths && T > 0 && B > 0 && W > 0, "pad_fill: bad arguments");
  launch_pdl(pad_fill_kernel, dim3(T * B), dim3(128), 0, as_stream(stream), x, lengths, T, B, W, value);
  SLNLP_LAUNCH_OK("pad_fill");
  return 0;
}

int slnlp_pad_fill_copy(const float* src, float* dst, const int64_t* lengths, int T, int B, int W, float value,
                        slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(src && dst && src != dst && lengths && T > 0 && B > 0 && W > 0, "pad_fill_copy: bad arguments");
  launch_pdl(pad_fill_copy_kernel, dim3(T * B), dim3(64), 0, as_stream(stream), src, dst, lengths, T, B, W, value);
  SLNLP_LAUNCH_OK("pad_fill_copy");
  return 0;
}

int slnlp_concat_dirs(const float* src, float* dst, int B, int H, int ndir, int inverse,
                      slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(src && dst && B > 0 && H > 0 && ndir > 0, "concat_dirs: bad arguments");
  launch_pdl(concat_dirs_kernel, dim3(ew_grid((int64_t)ndir * B * H)), dim3(256), 0, as_stream(stream), src, dst, B, H, ndir, inverse);
  SLNLP_LAUNCH_OK("concat_dirs");
  return 0;
}

int slnlp_tanh_fwd(float* x, int64_t n, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(x && n >= 0, "tanh_fwd: bad arguments");
  if (n == 0) return 0;
  launch_pdl(tanh_fwd_kernel, dim3(ew_grid(n)), dim3(256), 0, as_stream(stream), x, n);
  SLNLP_LAUNCH_OK("tanh_fwd");
  return 0;
}
int slnlp_tanh_bwd(float* dy, const float* y, int64_t n, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(dy && y && n >= 0, "tanh_bwd: bad arguments");
  if (n == 0) return 0;
  launch_pdl(tanh_bwd_kernel, dim3(ew_grid(n)), dim3(256), 0, as_stream(stream), dy, y, n);
  SLNLP_LAUNCH_OK("tanh_bwd");
  return 0;
}
int slnlp_relu_fwd(float* x, int64_t n, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(x && n >= 0, "relu_fwd: bad arguments");
  if (n == 0) return 0;
  launch_pdl(relu_fwd_kernel, dim3(ew_grid(n)), dim3(256), 0, as_stream(stream), x, n);
  SLNLP_LAUNCH_OK("relu_fwd");
  return 0;
}
int slnlp_relu_bwd(float* dy, const float* y, int64_t n, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(dy && y && n >= 0, "relu_bwd: bad arguments");
  if (n == 0) return 0;
  launch_pdl(relu_bwd_kernel, dim3(ew_grid(n)), dim3(256), 0, as_stream(stream), dy, y, n);
  SLNLP_LAUNCH_OK("relu_bwd");
  return 0;
}
int slnlp_relu_dropout_fwd(float* x, int64_t n, float p, const uint64_t* rng, uint32_t site, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(x && rng && n >= 0 && p >= 0.f && p < 1.f, "relu_dropout_fwd: bad arguments");
  if (n == 0) return 0;
  launch_pdl(relu_dropout_fwd_kernel, dim3(ew_grid((n + 3) / 4)), dim3(256), 0, as_stream(stream), x, n, p, rng, site);
  SLNLP_LAUNCH_OK("relu_dropout_fwd");
  return 0;
}
int slnlp_relu_dropout_bwd(float* dy, const float* y, int64_t n, float p, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(dy && y && n >= 0 && p >= 0.f && p < 1.f, "relu_dropout_bwd: bad arguments");
  if (n == 0) return 0;
  launch_pdl(relu_scaled_bwd_kernel, dim3(ew_grid(n)), dim3(256), 0, as_stream(stream), dy, y, 1.f / (1.f - p), n);
  SLNLP_LAUNCH_OK("relu_dropout_bwd");
  return 0;
}
int slnlp_axpy(float* y, const float* x, float a, int64_t n, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(y && x && n >= 0, "axpy: bad arguments");
  if (n == 0) return 0;
  launch_pdl(axpy_kernel, dim3(ew_grid(n)), dim3(256), 0, as_stream(stream), y, x, a, n);
  SLNLP_LAUNCH_OK("axpy");
  return 0;
}
int slnlp_dropout(const float* x, float* y, int64_t n, float p, const uint64_t* rng, uint32_t site,
                  slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(x && y && rng && n >= 0 && p >= 0.f && p < 1.f, "dropout: bad arguments");
  if (n == 0) return 0;
  launch_pdl(dropout_kernel, dim3(ew_grid((n + 3) / 4)), dim3(256), 0, as_stream(stream), x, y, n, p, rng, site);
  SLNLP_LAUNCH_OK("dropout");
  return 0;
}
int slnlp_dropout_mask(float* y, int64_t n, float p, const uint64_t* rng, uint32_t site, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(y && rng && n >= 0 && p >= 0.f && p < 1.f, "dropout_mask: bad arguments");
  if (n == 0) return 0;
  launch_pdl(dropout_mask_kernel, dim3(ew_grid((n + 3) / 4)), dim3(256), 0, as_stream(stream), y, n, p, rng, site);
  SLNLP_LAUNCH_OK("dropout_mask");
  return 0;
}
int slnlp_rng_advance(uint64_t* rng, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(rng, "rng_advance: null");
  launch_pdl(rng_advance_kernel, dim3(1), dim3(1), 0, as_stream(stream), rng);
  SLNLP_LAUNCH_OK("rng_advance");
  return 0;
}
int slnlp_dec_input_fwd(const float* row, const float* src, float* dst, int B, int E, int W,
                        slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(row && src && dst && B > 0 && E > 0 && W > 0, "dec_input_fwd: bad arguments");
  launch_pdl(dec_input_fwd_kernel, dim3(B), dim3(256), 0, as_stream(stream), row, src, dst, B, E, W);
  SLNLP_LAUNCH_OK("dec_input_fwd");
  return 0;
}
int slnlp_dec_input_bwd(const float* ddst, float* drow, float* dsrc, int B, int E, int W,
                        slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(ddst && drow && dsrc && B > 0 && E > 0 && W > 0, "dec_input_bwd: bad arguments");
  launch_pdl(dec_input_bwd_kernel, dim3(ceil_div(E + W, 256)), dim3(256), 0, as_stream(stream), ddst, drow, dsrc, B, E, W);
  SLNLP_LAUNCH_OK("dec_input_bwd");
  return 0;
}

int slnlp_log_softmax_fwd(const float* logits, float* logp, int B, int V, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(logits && logp && B > 0 && V > 0, "log_softmax_fwd: bad arguments");
  launch_pdl(log_softmax_fwd_kernel, dim3(B), dim3(256), 0, as_stream(stream), logits, logp, V);
  SLNLP_LAUNCH_OK("log_softmax_fwd");
  return 0;
}
int slnlp_log_softmax_bwd(const float* dlogp, const float* logp, float* dlogits, int B, int V,
                          int ld_dlogits, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(ld_dlogits >= V, "log_softmax_bwd: ld_dlogits < V");
  SLNLP_CHECK_ARG(dlogp && logp && dlogits && B > 0 && V > 0, "log_softmax_bwd: bad arguments");
  launch_pdl(log_softmax_bwd_kernel, dim3(B), dim3(256), 0, as_stream(stream), dlogp, logp, dlogits, V, ld_dlogits);
  SLNLP_LAUNCH_OK("log_softmax_bwd");
  return 0;
}
int slnlp_ce_on_logp(const float* logp, const int64_t* y, int64_t ignore_index, int B, int V,
                     float* loss_out, float* dlogits, int ld_dlogits, float* row_ws, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(!dlogits || ld_dlogits >= V, "ce_on_logp: ld_dlogits < V");
  SLNLP_CHECK_ARG(logp && y && loss_out && row_ws && B > 0 && V > 0, "ce_on_logp: bad arguments");
  launch_pdl(ce_rows_kernel, dim3(B), dim3(256), 0, as_stream(stream), logp, y, ignore_index, B, V, row_ws);
  launch_pdl(ce_reduce_kernel, dim3(1), dim3(256), 0, as_stream(stream), row_ws, B, loss_out);
  if (dlogits) launch_pdl(ce_grad_kernel, dim3(B), dim3(256), 0, as_stream(stream), logp, y, B, V, row_ws, loss_out, dlogits, ld_dlogits);
  note_launches(dlogits ? 2 : 1);
  SLNLP_LAUNCH_OK("ce_on_logp");
  return 0;
}

int slnlp_logsoftmax_ce_fused(const float* logits, const int64_t* y, int64_t ignore_index, int B, int V,
                              float* logp, float* loss_out, float* dlogits, int ld_dlogits, float* row_ws,
                              slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(logits && y && logp && row_ws && B > 0 && V > 0, "logsoftmax_ce_fused: bad arguments");
  SLNLP_CHECK_ARG(!dlogits || ld_dlogits >= V, "logsoftmax_ce_fused: ld_dlogits < V");
  launch_pdl(logsoftmax_ce_fused_kernel, dim3(B), dim3(256), 0, as_stream(stream), logits, y, ignore_index, B, V, logp,
             row_ws, dlogits, ld_dlogits);
  if (loss_out) {     // NULL: the caller reduces the row losses itself (slnlp_ce_reduce), off the backward chain
    launch_pdl(ce_reduce_kernel, dim3(1), dim3(256), 0, as_stream(stream), row_ws, B, loss_out);
    note_launches(1);
  }
  SLNLP_LAUNCH_OK("logsoftmax_ce_fused");
  return 0;
}
int slnlp_ce_reduce(const float* row_ws, int B, float* loss_out, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(row_ws && loss_out && B > 0, "ce_reduce: bad arguments");
  launch_pdl(ce_reduce_kernel, dim3(1), dim3(256), 0, as_stream(stream), row_ws, B, loss_out);
  SLNLP_LAUNCH_OK("ce_reduce");
  return 0;
}

int slnlp_sumsq_partials(void) { return kSumsqBlocks + 4; }   // + the ticket counter (must start at 0)
int slnlp_gradnorm(const float* g, int64_t n, float* partials, float* norm_out, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(g && partials && norm_out && n > 0, "gradnorm: bad arguments");
  SLNLP_CHECK_ARG((uintptr_t)g % 16 == 0, "gradnorm: g must be 16-byte aligned");
  int64_t want = (n / 4 + 255) / 256;
  int grid = (int)(want < 1 ? 1 : (want > kSumsqBlocks ? kSumsqBlocks : want));
  launch_pdl(sumsq_partial_kernel, dim3(grid), dim3(256), 0, as_stream(stream), g, n, partials,
             reinterpret_cast<unsigned*>(partials + kSumsqBlocks), norm_out);
  SLNLP_LAUNCH_OK("gradnorm");
  return 0;
}
static int sgd_launch(float* p, float* g, float* buf, int64_t n, const float* hyper, const float* norm,
                      float grad_scale, bool zero_grad, slnlp_stream_t stream) {
  SLNLP_CHECK_ARG(p && g && buf && hyper && n > 0, "sgd_momentum_clip: bad arguments");
  SLNLP_CHECK_ARG(((uintptr_t)p | (uintptr_t)g | (uintptr_t)buf) % 16 == 0,
                  "sgd_momentum_clip: buffers must be 16-byte aligned");
  int64_t want = (n / 4 + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
  launch_pdl(sgd_kernel, dim3(grid), dim3(256), 0, as_stream(stream), p, (const float*)g, buf, n, hyper, norm, grad_scale,
             zero_grad ? g : (float*)nullptr);
  SLNLP_LAUNCH_OK("sgd_momentum_clip");
  return 0;
}
int slnlp_sgd_momentum_clip(float* p, const float* g, float* buf, int64_t n, const float* hyper,
                            const float* norm, float grad_scale, slnlp_stream_t stream) {
  return sgd_launch(p, const_cast<float*>(g), buf, n, hyper, norm, grad_scale, false, stream);
}
int slnlp_sgd_momentum_clip_zero(float* p, float* g, float* buf, int64_t n, const float* hyper,
                                 const float* norm, float grad_scale, slnlp_stream_t stream) {
  return sgd_launch(p, g, buf, n, hyper, norm, grad_scale, true, stream);
}

}  // extern "C"
