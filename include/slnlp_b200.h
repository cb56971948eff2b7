/* slnlp_b200.h - C ABI of the B200-native hot path of sign-language-nlp.
 *
 * The reference (amorim-cleison/sign-language-nlp) has no FFI of its own: its hot
 * path is a stack of torch.nn modules called by skorch.  This
 * header is the boundary a maintainer binds instead: every entry point names the
 * reference call site it replaces ("bkp" = model/base/encoder_decoder_attn_bkp.py).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into caller-owned memory unless it says
 *     "host"; the library never allocates or frees device memory;
 *   - `stream` is a cudaStream_t passed as void*; all calls are asynchronous on it
 *     and capturable into a CUDA graph (no synchronisation, no allocation);
 *   - return value 0 = ok, non-zero = error; slnlp_last_error_string() describes the
 *     last error of the calling thread.  There is no CPU fallback anywhere;
 *   - activations are fp32, row-major; "time-major" means row index t*B + b;
 *   - RNN gate order is torch's: LSTM i,f,g,o (G=4); GRU r,z,n (G=3);
 *   - per-layer RNN weights of the two directions are adjacent in memory:
 *     w_ih [ndir*G*H, D], w_hh [ndir, G*H, H], b_ih / b_hh [ndir*G*H].
 */
#ifndef SLNLP_B200_H
#define SLNLP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLNLP_ABI_VERSION 1
#define SLNLP_MODE_LSTM 0
#define SLNLP_MODE_GRU 1
#define SLNLP_MAX_FIELDS 8

typedef void* slnlp_stream_t;

int slnlp_abi_version(void);
const char* slnlp_last_error_string(void);
/* kernels launched through this ABI by the process so far (bench.py's gpu_launches) */
int64_t slnlp_launch_count(void);
/* number of SMs of the current device (grid sizing), or -1 */
int slnlp_device_sm_count(void);
/* a NEW non-blocking CUDA stream on the current device (NULL on failure) / its release.  The host side
 * wraps it as an external stream: framework stream POOLS hand the same stream to several owners, and a
 * fit that captures its step graph on a stream another fit of the process is launching on would swallow
 * that fit's work into the graph (grid search with several fits per GPU). */
void* slnlp_stream_create(void);
/* the same with the device's highest (high != 0) or lowest stream priority: kernels captured from a
 * high-priority stream keep that priority as graph nodes, so the dependency chain of a step (the recurrent
 * kernels) is scheduled ahead of the weight-gradient GEMMs that run beside it on side lanes */
void* slnlp_stream_create_priority(int high);
int slnlp_stream_destroy(void* stream);

/* ---- K1: phonological embedding (nn.Embedding, bkp:49,60,374-379; Transformer
 * model/transformer.py:32-37,106-109).  One fused gather(+concat)(+scale)(+PE).
 * idx [B,T,F] int64; field f reads table + field_off[f] (row width field_w[f]);
 * out row (t*B+b if time_major else b*T+t) is the concat of the F rows, times
 * `scale`, plus pe[t, :] when pe != NULL (pe is [T, sum(field_w)]).
 * field_off / field_w / field_rows (rows of each table) are HOST arrays of length F
 * (F <= SLNLP_MAX_FIELDS).  An index outside [0, field_rows[f]) yields a NaN row
 * (loud, but no fault); the reference would raise an IndexError. */
int slnlp_embed_gather_fwd(const float* table, const int64_t* idx, float* out,
                           int B, int T, int F, const int64_t* field_off, const int* field_w,
                           const int64_t* field_rows, int time_major, float scale,
                           const float* pe, slnlp_stream_t stream);
/* dtable[row] += scale * dout[...]; rows whose index == padding_idx get no gradient
 * (nn.Embedding(padding_idx=...)); pass padding_idx = -1 for none. */
int slnlp_embed_gather_bwd(float* dtable, const int64_t* idx, const float* dout,
                           int B, int T, int F, const int64_t* field_off, const int* field_w,
                           const int64_t* field_rows, int time_major, float scale,
                           int64_t padding_idx, slnlp_stream_t stream);

/* ---- K3/K6/K10 and every other dense contraction (nn.Linear call sites bkp:73-76,
 * 193-194,246,297-299; the x*W_ih^T hoisted out of nn.LSTM/GRU, bkp:114,216).
 * Row-major C[M,N] = op(A) op(B) + bias[N] + beta*C with fp32 FMA accumulation.
 * op(A) is A[M,K] (transA=0) or A[K,M]^T (transA=1); likewise B[K,N] / B[N,K]^T.
 * workspace (may be NULL): scratch for deterministic split-K when M*N is small and K
 * long (the dW GEMMs over K = B*T); slnlp_gemm_workspace_floats() is always enough. */
int64_t slnlp_gemm_workspace_floats(void);
int slnlp_gemm_f32(int transA, int transB, int M, int N, int K,
                   const float* A, int lda, const float* B, int ldb,
                   float* C, int ldc, const float* bias, float beta,
                   float* workspace, int64_t workspace_floats, slnlp_stream_t stream);
/* The same contraction on the 5th-gen tensor cores, TMA-fed (north_star: "x W_ih hoisted into one TMA-fed GEMM
 * over all timesteps"):
 * operands stay fp32 in HBM and are consumed as TF32 by tcgen05.mma kind::tf32 from 128-byte
 * swizzled TMA tiles (4-stage mbarrier ring, warp-specialised producer / MMA / epilogue),
 * transposed operands as MN-major tiles.  Same contract as slnlp_gemm_f32 (shapes the tile kernel does not cover are computed by it); needs 16-byte
 * aligned A, B and lda, ldb multiples of 4, else it computes with slnlp_gemm_f32. */
int slnlp_gemm_tf32(int transA, int transB, int M, int N, int K,
                    const float* A, int lda, const float* B, int ldb,
                    float* C, int ldc, const float* bias, float beta,
                    float* workspace, int64_t workspace_floats, slnlp_stream_t stream);
/* slnlp_gemm_tf32's kernel with fp32-ACCURATE products (the 1e-5 / identical-argmax path on the tensor cores):
 * every operand tile is split in shared memory into the 19 bits the tensor core reads and the exact fp32
 * remainder, and each k-step issues hi*hi + hi*lo + lo*hi (error ~2^-22 per product, fp32 accumulation).
 * Same contract and fallbacks as slnlp_gemm_tf32. */
int slnlp_gemm_tf32x3(int transA, int transB, int M, int N, int K,
                      const float* A, int lda, const float* B, int ldb,
                      float* C, int ldc, const float* bias, float beta,
                      float* workspace, int64_t workspace_floats, slnlp_stream_t stream);
/* The same contraction at data-parallel batch sizes (BASELINE.json configs[3]: the hoisted x W_ih^T of
 * bkp:114 over T*B = 262,144 rows, and its dX / dW twins) on CTA PAIRS: A and B are bf16 in HBM (row strides
 * lda / ldb in ELEMENTS, multiples of 8; 16-byte aligned bases), C / bias / beta as above (fp32).  Persistent
 * 2-CTA clusters, 256 x 256 tiles, tcgen05.mma.cta_group::2 kind::f16 with fp32 accumulators double-buffered
 * in tensor memory, TMA 128-byte-swizzled operand ring, transposed operands consumed in place as MN-major
 * tiles.  Any M, N, K > 0 computes correctly; slnlp_gemm_bf16_supported says whether the shape is worth a
 * 256 x 256 tile (the host keeps smaller products on slnlp_gemm_tf32). */
int slnlp_gemm_bf16_supported(int transA, int transB, int M, int N, int K);
int slnlp_gemm_bf16(int transA, int transB, int M, int N, int K,
                    const uint16_t* A, int64_t lda, const uint16_t* B, int64_t ldb,
                    float* C, int ldc, const float* bias, float beta, slnlp_stream_t stream);
/* dst (bf16) = src (fp32), rows x cols with row strides lds / ldd in elements; transpose != 0 writes
 * dst[c*rows + r] (dense operands only): the K-major copy of a transposed weight matrix. */
int slnlp_cast_bf16(const float* src, int64_t lds, uint16_t* dst, int64_t ldd, int64_t rows, int64_t cols,
                    int transpose, slnlp_stream_t stream);
/* y (bf16) = dropout(x): the mask slnlp_dropout(site) draws, written as the bf16 operand the next layer's
 * CTA-pair GEMM reads (nn.LSTM inter-layer dropout, bkp:100, at data-parallel batch sizes). */
int slnlp_dropout_bf16(const float* x, uint16_t* y, int64_t n, float p, const uint64_t* rng, uint32_t site,
                       slnlp_stream_t stream);
/* the same from a bf16 input, also leaving the keep mask (bit i % 32 of word i / 32; n a multiple of 128) */
int slnlp_dropout_bf16_masked(const uint16_t* xb, uint16_t* y, uint32_t* keep_bits, int64_t n, float p,
                              const uint64_t* rng, uint32_t site, slnlp_stream_t stream);
/* slnlp_colsum_f32 over a bf16 matrix (the bf16 d(pre-activations) of the large-batch path); cols and lda
 * multiples of 4. */
int slnlp_colsum_bf16(const uint16_t* A, int rows, int cols, int64_t lda, float* out, float beta,
                      slnlp_stream_t stream);
/* out[c] = beta*out[c] + sum_r A[r*lda + c]   (bias gradients) */
int slnlp_colsum_f32(const float* A, int rows, int cols, int lda, float* out, float beta,
                     slnlp_stream_t stream);

/* tuning / diagnostics of the persistent recurrent kernel (H = 128, rnn_persistent.cu).
 *   variant >= 0: step-loop variant of the forward kernel (0 = k-major issue, one commit; 1 = gate-major
 *                 issue, early commit, last gate tile overlapped with the early gates' activations; the
 *                 default, or $SLNLP_PERSIST_VAR); -1 keeps the current one;
 *   profile  1/0: launch the instrumented instantiations, which sum %clock deltas per phase of every step
 *                 (0 MMA issue, 1 deferred stores + prefetch issue, 2 MMA wait, 3 tcgen05.ld, 4 gate math +
 *                 h tile store, 5 proxy fence, 6 CTA barrier); -1 keeps the current setting;
 *   out48 != NULL: synchronises the device and copies the table [fwd|bwd][3 threads of CTA (0,0): the MMA
 *                 issuer, thread 160, thread 511][8 phases] (cycles summed over the steps of the last layer
 *                 launched) to host memory.
 * Process-wide switches, not per stream: set them before launching (profiles/prof_persist_phases.py). */
int slnlp_debug_persist_config(int variant, int profile, uint32_t* out48);

/* diagnostic: clusters of the cluster-persistent recurrent kernel (H = 256: 8 CTAs, H = 512: 16 CTAs)
 * the device holds at once, or -1 */
int slnlp_max_active_clusters(int H, int nseq);

/* ---- K4/K8: one (bi)directional recurrent layer over all timesteps (nn.LSTM / nn.GRU
 * on a packed sequence, bkp:95-100,110-123; decoder step bkp:215-216 with T=1).
 * gates [T,B,ndir,G,H]: in = x W_ih^T + b_ih; out = activated gates (stash for bwd).
 * State is frozen and out = 0 for t >= lengths[b]; direction 1 walks t = len-1..0.
 * lengths may be NULL (all T).  h0/c0 [ndir,B,H] may be NULL (zeros).
 * out [T,B,ndir*H]; stash [T,B,ndir,H] (LSTM: c_t; GRU: W_hn h + b_hn);
 * h_final [ndir,B,H] (may be NULL).  precision: 0 = fp32 FMA, 1 = bf16 tcgen05. */
int slnlp_rnn_layer_fwd(int mode, int precision, int T, int B, int H, int ndir,
                        float* gates, const float* w_hh, const float* b_hh,
                        const int64_t* lengths, const float* h0, const float* c0,
                        float* out, float* stash, float* h_final, slnlp_stream_t stream);
/* ---- K8 fused: ONE decoder step of one layer (bkp:215-216; MAX_OUTPUT_LEN = 1, bkp:332) in one launch:
 * gates = x W_ih^T + b_ih + h0 W_hh^T + b_hh, gate nonlinearities, cell update and - between layers - the
 * inter-layer dropout of nn.LSTM / nn.GRU (bkp:190).  fp32 FMA, full-precision expf / tanhf (serves both
 * precision paths).  x [B,D]; h0 / c0 [B,H] (LSTM: c0 may alias h0, bkp:278-279; GRU: c0 NULL);
 * w_ih [G*H,D], w_hh [G*H,H], b_ih / b_hh [G*H].  Outputs: gates [B,G,H] = activated gates and
 * stash [B,H] (LSTM c_1; GRU W_hn h0 + b_hn) - exactly what slnlp_rnn_layer_bwd(T = 1) consumes;
 * h [B,H]; h_drop [B,H] = dropout(h, p_drop) with the mask slnlp_dropout(site) draws, or NULL. */
int slnlp_dec_cell_fwd(int mode, int B, int H, int D, const float* x, const float* h0, const float* c0,
                       const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                       float* gates, float* stash, float* h, float* h_drop, float p_drop,
                       const uint64_t* rng, uint32_t site, slnlp_stream_t stream);

/* Backward of slnlp_dec_cell_fwd as ONE launch (replaces slnlp_rnn_layer_bwd(T = 1) + axpy + the d(input) GEMM + its
 * dropout; bkp:190,215-216,278-279).  In: the forward's gates / stash, h0 / c0 (LSTM: c0 aliases h0; GRU: c0 NULL),
 * dh [B,H] = d(h_1).  Out: dgx [B,G,H] = d(x-side pre-activations) and, GRU only, dnh [B,H] = d(W_hn h0 + b_hn) - the
 * operands of the dW_ih / dW_hh / bias products, written to buffers of their OWN (gates stays intact);
 * dx [B,D] = dgx W_ih, multiplied by the keep / scale factor slnlp_dropout(site) draws when rng != NULL (the layer's
 * input was dropout(h of the layer below)); dh0 [B,H] = d(h0) (LSTM: + d(c0)).  dx and dh0 must not alias dh.
 * H and D multiples of 16 (slnlp_dec_cell_bwd_supported). */
int slnlp_dec_cell_bwd_supported(int mode, int B, int H, int D);
int slnlp_dec_cell_bwd(int mode, int B, int H, int D, const float* gates, const float* stash, const float* h0,
                       const float* c0, const float* dh, const float* w_ih, const float* w_hh, float* dgx,
                       float* dnh, float* dx, float* dh0, float p_drop, const uint64_t* rng, uint32_t site,
                       slnlp_stream_t stream);

/* BPTT of the same layer.  gates: in = activated gates, out = d(x-side pre-activations)
 * (zeros at frozen steps) ready for the hoisted dW_ih / dx GEMMs.  stash: GRU only,
 * out = d(W_hn h + b_hn).  dout [T,B,ndir*H] / dh_final / dc_final may be NULL.
 * dh0 / dc0 [ndir,B,H] may be NULL.  carry: workspace of 2*ndir*B*H floats. */
int slnlp_rnn_layer_bwd(int mode, int precision, int T, int B, int H, int ndir,
                        float* gates, float* stash, const float* out,
                        const float* w_hh, const int64_t* lengths,
                        const float* h0, const float* c0,
                        const float* dout, const float* dh_final, const float* dc_final,
                        float* dh0, float* dc0, float* carry, slnlp_stream_t stream);

/* Extras of the persistent recurrent kernels (H = 128, T > 1, no initial state; ask
 * slnlp_rnn_extras_supported first - the other kernel families reject them), each of which removes
 * launches from the train step's critical path:
 *   hfinal_cat   h_final (fwd) and dh_final / dc_final (bwd) are laid out [B, ndir*H], i.e. already as
 *                concatenate_directions (bkp:155-159) / the bridge input wants them: no concat kernel;
 *   out_drop     fwd: also write dropout(out, p_drop) - the next layer's input under nn.LSTM / nn.GRU's
 *                inter-layer dropout (bkp:100) - with the mask slnlp_dropout(site) draws; NULL = off;
 *   dout_dropped bwd: dout is the gradient of that DROPPED output: the forward's mask (p_drop, rng, site)
 *                is applied while it is read, instead of a separate dropout pass over the gradient. */
typedef struct slnlp_rnn_extras {
  int hfinal_cat;
  float* out_drop;
  float p_drop;
  const uint64_t* rng;
  uint32_t site;
  int dout_dropped;
  const float* mask;  /* optional: the factors slnlp_dropout_mask(site) wrote (out's layout) - then neither kernel
                       * runs Philox in its step loop (ten dependent rounds the tcgen05 kernel cannot hide) */
} slnlp_rnn_extras;
int slnlp_rnn_extras_supported(int precision, int T, int B, int H, int ndir);
int slnlp_rnn_layer_fwd_ex(int mode, int precision, int T, int B, int H, int ndir,
                           float* gates, const float* w_hh, const float* b_hh,
                           const int64_t* lengths, const float* h0, const float* c0,
                           float* out, float* stash, float* h_final, const slnlp_rnn_extras* extras,
                           slnlp_stream_t stream);
int slnlp_rnn_layer_bwd_ex(int mode, int precision, int T, int B, int H, int ndir,
                           float* gates, float* stash, const float* out,
                           const float* w_hh, const int64_t* lengths,
                           const float* h0, const float* c0,
                           const float* dout, const float* dh_final, const float* dc_final,
                           float* dh0, float* dc0, float* carry, const slnlp_rnn_extras* extras,
                           slnlp_stream_t stream);

/* The same recurrence with bf16 OPERANDS for the per-step family (batches beyond the persistent / cluster
 * kernels' reach, BASELINE.json configs[3]: batch 4096, H 512; nn.LSTM recurrence bkp:95-123): h_{t-1} is read
 * from out_bf - a bf16 copy [T][B][ndir*H] of `out` that the forward kernels write next to the fp32 one - and
 * W_hh from a bf16 copy (slnlp_cast_bf16), through tcgen05.mma kind::f16; everything else (hoisted
 * projection, gates, stash, h_final [ndir][B][H]) is fp32 exactly as in slnlp_rnn_layer_fwd.  Zero initial state
 * only.  The backward (LSTM only) reads d(pre-activations)_{t+1} from dg_bf, a bf16 copy [T][B][ndir*G*H] it
 * writes next to the in-place fp32 one (the dX / dW GEMMs read it too), and W_hh^T from w_hhT_bf
 * [ndir][H][G*H] (slnlp_cast_bf16 with transpose); write_f32 = 0 lets a kernel skip the fp32 in-place copy when every
 * consumer of d(pre-activations) reads dg_bf.  slnlp_rnn_bf16_step_supported: 1 where this family is
 * the right one (LSTM, B > 256, H a multiple of 128). */
int slnlp_rnn_bf16_step_supported(int mode, int T, int B, int H, int ndir);
int slnlp_rnn_layer_fwd_bf16(int mode, int T, int B, int H, int ndir, float* gates, const uint16_t* w_hh_bf,
                             const float* b_hh, const int64_t* lengths, float* out, uint16_t* out_bf, float* stash,
                             float* h_final, slnlp_stream_t stream);
int slnlp_rnn_layer_bwd_bf16(int mode, int T, int B, int H, int ndir, float* gates, uint16_t* dg_bf, float* stash,
                             const float* out, const uint16_t* w_hhT_bf, const int64_t* lengths, const float* dout,
                             const float* dh_final, const float* dc_final, float* carry, int write_f32,
                             const uint32_t* dout_keep, float dout_scale, int gates_in_dg, slnlp_stream_t stream);
/* slnlp_rnn_layer_fwd_bf16 with the BPTT stash of the activated gates as bf16: gates_act_bf [T][B][ndir*G*H] (CTA-pair
 * kernels only; `gates` then keeps the hoisted projection).  Pass the SAME buffer as dg_bf with gates_in_dg = 1 to
 * slnlp_rnn_layer_bwd_bf16: a step reads its activated gates there and overwrites them with d(pre-activations) -
 * one bf16 buffer, two lives, 8 instead of 16 bytes per element in each direction. */
int slnlp_rnn_layer_fwd_bf16_ex(int mode, int T, int B, int H, int ndir, float* gates, const uint16_t* w_hh_bf,
                                const float* b_hh, const int64_t* lengths, float* out, uint16_t* out_bf, float* stash,
                                float* h_final, uint16_t* gates_act_bf, slnlp_stream_t stream);
/* 1 where the two calls above run the persistent CTA-pair step kernels (rnn_step_pair.cu: LSTM, two directions,
 * B > 256, H = 256 / 512 / 1024).  Only those accept out = NULL in the forward (nobody reads the fp32 copy of a
 * lower layer's output once its consumers take out_bf) and, in the backward, dout_keep: the keep mask of the
 * inter-layer dropout above this layer (slnlp_dropout_bf16_masked), one bit per element of dout, applied with
 * dout_scale = 1 / (1 - p) while dout is read - instead of a dropout pass over the gradient. */
int slnlp_rnn_bf16_pair_supported(int mode, int T, int B, int H, int ndir);
/* pad_packed_sequence(padding_value) on the top layer (bkp:120-123): rows t >= len
 * of x [T,B,W] are set to `value` (1.0 going forward, 0.0 before BPTT). */
int slnlp_pad_fill(float* x, const int64_t* lengths, int T, int B, int W, float value,
                   slnlp_stream_t stream);
/* out-of-place: dst = src with rows t >= lengths[b] set to `value`.  The encoder output then exists twice -
 * zero-padded (what BPTT and dW_hh read) and pad-filled (what the key projection / attention read, bkp:121-123) -
 * and nothing has to be un-filled on the backward critical path. */
int slnlp_pad_fill_copy(const float* src, float* dst, const int64_t* lengths, int T, int B, int W, float value,
                        slnlp_stream_t stream);
/* concatenate_directions (bkp:155-159): [ndir,B,H] -> [B, ndir*H], and its inverse */
int slnlp_concat_dirs(const float* h_final, float* enc_final, int B, int H, int ndir,
                      int inverse, slnlp_stream_t stream);

/* ---- elementwise pieces: tanh of the bridge (bkp:276), inter-layer dropout
 * (nn.LSTM dropout=p, bkp:100,190), decoder input [emb_bos || ctx] (bkp:215). */
int slnlp_tanh_fwd(float* x, int64_t n, slnlp_stream_t stream);
/* dx = dy * (1 - y*y) in place on dy */
int slnlp_tanh_bwd(float* dy, const float* y, int64_t n, slnlp_stream_t stream);
/* y = x * keep/(1-p), keep ~ Philox(rng[0] seed, rng[1] step, site, i); rng is a
 * device array of two uint64.  In-place (y == x) allowed; bwd is the same call. */
int slnlp_dropout(const float* x, float* y, int64_t n, float p, const uint64_t* rng,
                  uint32_t site, slnlp_stream_t stream);
/* y[i] = the keep / scale factor (0 or 1/(1-p)) slnlp_dropout(site) applies to element i under the same rng state */
int slnlp_dropout_mask(float* y, int64_t n, float p, const uint64_t* rng, uint32_t site, slnlp_stream_t stream);
int slnlp_rng_advance(uint64_t* rng, slnlp_stream_t stream);
/* dst[b, 0:E] = row[E]; dst[b, E:E+W] = src[b, 0:W]  (B rows) */
int slnlp_dec_input_fwd(const float* row, const float* src, float* dst, int B, int E, int W,
                        slnlp_stream_t stream);
/* drow[E] += sum_b ddst[b,0:E]; dsrc[b,:] = ddst[b,E:E+W] */
int slnlp_dec_input_bwd(const float* ddst, float* drow, float* dsrc, int B, int E, int W,
                        slnlp_stream_t stream);
int slnlp_axpy(float* y, const float* x, float a, int64_t n, slnlp_stream_t stream);

/* ---- K7: Bahdanau attention, one fused decode step (bkp:304-327).
 * q [B,H] (= query_layer(h)), pk [T,B,H] (= key_layer(enc_out)), v [H] energy weights,
 * val [T,B,W] (enc_out, W = 2H), X [B,T] int64 tokens (mask = X != pad_idx).
 * alpha [B,T], ctx [B,W]. */
int slnlp_attn_step_fwd(const float* q, const float* pk, const float* v, const float* val,
                        const int64_t* X, int64_t pad_idx, int T, int B, int H, int W,
                        float* alpha, float* ctx, slnlp_stream_t stream);
/* dval [T,B,W] (written), dpk [T,B,H] (written), dq [B,H] (written),
 * dv_part [B,H] per-sequence partials of d(energy weight) (reduce with colsum). */
int slnlp_attn_step_bwd(const float* dctx, const float* q, const float* pk, const float* v,
                        const float* val, const float* alpha, int T, int B, int H, int W,
                        float* dval, float* dpk, float* dq, float* dv_part,
                        slnlp_stream_t stream);

/* ---- K7b: the decoder head of the single decode step, one launch per direction of autograd, one CTA per sequence:
 * bridge (bkp:268-280: hidden0[l] = tanh(W_b enc_final[l] + b_b), every layer) -> query (bkp:312) -> the attention of
 * slnlp_attn_step_fwd -> decoder input [trg_embed[bos] || ctx] (bkp:202-216).  enc_final [L,B,2H], hidden0 [L,B,H],
 * dec_xin [B,E+2H]; the other arguments as slnlp_attn_step_fwd (W = 2H).  w_bridge == NULL: hidden0 is an INPUT (the
 * caller ran the bridge); w_query == NULL: q is an input.  A fused bridge needs the fused query.  ctx may be NULL. */
int slnlp_dec_head_supported(int T, int B, int H, int L, int fuse_query, int fuse_bridge);
int slnlp_dec_head_fwd(int T, int B, int H, int L, int E, const float* enc_final, const float* w_bridge,
                       const float* b_bridge, const float* w_query, const float* pk, const float* v,
                       const float* val, const int64_t* X, int64_t pad_idx, const float* bos_row,
                       float* hidden0, float* q, float* alpha, float* ctx, float* dec_xin,
                       slnlp_stream_t stream);
/* Backward twin.  d_decx [B,E+2H] = d(decoder input) (its last 2H columns are d(ctx)); dval / dpk / dq / dv_part as
 * slnlp_attn_step_bwd.  w_query != NULL: d_hidden0[L-1] += dq W_q (d_hidden0 [L,B,H] holds the decoder cells' d(h0) on
 * entry).  w_bridge != NULL: d_hidden0 *= 1 - hidden0^2 (every layer, written back: the operand of the bridge's weight
 * gradient) and d_enc_final [L,B,2H] = d_hidden0 W_b. */
int slnlp_dec_head_bwd(int T, int B, int H, int L, int E, const float* d_decx, const float* q, const float* pk,
                       const float* v, const float* val, const float* alpha, const float* hidden0,
                       const float* w_query, const float* w_bridge, float* dval, float* dpk, float* dq,
                       float* dv_part, float* d_hidden0, float* d_enc_final, slnlp_stream_t stream);

/* ---- K10: generator log_softmax (bkp:75-76) + skorch CrossEntropyLoss(ignore_index)
 * applied on the log-probs (config/*.yaml:36, helper.py:67-70).
 * logits [B,V] -> logp [B,V]. */
int slnlp_log_softmax_fwd(const float* logits, float* logp, int B, int V, slnlp_stream_t stream);
/* dlogits = dlogp - exp(logp) * rowsum(dlogp).  dlogits rows are ld_dlogits floats apart (>= V; a
 * multiple of 4 keeps the generator's dW / dx GEMMs on the TMA path when V_tgt is not one). */
int slnlp_log_softmax_bwd(const float* dlogp, const float* logp, float* dlogits, int B, int V,
                          int ld_dlogits, slnlp_stream_t stream);
/* loss_out[0] = mean over y != ignore of -log_softmax(logp)[y]; loss_out[1] = count.
 * dlogits (may be NULL) = d loss / d logits through both log-softmaxes. */
int slnlp_ce_on_logp(const float* logp, const int64_t* y, int64_t ignore_index, int B, int V,
                     float* loss_out, float* dlogits, int ld_dlogits, float* row_ws, slnlp_stream_t stream);
/* The two above in ONE pass per row (north_star: "fused CrossEntropy+log-softmax"): logits -> logp
 * (written), loss_out = {mean loss, n_valid}, dlogits (may be NULL) with row stride ld_dlogits. */
int slnlp_logsoftmax_ce_fused(const float* logits, const int64_t* y, int64_t ignore_index, int B, int V,
                              float* logp, float* loss_out, float* dlogits, int ld_dlogits, float* row_ws,
                              slnlp_stream_t stream);
/* the mean-loss reduction of slnlp_logsoftmax_ce_fused as a call of its own: with loss_out = NULL there, the row
 * losses stay in row_ws and this reduces them to loss_out = {mean loss over valid rows, n_valid} - nothing on the
 * backward chain reads the loss, so the host runs it on a side stream. */
int slnlp_ce_reduce(const float* row_ws, int B, float* loss_out, slnlp_stream_t stream);

/* ---- K11/K12: GradientNormClipping -> clip_grad_norm_(max_norm, 2) (helper.py:227-229)
 * and torch.optim.SGD(momentum, nesterov=False) (config/*.yaml:39-42), over flat buffers.
 * partials: ZERO-INITIALISED workspace of slnlp_sumsq_partials() floats (block partial sums + the ticket
 * counter of the last-block reduction, which the kernel resets itself); norm_out[0] = ||g||_2.  One launch. */
int slnlp_sumsq_partials(void);
int slnlp_gradnorm(const float* g, int64_t n, float* partials, float* norm_out,
                   slnlp_stream_t stream);
/* hyper (device) = {lr, momentum, max_norm (<=0: no clipping), first_step flag}.
 * coef = min(1, max_norm/(norm+1e-6)); buf = first ? g*coef : momentum*buf + g*coef;
 * p -= lr*buf.  grad_scale multiplies g first (1/world_size after an all-reduce). */
int slnlp_sgd_momentum_clip(float* p, const float* g, float* buf, int64_t n,
                            const float* hyper, const float* norm, float grad_scale,
                            slnlp_stream_t stream);
/* the same update, and g <- 0 behind it: the gradient buffer is consumed and ready for the next step's
 * accumulating weight-gradient GEMMs (the fused train step then needs no separate zero-fill launch) */
int slnlp_sgd_momentum_clip_zero(float* p, float* g, float* buf, int64_t n, const float* hyper,
                                 const float* norm, float grad_scale, slnlp_stream_t stream);

/* ---- K13: nn.Transformer pieces (model/transformer.py:40-45,82-87; post-norm, ReLU, LayerNorm
 * eps 1e-5).  The projections are slnlp_gemm_*; the FFN activation is slnlp_relu_*.
 *
 * Scaled-dot-product attention of nn.MultiheadAttention, FlashAttention-style (no S x S matrix
 * in HBM).  Row (b*Sq + i) of q / o and row (b*Sk + j) of k / v hold all heads; head h is
 * columns [h*dh, (h+1)*dh); ldq/ldk/ldv/ldo are row strides in floats, so q, k, v may point
 * into one packed in-projection buffer.  dh a multiple of 4, <= 256.
 * causal != 0: key j > query i is masked (the reference applies this to the ENCODER self-
 * attention, model/transformer.py:68,84).  key_tokens [B,Sk] int64 (may be NULL): key j of
 * sequence b is masked when key_tokens[b,j] == pad_idx (src_key_padding_mask, util.py:45-61).
 * A row whose keys are all masked yields NaN, as torch.  lse [B,nhead,Sq] is saved for backward.
 * p_drop > 0: dropout on the attention weights with Philox(rng, site, element). */
int slnlp_mha_fwd(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                  float* o, int ldo, float* lse, int B, int Sq, int Sk, int nhead, int dh,
                  int causal, const int64_t* key_tokens, int64_t pad_idx, float p_drop,
                  const uint64_t* rng, uint32_t site, slnlp_stream_t stream);
/* o = forward output, dout = d o; dvec: workspace [B,nhead,Sq].  dq / dk / dv are written (not
 * accumulated) with the leading dimensions of q / k / v.  Deterministic (no atomics). */
int slnlp_mha_bwd(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                  const float* o, const float* dout, int ldo, const float* lse, float* dvec,
                  float* dq, float* dk, float* dv, int B, int Sq, int Sk, int nhead, int dh,
                  int causal, const int64_t* key_tokens, int64_t pad_idx, float p_drop,
                  const uint64_t* rng, uint32_t site, slnlp_stream_t stream);
/* Tensor-core forms of the two calls above (same arguments and semantics): every tile product of
 * the flash passes runs as warp-level tf32 MMA with fp32 accumulation (the "bf16" precision mode
 * of the modules, tolerance 2e-2; the softmax, masks and dropout stay fp32).  Head dimensions 16,
 * 32 and 64; any other shape runs the fp32 kernels.  Queries of <= 4 rows per sequence (the
 * reference's one-position decoder, model/transformer.py:82-87) use a one-warp-per-(sequence,
 * head) fp32 kernel in both forms. */
int slnlp_mha_tf32_fwd(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                       float* o, int ldo, float* lse, int B, int Sq, int Sk, int nhead, int dh,
                       int causal, const int64_t* key_tokens, int64_t pad_idx, float p_drop,
                       const uint64_t* rng, uint32_t site, slnlp_stream_t stream);
int slnlp_mha_tf32_bwd(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                       const float* o, const float* dout, int ldo, const float* lse, float* dvec,
                       float* dq, float* dk, float* dv, int B, int Sq, int Sk, int nhead, int dh,
                       int causal, const int64_t* key_tokens, int64_t pad_idx, float p_drop,
                       const uint64_t* rng, uint32_t site, slnlp_stream_t stream);
/* y = LayerNorm(x + res) * gamma + beta over rows of E floats (res may be NULL;
 * E <= 1024).  mean / rstd [rows] (may be NULL at inference) are saved for backward. */
int slnlp_add_layernorm_fwd(const float* x, const float* res, const float* gamma, const float* beta,
                            float* y, float* mean, float* rstd, int rows, int E, float eps,
                            slnlp_stream_t stream);
/* dx (= d x = d res) from dy; accumulate != 0 adds into dx.  partials: workspace of
 * slnlp_ln_bwd_blocks(rows) * 2 * E floats; row-block sums of d gamma (first E) and d beta
 * (next E) - reduce them with slnlp_colsum_f32(partials, blocks, 2E, 2E, ...). */
int slnlp_ln_bwd_blocks(int rows);
int slnlp_layernorm_bwd(const float* dy, const float* x, const float* res, const float* gamma,
                        const float* mean, const float* rstd, float* dx, float* partials,
                        int rows, int E, int accumulate, slnlp_stream_t stream);

int slnlp_relu_fwd(float* x, int64_t n, slnlp_stream_t stream);
/* dx = dy * (y > 0) in place on dy */
int slnlp_relu_bwd(float* dy, const float* y, int64_t n, slnlp_stream_t stream);
/* x = dropout(relu(x), p) in place, one pass, with the mask slnlp_dropout(site) draws (nn.TransformerEncoderLayer /
 * DecoderLayer: linear2(dropout(relu(linear1(x)))), model/transformer.py:40-45); and its backward
 * dy = y > 0 ? dy / (1 - p) : 0 on the stored y (positive exactly where the unit was active and kept: no random numbers). */
int slnlp_relu_dropout_fwd(float* x, int64_t n, float p, const uint64_t* rng, uint32_t site, slnlp_stream_t stream);
int slnlp_relu_dropout_bwd(float* dy, const float* y, int64_t n, float p, slnlp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SLNLP_B200_H */
