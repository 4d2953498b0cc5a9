/* mmvqa.h -- C ABI of libmmvqa_sm100.so: the B200 (sm_100a) kernels behind the MMBERT
 * fusion-encoder hot path of DannielSilva/MM-VQA.
 *
 * The reference has no FFI: its boundary is the Python nn.Module surface of models/
 * (SURVEY.md section 8b).  Each entry point below names the reference code whose forward or
 * backward arithmetic it replaces (file:line relative to the reference root).  The host
 * side (mmvqa_b200/*.py) mirrors the reference classes and calls these through ctypes
 * from torch.autograd.Function.forward/backward.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / pybind types.
 *   - every pointer is a DEVICE pointer unless the name ends in _host; buffers (including
 *     workspaces) are owned and allocated by the caller; nothing here allocates.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and the call
 *     returns without synchronising (CUDA-graph capturable).
 *   - return 0 on success, a negative MMVQA_ERR_* otherwise; mmvqa_last_error() returns a
 *     thread-local message.  Nothing throws.
 *   - `dtype` selects the activation storage type: MMVQA_F32 (fp32-accurate validation
 *     path, SIMT FFMA GEMMs) or MMVQA_BF16 (production path: tcgen05/TMEM GEMMs fed by
 *     TMA, fp32 accumulate).  Parameter vectors (bias, gamma, beta), LayerNorm statistics,
 *     RealFormer scores and all gradients of parameters are fp32 in both modes.
 *   - masks are float 1.0 / 0.0 of shape [B, T].
 */
#ifndef MMVQA_H_
#define MMVQA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMVQA_ABI_VERSION 4

enum { MMVQA_F32 = 0, MMVQA_BF16 = 1 };
enum { MMVQA_ACT_NONE = 0, MMVQA_ACT_SERF = 1, MMVQA_ACT_GELU = 2, MMVQA_ACT_RELU = 3 };
enum {
  MMVQA_OK = 0,
  MMVQA_ERR_ARG = -1,      /* bad shape / alignment / enum */
  MMVQA_ERR_CUDA = -2,     /* a CUDA runtime or driver call failed */
  MMVQA_ERR_ARCH = -3,     /* device is not sm_100 */
  MMVQA_ERR_SMEM = -4      /* problem does not fit the kernel's shared-memory budget */
};

typedef void* mmvqa_stream_t;

int mmvqa_abi_version(void);
const char* mmvqa_last_error(void);
/* compute capability major*10+minor of the current device, or a negative error */
int mmvqa_device_sm(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
int64_t mmvqa_launch_count(void);
/* Dropout under CUDA-graph replay.  nn.Dropout (models/realformer.py:22-26,44, transformer.py:26,45,58, BertEmbeddings)
 * draws a new mask on every call; a captured graph would freeze the by-value seeds.  When `counter` (a DEVICE uint64,
 * NULL = off) is registered, every kernel launched afterwards that draws a dropout mask uses
 * seed + *counter * 0x9E3779B97F4A7C15 instead of seed, read on the device at run time; the graph itself increments
 * the counter once per replay, so forward and backward of one replay agree and consecutive replays differ. */
int mmvqa_set_dropout_counter(const uint64_t* counter);

/* ------------------------------------------------------------------------------------
 * Caption-similarity mask.  replaces SimilarityCalculator.jaccard /
 * jaccard_similarity, models/SupConLoss/supcon_utils.py:110-138 (two nested Python loops over word sets).
 * ids_a / ids_b: [n, lmax] int32, each row the SORTED UNIQUE word ids of one document (padding ignored),
 * len_a / len_b: [n] int32 set sizes.  mask[c1, c2] = 1 if c1 == c2 else |A_c1 & B_c2| / |A_c1 | B_c2| (0 when the
 * union is empty), fp32, computed as the reference does (double division, rounded once): bit-exact.
 * ---------------------------------------------------------------------------------- */
int mmvqa_jaccard_mask(const int* ids_a, const int* len_a, const int* ids_b, const int* len_b, float* mask, int na,
                       int nb, int lmax, mmvqa_stream_t stream);

/* ------------------------------------------------------------------------------------
 * GEMM:  C[M,N] = epilogue( sum_k opA(A)[m,k] * opB(B)[n,k] )
 *   a_trans = 0: A stored [M,K] row-major (K contiguous);  1: stored [K,M] (M contiguous)
 *   b_trans = 0: B stored [N,K] row-major (nn.Linear weight layout); 1: stored [K,N]
 *   replaces: every nn.Linear / einsum contraction on the path -- models/transformer.py:20,48,78
 *   models/realformer.py:33,45,22-26, models/mmbert.py:133-148,164-166, image_encoding.py:74-113
 *   (1x1 conv), SupConLoss/loss.py:69-71, and their autograd backward (dgrad / wgrad).
 * Epilogues (acc = fp32 accumulator, bias fp32 [N] optional everywhere):
 *   EPI_STORE      C = (acc + bias) * (rowscale ? rowscale[batch*M + m] * scale : 1)
 *   EPI_ACT        aux_out = acc + bias (pre-activation, optional) ; C = act(acc + bias)
 *   EPI_RESIDUAL   C = dropout(acc + bias) + aux_in[m,n]   (nn.Dropout on the branch output,
 *                  realformer.py:45,26 / transformer.py:79,86; dropout_p = 0 -> identity.  The keep
 *                  decision is a counter hash of (dropout_seed, m*N+n), reproduced by mmvqa_dropout)
 *   EPI_DACT       C = acc * act'(aux_in[m,n])          (dgrad through an activation)
 *   EPI_ACT_ROWSUM rowsum_out[batch*M + m] += scale * sum_n act(acc)   (visual-token pooling;
 *                  C unused; columns n >= N are masked).  If aux_out != NULL, act'(acc) is stored at
 *                  aux_out[(batch*M + m) * ld_aux_out + n] for the backward pass (the [B,hidden,H,W]
 *                  activation map itself is never written)
 *   EPI_DACT_SCALE C = act'(acc) * rowscale[batch*M + m] * scale  (projector backward recompute)
 * c_dtype may differ from dtype (fp32 output for weight gradients / logits).
 * accumulate != 0: C += result with fp32 atomics (C must be fp32; used with split_k > 1
 * or batch-reduction); the caller zero-fills C first.
 * batch > 1: operands advance by a_batch_rows / b_batch_rows rows of their stored matrix
 * per batch, C by c_batch_stride elements (0 with accumulate = sum over the batch).
 * ---------------------------------------------------------------------------------- */
enum { MMVQA_EPI_STORE = 0, MMVQA_EPI_ACT = 1, MMVQA_EPI_RESIDUAL = 2, MMVQA_EPI_DACT = 3,
       MMVQA_EPI_ACT_ROWSUM = 4, MMVQA_EPI_DACT_SCALE = 5 };

typedef struct mmvqa_gemm_args {
  int dtype;                 /* operand dtype: MMVQA_F32 -> SIMT FFMA, MMVQA_BF16 -> tcgen05 */
  int M, N, K;
  const void* A; int64_t lda; int a_trans;
  const void* B; int64_t ldb; int b_trans;
  void* C; int64_t ldc; int c_dtype;
  const float* bias;         /* [N] or NULL */
  int epilogue; int act;
  const void* aux_in; int64_t ld_aux_in;     /* dtype = `dtype` */
  void* aux_out; int64_t ld_aux_out;         /* dtype = `dtype` */
  float* rowsum_out;         /* EPI_ACT_ROWSUM */
  float* colsum_out;         /* optional, any storing epilogue: colsum_out[n] += sum_m C[m,n] (fp32 atomics, caller
                                zero-fills) -- the bias gradient of the layer below, fused into the dgrad GEMM */
  const float* rowscale;     /* EPI_DACT_SCALE */
  float scale;
  int accumulate;
  int split_k;               /* >= 1 */
  int batch; int64_t a_batch_rows, b_batch_rows, c_batch_stride;
  float dropout_p; uint64_t dropout_seed;    /* EPI_RESIDUAL only: C = dropout(acc + bias) + aux_in */
  int64_t c_split_stride;    /* split_k > 1 without accumulate: split s stores its fp32 partial tile at
                                C + s * c_split_stride (no atomics, no zero-fill; bias goes into split 0).  The
                                consumer sums the slabs: mmvqa_add_layernorm_fwd_parts / mmvqa_layernorm_bwd_parts */
  int b_static;              /* non-zero: B (a weight matrix) is not written by the kernel that precedes this one on the
                                stream, so with programmatic dependent launch its first tiles may be fetched while
                                that kernel is still running */
  uint64_t* trace;           /* optional (tuning only): [ctas, 16] device buffer; the bf16 kernel stores %globaltimer stamps
                                of its phases per CTA (entry, setup, dependency wait, loads issued, first tile landed,
                                last MMA issued, accumulator ready, epilogue done) */
} mmvqa_gemm_args;

int mmvqa_gemm(const mmvqa_gemm_args* args, mmvqa_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Elementwise / reductions
 * ---------------------------------------------------------------------------------- */
/* y = act(x + bias).  replaces models/serf.py:23-24 (5 ATen kernels), transformer.py:7-8 */
int mmvqa_bias_act_fwd(const void* x, const float* bias, void* y, int64_t rows, int cols, int act, int dtype,
                       mmvqa_stream_t stream);
/* dx = dy * act'(x + bias) */
int mmvqa_bias_act_bwd(const void* x, const float* bias, const void* dy, void* dx, int64_t rows, int cols, int act,
                       int dtype, mmvqa_stream_t stream);
/* out[c] = sum_r x[r,c]  (bias gradients); out is overwritten */
int mmvqa_colsum(const void* x, int64_t ldx, float* out, int64_t rows, int cols, int dtype, mmvqa_stream_t stream);
/* dst = (dst_dtype) src, n elements; fp32 <-> bf16 casts of weights and feature maps.
 * rows/cols/ld form: copies [rows, cols] from src (ld_src) to dst (ld_dst), zero-filling cols..ld_dst */
int mmvqa_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, mmvqa_stream_t stream);
int mmvqa_cast_pad(const void* src, int src_dtype, int64_t ld_src, void* dst, int dst_dtype, int64_t ld_dst,
                   int64_t rows, int cols, mmvqa_stream_t stream);
/* Hint: pull up to MMVQA_PREFETCH_MAX address ranges into L2 (prefetch.global.L2, one request per 128-byte line, a
 * few CTAs).  No result, no dependency: launched on a side stream one encoder layer ahead of the main chain so the next
 * layer's bf16 weights (forward: models/realformer.py:47-51 walks 12 layers x 14 MB, more than L2 keeps) and saved
 * activations (backward) are L2 hits when the latency-bound M = B*T kernels ask for them. */
#define MMVQA_PREFETCH_MAX 16
typedef struct mmvqa_prefetch_list {
  const void* ptr[MMVQA_PREFETCH_MAX];
  int64_t bytes[MMVQA_PREFETCH_MAX];
  int n;
} mmvqa_prefetch_list;
int mmvqa_l2_prefetch(const mmvqa_prefetch_list* list, int ctas, mmvqa_stream_t stream);
/* Several fp32 / bf16 -> bf16 cast_pad problems in ONE launch (the five feature maps of the pyramid, models/image_encoding.py:
 * 71-87: five launches of 4-13 us each headed the critical path of every step).  ld_dst % 4 == 0, dst 8-byte aligned. */
#define MMVQA_CAST_MULTI_MAX 8
typedef struct mmvqa_cast_list {
  const void* src[MMVQA_CAST_MULTI_MAX];       /* fp32, or bf16 where src_bf16[i] != 0 (a pure re-pad of the rows) */
  void* dst[MMVQA_CAST_MULTI_MAX];             /* bf16 */
  int64_t ld_src[MMVQA_CAST_MULTI_MAX], ld_dst[MMVQA_CAST_MULTI_MAX], rows[MMVQA_CAST_MULTI_MAX];
  int cols[MMVQA_CAST_MULTI_MAX];
  int src_bf16[MMVQA_CAST_MULTI_MAX];
  int n;
} mmvqa_cast_list;
int mmvqa_cast_pad_multi(const mmvqa_cast_list* list, mmvqa_stream_t stream);
/* x *= *scalar (device scalar); used to apply the incoming loss gradient without a host sync */
int mmvqa_scale_by_device_scalar(void* x, int dtype, const float* scalar, float host_factor, int64_t n,
                                 mmvqa_stream_t stream);

/* y = x * keep / (1-p) with keep = hash(seed, element index) >= p: the same mask the EPI_RESIDUAL
 * epilogue applied in forward, used on the incoming gradient in backward. */
int mmvqa_dropout(const void* x, void* y, int64_t n, float p, uint64_t seed, int dtype, mmvqa_stream_t stream);

/* fused residual + LayerNorm.  replaces transformer.py:78,83 (norm1), realformer.py:49-50 (ln1/ln2),
 * mmbert.py:136 (classifier[1]).  y = LN(x + res) * gamma + beta; sum_out (optional) = x + res;
 * mean/rstd [rows] fp32 are saved for backward. */
int mmvqa_add_layernorm_fwd(const void* x, const void* res, const float* gamma, const float* beta, void* y,
                            void* sum_out, float* mean, float* rstd, int64_t rows, int cols, float eps, int dtype,
                            mmvqa_stream_t stream);
/* dx = d(LN)/d(xsum) . dy (+ dres_extra if non-NULL: an extra gradient of the same shape added into dx,
 * i.e. the residual branch); dgamma/dbeta [cols] are ACCUMULATED with atomics (caller zero-fills).
 * Optional fused outputs for the branch that was added to the residual before the LayerNorm:
 *   dx_drop (dtype `dtype`) = dropout(dx) with the forward mask of (dropout_p, dropout_seed)  [= dx if p == 0]
 *   dxsum [cols] fp32      += column sums of dx_drop (the bias gradient of that branch's last Linear). */
int mmvqa_layernorm_bwd(const void* dy, const void* xsum, const float* gamma, const float* mean, const float* rstd,
                        const void* dx_extra, void* dx, float* dgamma, float* dbeta, void* dx_drop, float* dxsum,
                        float dropout_p, uint64_t dropout_seed, int64_t rows, int cols, int dtype,
                        mmvqa_stream_t stream);

/* The same two LayerNorm passes fed by the fp32 split-K partial tiles of mmvqa_gemm (c_split_stride): at the
 * reference's fine-tune shape (M = 16 x 28 = 448 rows) a K = 3072 GEMM only fills the chip when K is split, and the
 * slab reduction rides on the LayerNorm pass that follows it in realformer.py:49-50 instead of atomics.
 *   fwd:  s = dropout_p(sum_i parts[i]) + res (rounded to `dtype`, stored in sum_out);  y = LN(s) * gamma + beta
 *   bwd:  dy = sum_i dy_parts[i] + dy_res (dy_res may be NULL); everything else as mmvqa_layernorm_bwd.
 * parts[i] = parts + i * part_stride, each [rows, cols] fp32 contiguous.  cols % 8 == 0 (bf16) / % 4 (f32), <= 1024. */
int mmvqa_add_layernorm_fwd_parts(const float* parts, int nparts, int64_t part_stride, const void* res,
                                  const float* gamma, const float* beta, void* y, void* sum_out, float* mean,
                                  float* rstd, int64_t rows, int cols, float eps, float dropout_p,
                                  uint64_t dropout_seed, int dtype, mmvqa_stream_t stream);
int mmvqa_layernorm_bwd_parts(const float* dy_parts, int nparts, int64_t part_stride, const void* dy_res,
                              const void* xsum, const float* gamma, const float* mean, const float* rstd, void* dx,
                              float* dgamma, float* dbeta, void* dx_drop, float* dxsum, float dropout_p,
                              uint64_t dropout_seed, int64_t rows, int cols, int dtype, mmvqa_stream_t stream);

/* LayerNorm backward with the column sums DEFERRED (transformer.py:78,83 / realformer.py:49-50 backward at M = B*T rows):
 * the 112 CTAs of the small-batch problem each flushing 3 x cols sums with atomics onto the same addresses was the
 * slowest part of a kernel that sits twice on the critical path of every layer.  Here every CTA stores its sums to
 * partials[cta][3][cols] (dgamma | dbeta | column sums of dx_drop) with plain stores and returns; the caller folds them
 * with mmvqa_ln_partials_reduce on a side stream, next to the weight-gradient GEMMs nothing waits for.
 * dy may be NULL when dy_parts is given (split-K slabs, as mmvqa_layernorm_bwd_parts).  partial_rows must equal
 * mmvqa_layernorm_bwd_partial_rows(rows, cols, dtype) (0 = this shape cannot use the deferred form). */
int mmvqa_layernorm_bwd_partial_rows(int64_t rows, int cols, int dtype);
int mmvqa_layernorm_bwd_deferred(const void* dy, const float* dy_parts, int nparts, int64_t part_stride, const void* xsum,
                                 const float* gamma, const float* mean, const float* rstd, void* dx, void* dx_drop,
                                 float dropout_p, uint64_t dropout_seed, int64_t rows, int cols, int dtype, float* partials,
                                 int partial_rows, mmvqa_stream_t stream);
/* dgamma / dbeta / dxsum [cols] (any may be NULL) += sum over the partial rows (fp32 atomics, caller zero-fills) */
int mmvqa_ln_partials_reduce(const float* partials, int partial_rows, int cols, float* dgamma, float* dbeta, float* dxsum,
                             mmvqa_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Attention (short sequence: T <= 128, head dim <= 128; one CTA per (batch, head))
 * ---------------------------------------------------------------------------------- */
/* Transformer MHSA core, models/transformer.py:21-27.  qkv [B*T, 3*heads*d] packed q|k|v
 * (output of the fused QKV GEMM), key-side mask, probs [B,heads,T,T] (dtype `dtype`) is the
 * post-softmax matrix the reference keeps in self.scores and is reused by backward.
 * dropout_p > 0 applies nn.Dropout to the probabilities before P.v (transformer.py:26) with the
 * counter-hash mask of (dropout_seed, element index); probs holds the pre-dropout values. */
int mmvqa_mhsa_fwd(const void* qkv, const float* mask, void* out, void* probs, int B, int T, int heads, int d,
                   float dropout_p, uint64_t dropout_seed, int dtype, mmvqa_stream_t stream);
int mmvqa_mhsa_bwd(const void* qkv, const void* probs, const void* dout, void* dqkv, int B, int T, int heads, int d,
                   float dropout_p, uint64_t dropout_seed, int dtype, mmvqa_stream_t stream);
/* RealFormer residual attention core, models/realformer.py:33-44.  kqv [B*T*heads, 3*d] packed
 * k|q|v per (token, head) (output of the shared-weight kqv GEMM).  prev / scores / dscores are
 * fp32 in the kernel-native layout [B, heads, T, T] (the Python side exposes the reference's
 * [B,T,T,heads] as a permuted view).  scores = q.k/sqrt(d) + prev - 10000*(1-mask[b,i]) is both
 * the softmax input and the tensor handed to the next layer.  Backward adds dscores_in (gradient
 * arriving from the next layer through prev) and writes the total as dprev. */
int mmvqa_rf_attn_fwd(const void* kqv, const float* prev, const float* mask, void* out, float* scores, int B, int T,
                      int heads, int d, int dtype, mmvqa_stream_t stream);
/* Same forward with the shared-weight kqv projection (realformer.py:13,33) inside the kernel (bf16 path): x is the layer
 * input [B*T, heads*d], wkqv the [3d, d] weight; kqv_out [B*T*heads, 3d] is WRITTEN (the backward pass reads it).
 * Removes one GEMM launch per layer from the latency-bound small-batch chain. */
int mmvqa_rf_attn_fwd_fused(const void* x, const void* wkqv, const float* prev, const float* mask, void* out, float* scores,
                            void* kqv_out, int B, int T, int heads, int d, int dtype, mmvqa_stream_t stream);
int mmvqa_rf_attn_bwd(const void* kqv, const float* scores, const void* dout, const float* dscores_in, void* dkqv,
                      float* dprev, int B, int T, int heads, int d, int dtype, mmvqa_stream_t stream);
/* Same backward with the input gradient of the kqv projection inside (bf16 path):
 * dx [B*T, heads*d] = dkqv . wkqv + dres  (dres = gradient of the residual connection around the attention block, may be
 * NULL).  dkqv is still written (the weight-gradient GEMM reads it).  Removes the dgrad GEMM launch from the chain. */
int mmvqa_rf_attn_bwd_fused(const void* kqv, const float* scores, const void* dout, const float* dscores_in, void* dkqv,
                            float* dprev, const void* wkqv, const void* dres, void* dx, int B, int T, int heads, int d,
                            int dtype, mmvqa_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Input fusion and pooling
 * ---------------------------------------------------------------------------------- */
/* BertEmbeddings (word+pos+type -> LN eps) fused with the visual-token overwrite of positions
 * 0..nvis-1.  replaces models/mmbert.py:60-67 (B*nvis Python-level copies) + HF BertEmbeddings.
 * vis is [nvis, B, H] fp32.  dropout_p is BertEmbeddings' dropout (text positions only). */
int mmvqa_embed_ln_scatter_fwd(const int64_t* ids, const int64_t* seg, const float* word, const float* pos,
                               const float* typ, const float* gamma, const float* beta, const float* vis, void* h,
                               float* mean, float* rstd, int B, int T, int H, int nvis, int vocab, float eps,
                               float dropout_p, uint64_t dropout_seed, int dtype, mmvqa_stream_t stream);
/* dword/dpos/dtyp/dgamma/dbeta are accumulated with atomics (caller zero-fills); word row
 * `padding_idx` receives no gradient (nn.Embedding(padding_idx=0)); dvis [nvis,B,H] is overwritten. */
int mmvqa_embed_ln_scatter_bwd(const void* dh, const int64_t* ids, const int64_t* seg, const float* word,
                               const float* pos, const float* typ, const float* gamma, const float* mean,
                               const float* rstd, float* dword, float* dpos, float* dtyp, float* dgamma,
                               float* dbeta, float* dvis, int B, int T, int H, int nvis, int padding_idx,
                               float dropout_p, uint64_t dropout_seed, int dtype, mmvqa_stream_t stream);
/* mask-weighted mean over tokens, models/mmbert.py:169-172 */
int mmvqa_masked_mean_fwd(const void* h, const float* mask, void* out, int B, int T, int H, int dtype,
                          mmvqa_stream_t stream);
int mmvqa_masked_mean_bwd(const void* dout, const float* mask, void* dh, int B, int T, int H, int dtype,
                          mmvqa_stream_t stream);
/* row-wise L2 normalise (F.normalize(dim=1), mmbert.py:157), fp32 in/out */
int mmvqa_l2norm_fwd(const float* x, float* y, float* inv_norm, int rows, int cols, mmvqa_stream_t stream);
int mmvqa_l2norm_bwd(const float* y, const float* inv_norm, const float* dy, float* dx, int rows, int cols,
                     mmvqa_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Losses
 * ---------------------------------------------------------------------------------- */
/* ASLSingleLabel, models/asl_singlelabel.py:23-52.  One pass: per-row loss, d(loss_row)/d(logits)
 * (unscaled) and the smoothed one-hot the reference leaves in self.targets_classes (optional). */
int mmvqa_asl_fwd_bwd(const void* logits, int64_t ld, const int64_t* target, float* loss_rows, float* dlogits,
                      float* targets_classes, int B, int C, float gamma_pos, float gamma_neg, float eps, int dtype,
                      mmvqa_stream_t stream);
/* MLM loss: NLL(log_softmax(logits)) per row, pretrain/roco_utils.py:235-236; dlogits (dtype `dtype`)
 * = (softmax - onehot) * scale, written in place of / next to the logits */
int mmvqa_ce_fwd_bwd(const void* logits, int64_t ld, const int64_t* target, float* loss_rows, void* dlogits,
                     int64_t ld_d, int64_t rows, int C, float scale, int dtype, mmvqa_stream_t stream);
/* Chunked vocabulary cross entropy: the [rows, V] logits of the MLM head
 * (classifier[2], models/mmbert.py:137,154-155) and their log-softmax / gradient (pretrain/roco_utils.py:235-236) are
 * never materialised.  The caller produces the fp32 logits of columns [col0, col0 + Vc) with mmvqa_gemm into a
 * [rows, ld] scratch chunk; _stats folds the chunk into a running (row max, row sum-exp) and picks up the target logit
 * (loss_row = rowmax + log(rowsum) - tgt_logit after the last chunk); in the backward pass the chunk is recomputed and
 * _grad turns it into dlogits = (softmax - onehot) * row_scale[row] (storage type dl_dtype) for the dgrad / wgrad GEMMs. */
int mmvqa_ce_chunk_stats(const float* logits, int64_t ld, const int64_t* target, int64_t rows, int col0, int Vc,
                         float* rowmax, float* rowsum, float* tgt_logit, int first, mmvqa_stream_t stream);
int mmvqa_ce_chunk_grad(const float* logits, int64_t ld, const int64_t* target, int64_t rows, int col0, int Vc,
                        const float* rowmax, const float* rowsum, const float* row_scale, void* dlogits, int64_t ld_d,
                        int dl_dtype, mmvqa_stream_t stream);
/* SupCon row pass, models/SupConLoss/loss.py:72-96, over logits = anchor.contrast^T (NOT yet divided
 * by temperature) [R, N] fp32 for anchors row_offset..row_offset+R of the global N.  mask is the
 * un-tiled [bsz, bsz] float mask (NULL = SimCLR identity).  Writes per-anchor loss terms
 * -(T/Tb) * mean_log_prob_pos and G = d(sum of those)/d(logits) [R, N] fp32. */
int mmvqa_supcon_rows(const float* logits, const float* mask, float* loss_rows, float* G, int R, int N, int bsz,
                      int row_offset, float temperature, float base_temperature, mmvqa_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Whole RealFormer encoder forward in ONE launch (small batches; bf16 only).
 * replaces: RealFormer.forward, models/mmbert.py:103-108 = n_layers x ResEncoderBlock.forward,
 * models/realformer.py:30-51 (resmha + proj + dropout + ln1 + ff + dropout + ln2), i.e. the same arithmetic as the
 * per-operator chain mmvqa_rf_attn_fwd_fused / mmvqa_gemm / mmvqa_add_layernorm_fwd*, kept on chip by one 8-CTA
 * cluster per group of samples (mmvqa_b200/csrc/rf_encoder.cu).  Every intermediate the backward pass needs is
 * written to the stacked [n_layers, ...] buffers below, in the layouts of the per-operator entry points.
 * Per-layer parameter pointers are HOST arrays of DEVICE pointers (weights bf16 row-major as nn.Linear stores them,
 * biases / LayerNorm parameters fp32).  mmvqa_rf_encoder_fwd_supported() tells whether a shape can take this path (hidden 768,
 * 8 heads, ff 3072, T <= 32, <= 16 layers, few enough sample groups to be co-resident); other shapes use the
 * per-operator entry points. */
typedef struct mmvqa_rf_encoder_args {
  int B, T, hidden, heads, ff, n_layers;
  const void* const* wkqv;      /* [n_layers] -> bf16 [3*d, d]      realformer.py:13 */
  const void* const* wproj;     /* [n_layers] -> bf16 [hidden, hidden]  :14 */
  const void* const* w1;        /* [n_layers] -> bf16 [ff, hidden]  ff.0  :22 */
  const void* const* w2;        /* [n_layers] -> bf16 [hidden, ff]  ff.2  :25 */
  const float* const* b1; const float* const* b2;
  const float* const* ln1_w; const float* const* ln1_b; const float* const* ln2_w; const float* const* ln2_b;
  const void* x0;               /* bf16 [M, hidden]: encoder input (M = B*T) */
  void* xout;                   /* bf16 [n_layers, M, hidden]: output of layer l = input of layer l+1 */
  void* kqv;                    /* bf16 [n_layers, M*heads, 3*d] */
  float* scores;                /* fp32 [n_layers, B, heads, T, T] */
  void* att; void* y1; void* x1;        /* bf16 [n_layers, M, hidden] */
  void* hpre; void* hact;               /* bf16 [n_layers, M, ff] */
  void* y2;                             /* bf16 [n_layers, M, hidden] */
  float* mean1; float* rstd1; float* mean2; float* rstd2;   /* fp32 [n_layers, M] */
  const float* prev;            /* optional fp32 [B, heads, T, T]: scores entering layer 0 */
  const float* mask;            /* optional fp32 [B, T] */
  float dropout_p1, dropout_p2, eps;
  uint64_t dropout_seed;        /* layer l uses seed + 2l (proj branch) and seed + 2l + 1 (ff branch) */
  void* trace;                  /* optional int64 [4, 16, 16]: clock64 stamps of CTA 0 per role / layer / phase (tuning) */
} mmvqa_rf_encoder_args;
int mmvqa_rf_encoder_fwd_supported(int B, int T, int hidden, int heads, int ff, int n_layers);
int mmvqa_rf_encoder_fwd(const mmvqa_rf_encoder_args* args, mmvqa_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Backward of the attention block of ONE RealFormer layer in one cluster launch (small batches; bf16 only):
 * the backward of  x1 = ln1(x + dropout(proj(resmha(x))))  (models/realformer.py:30-45,49) = LN1 backward, proj dgrad,
 * residual-attention backward and kqv dgrad (+ residual), i.e. mmvqa_layernorm_bwd_parts | mmvqa_gemm | mmvqa_rf_attn_bwd |
 * mmvqa_gemm of the per-operator chain (mmvqa_b200/csrc/rf_attn_block.cu).
 *   in : dy_parts [nparts, M, hidden] fp32 (+ dy_res [M, hidden], optional) = gradient w.r.t. x1; y1, mean1, rstd1, ln1_w of
 *        the forward LN1; wproj [hidden, hidden], wkqv [3d, d] (bf16 operand copies); kqv [M*heads, 3d], scores
 *        [B, heads, T, T] saved by the forward pass; dscores_in = score gradient arriving from layer l + 1 (optional).
 *   out: dpr [M, hidden] = dropout(d y1) (operand of the proj weight-gradient GEMM), dkqv [M*heads, 3d] (operand of the kqv
 *        weight-gradient GEMM), dprev (score gradient for layer l - 1, optional), dxin [M, hidden] = gradient w.r.t. the
 *        layer input (before the FF block of layer l - 1), dln1_w / dln1_b ACCUMULATED into zero-filled fp32 buffers. */
typedef struct mmvqa_rf_attn_block_bwd_args {
  int B, T, hidden, heads;
  const float* dy_parts; int nparts; int64_t part_stride;
  const void* dy_res;
  const void* y1; const float* mean1; const float* rstd1; const float* ln1_w;
  const void* wproj; const void* wkqv;
  const void* kqv; const float* scores; const float* dscores_in;
  void* dpr; void* dkqv; float* dprev; void* dxin; float* dln1_w; float* dln1_b;
  float dropout_p; uint64_t dropout_seed;      /* the forward's proj-branch dropout (seed + 2l) */
} mmvqa_rf_attn_block_bwd_args;
int mmvqa_rf_attn_block_bwd_supported(int B, int T, int hidden, int heads);
int mmvqa_rf_attn_block_bwd(const mmvqa_rf_attn_block_bwd_args* args, mmvqa_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Visual-token projector, forward with the weight-gradient contraction inside (models/image_encoding.py:74-87, 103-113):
 *   vis[b, m]      += mean_hw act(sum_c W[m, c] f[b, c, hw])
 *   pgrad[b, m, c] += sum_hw act'(.)[b, m, hw] f[b, c, hw]        (fp32 [B, M, C], caller zero-fills both)
 * W bf16 [M, C] (ldw), f bf16 [B * C, ldf] = the NCHW map read in place.  The backward pass is then
 *   dW[m, c] = scale * sum_b dv[b, m] pgrad[b, m, c]   with scale = 1 / HW  (mmvqa_vistok_dw)
 * and the [B, M, HW] act' map of the EPI_ACT_ROWSUM path (aux_out) is never written: use it when the feature maps need no
 * gradient.  C <= 128, C % 8 == 0, HW >= 128 (mmvqa_vistok_pgrad_supported); the other levels keep mmvqa_gemm.
 * ---------------------------------------------------------------------------------- */
int mmvqa_vistok_pgrad_supported(int M, int HW, int C);
int mmvqa_vistok_fwd_pgrad(const void* W, int64_t ldw, const void* f, int64_t ldf, float* vis, float* pgrad, int M, int HW,
                           int C, int B, int act, mmvqa_stream_t stream);
int mmvqa_vistok_dw(const float* pgrad, const float* dv, float scale, float* dW, int B, int M, int C, mmvqa_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Optimiser: multi-tensor Adam, torch.optim.Adam semantics
 * (vqamed2019/train.py:160, no amsgrad, L2 weight decay).  `table` is a DEVICE array of n_chunks
 * descriptors, each a contiguous chunk (8192 elements is a good size) of one parameter
 * tensor; one CTA per chunk.  If bf16_out != NULL the updated parameter is also written as
 * bf16 (refreshes the tensor-core weight cache in the same pass).
 * ---------------------------------------------------------------------------------- */
typedef struct mmvqa_adam_desc {
  float* p; float* m; float* v; const void* g; void* bf16_out; int64_t n;
  int64_t flags;             /* bit 0: g is bf16 (all-reduced bf16 gradient bucket) instead of fp32 */
  const unsigned char* row_live;   /* optional row gate (embedding tables, weight_decay == 0 only): the chunk is n / row_len
                                whole rows and row r is skipped while row_live[r] == 0.  A row whose gradient has been zero
                                in every step so far has m = v = 0, so its Adam update is exactly the identity: skipping it
                                changes no bit of the result and saves 28 B/parameter of HBM traffic on the 23.4 M-parameter
                                word-embedding table, of which a step touches <= B*T rows (models/mmbert.py:52-63) */
  int64_t row_len;           /* elements per row when row_live != NULL (multiple of 4, rows 16-byte aligned) */
} mmvqa_adam_desc;
/* row_live[ids[i]] = 1 for i < n (ids outside [0, rows) are ignored): the rows that receive gradient in this step */
int mmvqa_mark_rows(unsigned char* row_live, const int64_t* ids, int64_t n, int64_t rows, mmvqa_stream_t stream);
/* `step` (1-based) sets the bias corrections; if step_dev != NULL the kernel reads the step from
 * that device int instead, so a captured CUDA graph can be replayed while the host bumps it.
 * max_ctas > 0 caps the grid (the CTAs stride over the table): an update that runs underneath the backward pass
 * then takes a bounded share of the HBM bandwidth. */
int mmvqa_adam_step(const mmvqa_adam_desc* table, int n_chunks, float lr, float beta1, float beta2, float eps,
                    float weight_decay, int step, const int* step_dev, float grad_scale, int max_ctas, int background,
                    mmvqa_stream_t stream);
/* Same update with the two values a training loop changes between steps read from DEVICE memory:
 * hyper_dev[0] = lr (ReduceLROnPlateau, vqamed2019/train.py:161,233; pretrain/roco_train.py:91,162),
 * hyper_dev[1] = grad_scale.  A captured graph of this launch follows scheduler.step() as long as the caller refreshes
 * the two floats before the replay.
 * background != 0 (both entry points): the update runs underneath other work (a layer's update under the rest of the
 * backward pass) -- one 16-byte group in flight per thread and array, so it streams gently; 0: on the critical path,
 * two groups in flight (pair it with ~8192-element chunks). */
int mmvqa_adam_step_dev(const mmvqa_adam_desc* table, int n_chunks, const float* hyper_dev, float beta1, float beta2,
                        float eps, float weight_decay, const int* step_dev, int max_ctas, int background,
                        mmvqa_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Data-parallel gradient exchange (collective 1; the reference has no distributed code -- this is
 * what DDP's all-reduce would do around vqamed2019/train.py:173-174).  In-place SUM all-reduce of a bucket that lives in
 * symmetric memory bound to an NVLink multicast object: multimem.ld_reduce / multimem.st through the NVSwitch, two
 * shots, a few CTAs.  multicast_ptr = multicast address of the bucket (same offset on every rank), signal_pads_dev =
 * DEVICE array of `world` pointers to the ranks' signal pads (uint32 words, zero when idle; ctas * world words used).
 * Every rank must launch it with the same arguments in the same order (a collective).  dtype F32 or BF16 (fp32
 * accumulation in the switch). */
int mmvqa_multimem_allreduce(void* multicast_ptr, const void* signal_pads_dev, int rank, int world, int64_t nbytes, int dtype,
                             int max_ctas, mmvqa_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MMVQA_H_ */
