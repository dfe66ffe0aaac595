/* libb4cp — C ABI of the B200-native clickstream-transformer hot path.
 *
 * The reference (MiladShahidi/BERT4ClickPath) has no FFI layer: its hot path is a chain of
 * TensorFlow 2.3.1 ops called from Python classes.  Each export below replaces the TF op chain
 * at the cited reference call site (paths relative to the reference repository) with one or a
 * few hand-written sm_100a kernels.  INTEGRATION.md shows the ctypes binding a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the name says `h_` (host);
 *   - `stream` is a cudaStream_t passed as void*; calls are stream-ordered, never synchronise,
 *     never allocate (callers pass workspaces sized by the *_workspace_bytes queries);
 *   - return value: 0 = ok, <0 = bad argument, >0 = cudaError_t; b4cp_last_error() gives the
 *     thread-local message;
 *   - matrices are row-major; Dense kernels use the Keras (in, out) layout;
 *   - "bf16" buffers are passed as void* (raw __nv_bfloat16 bits).
 */
#ifndef B4CP_H_
#define B4CP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B4CP_VERSION 1
#define B4CP_MAX_FEATURES 8

const char* b4cp_last_error(void);
int b4cp_version(void);
/* number of CUDA kernels launched by this library since load (all threads) */
long b4cp_launch_count(void);
/* fails (<0) unless the current device is compute capability 10.x */
int b4cp_device_check(void);

/* ------------------------------------------------------------------ dense layers (tcgen05)
 * Replaces tf.keras.layers.Dense MatMul+BiasAdd(+ReLU) and its autodiff transposes:
 *   clickstream_transformer/transformer.py:112-116,139-141,158 (MHA projections)
 *   clickstream_transformer/transformer.py:163-167 (feed-forward)
 *   clickstream_transformer/head.py:10-11,16-19,35-45,56-62 (head MLPs)
 *
 *   C[M,N] = epilogue( alpha * sum_k A(m,k) * B(n,k) ),  bf16 operands, fp32 accumulation.
 * a_mn = 0: A is stored [M][K] (K contiguous);  a_mn = 1: A is stored [K][M] (M contiguous).
 * b_mn = 0: B is stored [N][K];                  b_mn = 1: B is stored [K][N] (Keras kernel).
 * Leading dimensions in elements, multiples of 8.  splits > 1 writes `splits` raw fp32 partial
 * products at out_f32 + z*split_stride (reduce them with b4cp_reduce_splits).
 */
typedef struct {
  float alpha;         /* scale on the accumulator (1.0f for a plain product) */
  const float* bias;   /* [N] added to every row, or NULL */
  int relu;            /* max(x, 0) after bias */
  const void* gate;    /* bf16 [M][ld_gate]: result zeroed where gate <= 0 (ReLU backward), or NULL */
  long ld_gate;
  const float* addend; /* fp32 [M][ld_addend] added last (residual / gradient accumulate), or NULL */
  long ld_addend;
  float* out_f32;      /* fp32 [M][ld_f32] or NULL */
  long ld_f32;
  long split_stride;   /* elements between split-K partials in out_f32 */
  void* out_bf16;      /* bf16 [M][ld_bf16] or NULL */
  long ld_bf16;
} b4cp_gemm_epilogue;

int b4cp_gemm_bf16(const void* A, int a_mn, long lda, const void* B, int b_mn, long ldb, int M,
                   int N, int K, int splits, const b4cp_gemm_epilogue* ep, void* stream);
/* number of K splits that fills the GPU for an M x N output */
int b4cp_gemm_splits_for(int M, int N, int K);

/* sums `splits` partial matrices (n elements each, split_stride apart) in fixed order */
int b4cp_reduce_splits(const float* partials, int splits, long n, long split_stride, float* out,
                       void* stream);
/* same, over a dense [M][N] matrix, with an optional ReLU-backward gate (bf16, zero where
 * gate <= 0) and fp32 (row stride ld_f32 >= N) and/or bf16 (row stride ld_bf16) outputs */
int b4cp_reduce_splits_ex(const float* partials, int splits, long M, int N, long split_stride,
                          const void* gate, long ld_gate, float* out_f32, long ld_f32,
                          void* out_bf16, long ld_bf16, void* stream);
/* fp32 [rows][ld_in] -> bf16 [rows][ld_out], columns >= cols are zero-filled */
int b4cp_cast_f32_bf16(const float* in, long rows, int cols, long ld_in, void* out, long ld_out,
                       void* stream);

/* ------------------------------------------------------------------ fp32-class ("parity") mode
 * The reference's Dense / attention / softmax arithmetic is fp32 (TensorFlow 2.3.1 kernels behind
 * clickstream_transformer/transformer.py:64-97, :112-116, :139-167 and head.py:35-45).  These
 * entry points let the whole path run at fp32-class accuracy (<= 1e-3 of the reference, measured
 * ~1e-5) on the same tcgen05 GEMM: a fp32 operand is split into bf16 hi + lo parts laid out along
 * the contraction axis, A as (hi | hi | lo) [order 0] and B as (hi | lo | hi) [order 1], so ONE
 * b4cp_gemm_bf16 over K' = 3K accumulates hi*hi + hi*lo + lo*hi in fp32.
 *   k_along_rows = 0: in [rows][ld_in] with `cols` = K valid columns -> out bf16 [rows][3*ld8(K)]
 *                     (ld_out must equal 3*ld8(K); pad columns are zero);
 *   k_along_rows = 1: in [K = rows][ld_in] with `cols` valid columns -> out bf16
 *                     [3*ld8(K)][ld_out], ld_out a multiple of 8 >= cols (pad rows / columns zero).
 * Either way the contraction length of the product becomes K' = 3*ld8(K).
 */
int b4cp_split_bf16x3(const float* in, long rows, int cols, long ld_in, void* out, long ld_out,
                      int k_along_rows, int order, void* stream);
/* fp32 masked self-attention (same contract as b4cp_attention_fwd/_bwd with fp32 buffers; exact
 * division by sqrt(dh), expf; S <= 256 forward, backward while 4 head tiles fit shared memory) */
int b4cp_attention_f32_fwd(const float* qkv, const int32_t* ids_first, int B, int S, int H, int dh,
                           float* out, float* lse, void* stream);
int b4cp_attention_f32_bwd(const float* qkv, const float* dout, const float* lse,
                           const int32_t* ids_first, int B, int S, int H, int dh, float* dqkv,
                           void* stream);
/* column sums of a fp32 [T][ld] matrix (workspace: b4cp_colsum_workspace_bytes) */
int b4cp_colsum_f32(const float* in, long T, int n, long ld, float* out, void* workspace,
                    void* stream);

/* ------------------------------------------------------------------ input embedding
 * Replaces Embedding gather x F, tf.concat, * sqrt(d_model), + pos_encoding[:, :S] and the
 * encoder's input Dropout: clickstream_transformer/transformer.py:376-398, :263.
 *   out[b,s,off_f+j] = fl32(fl32(E_f[ids_f[b,s], j] * fl32(sqrt(d_model))) + PE[s, off_f+j])
 * h_ids / h_tables / h_dims / h_rows are HOST arrays of length F (device pointers inside).
 * `pe` is the fp32 [>=S][d_model] sinusoid table.  dropout_rate = 0 disables dropout; otherwise
 * kept values are scaled by 1/(1-rate) and the keep bit of element i is the one
 * b4cp_dropout_mask(seed, site) reports.  Bit-exact against the oracle when dropout is off.
 * Every `seed` argument of this library is either the seed value (< 2^63) or
 * B4CP_SEED_FROM_DEVICE(ptr): the seed is the int32 at device address `ptr` when the kernel runs,
 * so a captured CUDA graph draws new masks on every replay.
 */
#define B4CP_SEED_FROM_DEVICE(ptr) ((uint64_t)1 << 63 | (uint64_t)(uintptr_t)(ptr))
int b4cp_embed_fwd(const int32_t* const* h_ids, const float* const* h_tables, const int* h_dims,
                   const int* h_rows, int F, const float* pe, int B, int S, float dropout_rate,
                   uint64_t seed, uint32_t site, float* out_f32, void* out_bf16, void* stream);

/* Backward of the gather for ONE feature (TF autodiff IndexedSlices + UnsortedSegmentSum,
 * implied by transformer.py:347-355): table_grad[r, :] = sqrt(d_model) * sum over tokens with
 * id r of dout[token, col_offset : col_offset+dim] (after the input-dropout mask).  Deterministic:
 * stable radix sort of (id, token), one segment per unique id, fixed summation order.
 * table_grad (rows x dim, fp32) is zero-filled by the call.  uniq_ids / n_unique are optional.
 */
long b4cp_embed_bwd_workspace_bytes(long tokens, int max_dim);
int b4cp_embed_bwd(const float* dout, int d_model, int col_offset, int dim, const int32_t* ids,
                   long tokens, int rows, float dropout_rate, uint64_t seed, uint32_t site,
                   float* table_grad, int32_t* uniq_ids, int32_t* n_unique, void* workspace,
                   long workspace_bytes, void* stream);

/* The same backward in two stream-ordered parts: the sort of (id, token) depends on the ids alone,
 * so it can run EARLY (e.g. on a side stream while the backward is still in the encoder layers);
 * b4cp_embed_bwd_sorted then only streams the gradient rows through the segment sums.  Both calls
 * take the same workspace (b4cp_embed_bwd_workspace_bytes(tokens, max_dim), max_dim >= dim). */
int b4cp_embed_sort(const int32_t* ids, long tokens, int rows, int max_dim, void* workspace,
                    long workspace_bytes, void* stream);
int b4cp_embed_bwd_sorted(const float* dout, int d_model, int col_offset, int dim, long tokens,
                          int rows, float dropout_rate, uint64_t seed, uint32_t site,
                          float* table_grad, void* workspace, long workspace_bytes, void* stream);

/* ------------------------------------------------------------------ encoder layer pieces
 * Fused short-sequence masked self-attention (S <= 256), one CTA per (sequence, head):
 * clickstream_transformer/transformer.py:64-97 (scaled_dot_product_attention, additive -1e9 key
 * padding mask from create_padding_mask :38-41) and :130-156 (split/merge heads).
 * qkv: bf16 [B*S][3*d] = (q | k | v) with head h in columns h*dh..; out: bf16 [B*S][d];
 * lse: fp32 [B][H][S] log-sum-exp of each score row (saved for the backward).
 */
int b4cp_attention_fwd(const void* qkv, const int32_t* ids_first, int B, int S, int H, int dh,
                       void* out, float* lse, void* stream);
/* backward: `out` is the forward output (bf16 [B*S][d]); the tensor-core path for 128 < S <= 256
 * takes delta = rowsum(dO o O) from it (may be NULL: the SIMT kernel is used for such S then). */
int b4cp_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse,
                       const int32_t* ids_first, int B, int S, int H, int dh, void* dqkv,
                       void* stream);

/* y = LayerNormalization(eps=1e-6)(x + Dropout(r)): transformer.py:204-206, :209-211.
 * Backward returns dx (residual branch, fp32), dr (gradient of the Dense output r, bf16) and the
 * reduced dgamma, dbeta and dbias = column sums of dr.  d <= 256.  dr_f32 (fp32 [T][d], optional)
 * receives the same dr unrounded (fp32-class mode).
 */
int b4cp_residual_ln_fwd(const float* x, const float* r, long T, int d, const float* gamma,
                         const float* beta, float dropout_rate, uint64_t seed, uint32_t site,
                         float* y_f32, void* y_bf16, long ld_bf16, void* stream);
long b4cp_residual_ln_bwd_workspace_bytes(int d);
int b4cp_residual_ln_bwd(const float* dy, const float* x, const float* r, long T, int d,
                         const float* gamma, float dropout_rate, uint64_t seed, uint32_t site,
                         float* dx, void* dr_bf16, long ld_bf16, float* dr_f32, float* dgamma,
                         float* dbeta, float* dbias, void* workspace, void* stream);

/* bias gradients: out[c] = sum over rows of a bf16 [T][ld] matrix (deterministic two-stage) */
long b4cp_colsum_workspace_bytes(long T, int n);
int b4cp_colsum_bf16(const void* in, long T, int n, long ld, float* out, void* workspace,
                     void* stream);

/* fills out[i] with 1/(1-rate) (kept) or 0 (dropped) for the n elements of a dropout site */
int b4cp_dropout_mask(float* out, long n, float dropout_rate, uint64_t seed, uint32_t site,
                      void* stream);

/* x[i] <- x[i] / (1-rate) (kept) or 0 (dropped), in place, with the bits of the same site: the
 * gradient side of a dropout site when the consumer cannot recompute the mask itself (the
 * data-parallel row exchange of table gradients gathers rows of OTHER ranks' tokens) */
int b4cp_dropout_apply(float* x, long n, float dropout_rate, uint64_t seed, uint32_t site,
                       void* stream);

/* ------------------------------------------------------------------ output selection
 * clickstream_transformer/clickstream_transformer.py:260-297 (_gather_output_by_raw_value):
 * token indices whose first-feature id == value, in (b, s) order.  row_index[i] = -1 for
 * count <= i < capacity.  gather_rows writes zero vectors for -1 (the reference's zero padding).
 */
long b4cp_select_workspace_bytes(long tokens);
int b4cp_select_masked(const int32_t* ids_first, long tokens, int value, int32_t* row_index,
                       long capacity, int32_t* count, void* workspace, void* stream);
/* reference label matrix (B, max_n_masked) float32 padded with -1.0 (input_pipeline.py:95-97,
 * :198-214) -> int32 labels of the valid entries in row-major order (= [MASK] order) */
int b4cp_compact_labels(const float* labels, long n, float label_pad, int32_t* out, long capacity,
                        int32_t* count, void* workspace, void* stream);
int b4cp_gather_rows(const float* x, int d, const int32_t* row_index, long M, float* out_f32,
                     void* out_bf16, long ld_bf16, void* stream);
int b4cp_scatter_rows(const float* src, int d, const int32_t* row_index, long M, float* dst,
                      void* stream);

/* ------------------------------------------------------------------ Cloze batch builder
 * On-device restatement of examples/BERT4Rec/source/input_pipeline.py:21-32, :59-133, :198-214
 * and the chaining of clickstream_transformer.py:38-63 over sessions resident in HBM as a CSR
 * (items int32 input-vocabulary ids, offsets int64 [n_sessions + 1]); batch row b is session
 * session_idx[b].  train != 0: the last item is dropped and n = clip(int(float32(len) *
 * float32(masked_percentage)), 0, max_masked) distinct positions (float32 product, truncated, as
 * input_pipeline.py:68-70 computes it) are replaced by mask_id; else only the last position.
 * Positions = the n smallest keys splitmix64(splitmix64(seed + session) + pos), ties by position
 * (the reference's tf.random.shuffle stream cannot be matched; oracle/ restates this rule).
 * Writes ids [B][L + 3] = cls sep items.. pad.. sep, labels [B][Mmax] = (id - label_offset) of the
 * masked items in ascending position, padded with label_pad; ADDS the number of masked items to
 * *n_masked (zero it first) and sets *status = b + 1 if row b does not fit (len > L, n > Mmax or an
 * empty session; that row is written as pads).  L <= 2048. */
int b4cp_cloze_build(const int32_t* items, const long long* offsets, const int32_t* session_idx,
                     int B, int L, int Mmax, int train, double masked_percentage, int max_masked,
                     unsigned long long seed, int cls_id, int sep_id, int mask_id, int pad_id,
                     int label_offset, float label_pad, int32_t* ids, float* labels,
                     int32_t* n_masked, int32_t* status, void* stream);
/* host-side: the key above for one (seed, session, pos); needs no device */
unsigned long long b4cp_cloze_position_key(unsigned long long seed, unsigned long long session,
                                           unsigned long long pos);

/* ------------------------------------------------------------------ host: vocabulary lookup
 * tf.lookup.StaticVocabularyTable(KeyValueTensorInitializer(keys, range(len(keys))),
 * num_oov_buckets=1) as clickstream_transformer/clickstream_transformer.py:247-258 builds it, on
 * the host (the reference runs it inside its graph on string tensors): key j -> j (first
 * occurrence of a duplicate), any other string -> n_keys.  Strings are fixed-width UCS4 code
 * points, NUL padded (NumPy's '<U<width>' layout).  No device is needed.  lookup fills
 * out_ids[n_tokens] with up to n_threads host threads. */
void* b4cp_vocab_table_create(const uint32_t* keys_ucs4, long n_keys, int width);
void b4cp_vocab_table_destroy(void* table);
long b4cp_vocab_table_size(const void* table);
int b4cp_vocab_table_lookup(const void* table, const uint32_t* tokens_ucs4, long n_tokens, int width,
                            int32_t* out_ids, int n_threads);

/* ------------------------------------------------------------------ Cloze loss, materialised
 * Small-vocabulary path of SoftMaxHead + ClozeMaskedLoss (head.py:38-47;
 * examples/BERT4Rec/source/utils.py:56-134; losses.py:31-98) in logits mode.
 * labels: int32 [M], -1 = padded row.  loss_stats[0] = sum of (lse - z_t) over valid rows,
 * loss_stats[1] = number of valid rows (the loss is their ratio, 0 when there are none).
 */
int b4cp_ce_rows_stats(const float* logits, long ld, long M, int V, const int32_t* labels,
                       float* lse, float* tgt, void* stream);
int b4cp_ce_loss_reduce(const float* lse, const float* tgt, const int32_t* labels, long M,
                        float* loss_stats, void* stream);
/* data-parallel form: loss_stats[0] = this rank's loss sum, loss_stats[1] = (float)*n_global, the
 * number of valid rows over ALL ranks (an int32 on the device, all-reduced while the forward runs:
 * the masked mean of losses.py:80-91 is global, and the backward needs only the count) */
int b4cp_ce_loss_reduce_n(const float* lse, const float* tgt, const int32_t* labels, long M,
                          const int32_t* n_global, float* loss_stats, void* stream);
/* dz = (softmax - onehot) / loss_stats[1] (bf16, optional) and/or probabilities (fp32, optional) */
int b4cp_ce_rows_grad(const float* logits, long ld, long M, int V, const int32_t* labels,
                      const float* lse, const float* loss_stats, void* dz_bf16, long ld_dz,
                      float* probs, long ld_probs, void* stream);

/* fp32-class mode: dz = (softmax - onehot) / loss_stats[1] written IN PLACE over the fp32 logits
 * [M][ld] (columns V..ld-1 zeroed), rows with label -1 -> 0 */
int b4cp_ce_rows_grad_f32(float* logits, long ld, long M, int V, const int32_t* labels,
                          const float* lse, const float* loss_stats, void* stream);

/* ------------------------------------------------------------------ Cloze loss, FUSED (tcgen05)
 * SoftMaxHead's Dense(V) + softmax + sparse categorical cross-entropy and its gradient without
 * ever writing logits / probabilities / dlogits to HBM (head.py:36,45;
 * examples/BERT4Rec/source/utils.py:116-134; losses.py:31-98, logits mode).
 *   x_bf16: bf16 [M][ldx] head hidden states; w_bf16: bf16 [h][ldw] Keras (in, out) kernel;
 *   bias: fp32 [V]; labels: int32 [M], -1 = padded row.
 * fwd: lse[M] = log sum_v exp(x W + b), tgt[M] = logit of the label (0 for padded rows); feed
 *      them to b4cp_ce_loss_reduce.  h in {64,128,192,256}.  With want_dx (h = 128) the kernel
 *      also accumulates, flash-attention style, U = sum_v exp(z_v - max) W[:, v] per row with a
 *      second tensor-core product per tile; the partials stay in `workspace`.
 * dx : d(loss)/dx[M][h] = gate * (U / sum - W[:, label]) / n from the forward's workspace, with
 *      n = loss_stats[1] (the global valid count); gate (bf16 [M][ld_gate], zero where <= 0) is
 *      the ReLU output that produced x, or NULL.  Deterministic, no atomics.
 * bwd: dW[h][V] = X^T dZ (fp32, accumulated in TMEM in a fixed order), db[V] = column sums of
 *      dZ, with dZ = (softmax - onehot)/n on valid rows recomputed tile by tile.  h must be 128.
 */
long b4cp_vocab_ce_workspace_bytes(long M, int V, int h);
/* host-only: how the forward of (M, V, h) is scheduled.  out[0] = 0: grid of (row tile x vocabulary
 * chunk) CTAs sweeping W in lock step (W larger than L2), 1: <= 148 persistent CTAs over contiguous
 * ranges of the (row tile, vocabulary tile) space (W fits L2); out[1] = CTAs; out[2] = partial
 * (max, sum, U) slots per row in the workspace; out[3] = tiles per range / per chunk. */
int b4cp_vocab_ce_plan(long M, int V, int h, long* out);
int b4cp_vocab_ce_fwd(const void* x_bf16, long ldx, long M, int h, const void* w_bf16, long ldw,
                      const float* bias, int V, const int32_t* labels, int want_dx, float* lse,
                      float* tgt, void* workspace, void* stream);
/* dx with lse_global != NULL is the vocabulary-parallel form: the shard's partial contribution
 * gate * (U * exp(max - lse_global) - [label < V] W[:, label]) / n, to be summed over shards. */
int b4cp_vocab_ce_dx(long M, int h, int V, const int32_t* labels, const float* loss_stats,
                     const float* lse_global, const void* w_bf16, long ldw, const void* gate_bf16,
                     long ld_gate,
                     float* out_f32, void* out_bf16, long ld_bf16, const void* workspace,
                     void* stream);
int b4cp_vocab_ce_bwd(const void* x_bf16, long ldx, long M, int h, const void* w_bf16, long ldw,
                      const float* bias, int V, const int32_t* labels, const float* lse,
                      const float* loss_stats, float* dW, float* db, void* stream);

/* Vocabulary-parallel helpers (the output kernel sharded by contiguous vocabulary ranges across
 * GPUs; the reference has no such mode - its MirroredStrategy replicates head.py:36's kernel):
 * remap labels for a shard (-1 stays, owned -> local id, others -> a valid id >= v_count), and
 * merge per-shard log-sum-exps: out[m] = log sum_r exp(parts[r][m]). */
int b4cp_shard_labels(const int32_t* labels, long M, int v_begin, int v_count, int32_t* out,
                      void* stream);
int b4cp_lse_merge(const float* parts, int n_parts, long M, float* out, void* stream);

/* Probability inputs (a materialised SoftMaxHead output): z = log(clip(p, lo, hi)) reproduces
 * K.sparse_categorical_crossentropy(from_logits=False) of TF 2.3 when followed by
 * b4cp_ce_rows_stats (losses.py:60 via examples/BERT4Rec/source/main.py:89). */
int b4cp_clip_log(const float* p, float* out, long n, float lo, float hi, void* stream);
/* MaskedLoss with K.binary_crossentropy and optional pos_weight (losses.py:31-98):
 * stats[0] = sum of weighted item losses over labels != label_pad, stats[1] = their count */
int b4cp_masked_bce(const float* y_true, const float* probs, long n, float label_pad,
                    float pos_weight, int use_pos_weight, float* stats, void* stream);
/* Accumulators of clickstream_transformer/metrics.py: counters[0..5] += (sum mask, sum y*mask,
 * sum round(p)*mask, tp, condition_true, predicted_true) with mask = (y != label_pad), tf.round
 * (half to even) as the 0.5 threshold, and F1Score's unmasked int32 comparisons (metrics.py:63-78). */
int b4cp_binary_metric_counts(const float* y_true, const float* probs, long n, float label_pad,
                              float* counters, void* stream);
/* Backward of BinaryClassificationHead's Dense(1, sigmoid) + MaskedLoss(binary_crossentropy,
 * pos_weight) (head.py:11,24-26; losses.py:31-98): y_true / probs [M] (one logit per item),
 * stats = b4cp_masked_bce's (sum, n) (n may already be the global count).  ab: bf16 [M][ld_ab]
 * input of the Dense(1) (ReLU output of the head MLP -> gated = 1, or the encoder rows), w_out:
 * fp32 [h] its kernel.  Writes dz [M], d(loss)/d(ab) as fp32 [M][h] and/or bf16 [M][ld_dab],
 * dw [h], db [1].  Deterministic. */
int b4cp_binary_head_bwd(const float* y_true, const float* probs, long M, float label_pad,
                         float pos_weight, int use_pos_weight, const float* stats,
                         const void* ab_bf16, long ld_ab, int h, const float* w_out, int gated,
                         float* dz, float* dab_f32, void* dab_bf16, long ld_dab, float* dw,
                         float* db, void* stream);

/* The same item-wise gradient over a (rows, cols) sigmoid output — MultiLabel_MultiClass_
 * classification (head.py:50-69) under MaskedLoss(binary_crossentropy, pos_weight): every
 * (row, class) cell whose label != label_pad is one item of the masked mean (losses.py:50-98).
 * y_true / probs fp32 [rows][cols]; writes dz as fp32 [rows][cols] and/or bf16 [rows][ld_bf16]
 * (pad columns zeroed) — the operand of the dW = ab^T dz and dx = dz W^T GEMMs. */
int b4cp_sigmoid_bce_dz(const float* y_true, const float* probs, long rows, int cols,
                        float label_pad, float pos_weight, int use_pos_weight, const float* stats,
                        float* dz_f32, void* dz_bf16, long ld_bf16, void* stream);

/* sigmoid output activation of BinaryClassificationHead / MultiLabel_MultiClass_classification
 * (head.py:11, :57) */
int b4cp_sigmoid(const float* z, float* out, long n, void* stream);

/* ------------------------------------------------------------------ ranking metrics
 * tf.math.top_k order (score desc, ties -> lower id): utils.py:176, :245.  k <= 256.
 * rank_metrics accumulates counters += (hits, sum 1/log2(rank+2), n_valid): utils.py:176-187,
 * :211, :225-252.
 */
int b4cp_topk_rows(const float* scores, long ld, long rows, int V, int k, int32_t* out_ids,
                   float* out_scores, long ld_out, void* stream);
/* same ranking over explicit candidate lists: row r ranks cand_scores[r*ld .. r*ld+n_cand) with
 * ids cand_ids[...] (negative = empty slot); ids < V */
int b4cp_topk_candidates(const float* cand_scores, const int32_t* cand_ids, long ld, long rows,
                         int n_cand, int V, int k, int32_t* out_ids, float* out_scores,
                         long ld_out, void* stream);
/* counted lists: row r holds base_count + extra_count[r] entries (capped at n_cand); later slots
 * are never read */
int b4cp_topk_candidates_counted(const float* cand_scores, const int32_t* cand_ids, long ld,
                                 long rows, int n_cand, const int* extra_count, int base_count, int V,
                                 int k, int32_t* out_ids, float* out_scores, long ld_out,
                                 void* stream);
/* b4cp_topk_candidates restricted to the rows whose out_ids[row][0] == -2 (redo marker) */
int b4cp_topk_candidates_redo(const float* cand_scores, const int32_t* cand_ids, long ld, long rows,
                              int n_cand, int V, int k, int32_t* out_ids, float* out_scores,
                              long ld_out, void* stream);
/* FUSED inference scoring + top-k: ranks x W + b over the whole vocabulary without writing the
 * (M x V) scores (head.py:36,45 + examples/BERT4Rec/source/utils.py:176,:245).  x_bf16: bf16
 * [M][ldx]; w_bf16: bf16 [h][ldw] Keras kernel; h in {64,128,256}; k <= 104.  Exact, ties -> lower
 * id.  V < 262144: per-row heaps in shared memory; longer vocabularies: the first 65536 entries
 * are ranked to give every row a threshold, the tcgen05 product over the rest appends the scores
 * above it to per-row lists, and an exact merge finishes (rows whose list overflows are redone by
 * the heap kernel). */
long b4cp_score_topk_workspace_bytes(long M, int V, int k);
/* id_base is added to every reported id and V_total (0 = V) bounds the reported ids: a vocabulary
 * shard [id_base, id_base + V) of a V_total-wide output layer reports global ids */
int b4cp_score_topk(const void* x_bf16, long ldx, long M, int h, const void* w_bf16, long ldw,
                    const float* bias, int V, int k, int id_base, int V_total, int32_t* out_ids,
                    float* out_scores, long ld_out, void* workspace, void* stream);
int b4cp_rank_metrics(const int32_t* topk_ids, long M, int k, long ld, const int32_t* labels,
                      float* counters, void* stream);

/* ------------------------------------------------------------------ optimizer
 * tf.keras.optimizers.Adam(lr, 0.9, 0.999, epsilon=1e-9) (examples/BERT4Rec/source/main.py:87):
 *   lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v updates; theta -= lr_t * m / (sqrt(v) + eps).
 * t comes from *step_dev when given (graph-friendly) else step_host (1-based).  grad is scaled by
 * grad_scale first.  Optionally refreshes a bf16 shadow [n/cols][ld_shadow] of the parameter.
 */
int b4cp_adam_step(float* theta, const float* grad, float* m, float* v, long n, float lr,
                   float beta1, float beta2, float eps, const int* step_dev, int step_host,
                   float grad_scale, void* shadow_bf16, int cols, long ld_shadow, void* stream);
int b4cp_step_increment(int* step_dev, void* stream);
/* The same update as ONE sweep over flat parameter / gradient / moment buffers of n elements that
 * hold many parameters back to back (gaps must hold zeros in all four buffers).  h_segs (HOST
 * array, sorted by `begin`, disjoint, `begin` a multiple of 4): the Dense kernels whose bf16
 * shadow [numel / cols][ld_shadow] is refreshed in the same pass. */
#define B4CP_ADAM_MAX_SEGS 48
typedef struct {
  long begin;         /* first element of the kernel in the flat buffers */
  long numel;         /* rows * cols */
  int cols;
  long ld_shadow;
  void* shadow_bf16;
} b4cp_adam_segment;
int b4cp_adam_flat(float* theta, const float* grad, float* m, float* v, long n, float lr,
                   float beta1, float beta2, float eps, const int* step_dev, int step_host,
                   float grad_scale, const b4cp_adam_segment* h_segs, int n_segs, void* stream);
/* stream-ordered zero fill (cudaMemsetAsync) */
int b4cp_zero(void* ptr, long bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B4CP_H_ */
