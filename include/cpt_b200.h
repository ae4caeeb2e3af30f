/* cpt_b200 C ABI — the B200-native implementation of CPT's cross-modal BERT hot path.
 *
 * The reference (thunlp/CPT) has NO FFI for this path: it is a stack of Python nn.Modules
 * (SURVEY.md 8b).  Each entry point below therefore cites the reference *Python* interface it replaces;
 * cpt_b200/modeling_*.py binds them with ctypes behind drop-in modules of the same names
 * (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - plain C: pointers + sizes only, no torch / C++ types.  All data pointers are DEVICE pointers on the
 *     handle's device unless noted "host".  `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - every function returns 0 on success; on failure a non-zero code, and cpt_last_error() (thread-local)
 *     describes it.  The Python binding raises RuntimeError (the reference's per-step
 *     `except RuntimeError: continue`, Oscar/oscar/fewshot/refcoco_cpt.py:244-253, keeps working).
 *   - nothing is allocated per call: scratch is a caller-provided workspace of cpt_workspace_bytes().
 *   - all launches are asynchronous on `stream`; no entry point synchronises except cpt_check_async_error.
 */
#ifndef CPT_B200_H
#define CPT_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPT_B200_ABI_VERSION 3

typedef struct cpt_handle cpt_handle;

/* Mirrors the BertConfig fields the path reads: Oscar/oscar/modeling/modeling_bert.py:96-97,159-181. */
typedef struct {
  int32_t hidden_size, num_hidden_layers, num_attention_heads, intermediate_size;
  int32_t vocab_size, max_position_embeddings, type_vocab_size, img_feature_dim;
  int32_t use_img_layernorm, num_contrast_classes;
  float layer_norm_eps, img_layer_norm_eps;
  int32_t dtype; /* arithmetic type of the tensor-core GEMM operands: 0 = fp16 (default), 1 = bf16 */
} cpt_config;

/* One CaptionBertLayer (modeling_bert.py:129-147); fp32, row-major [out, in] like nn.Linear.weight. */
typedef struct {
  const float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b; /* attention.self.{query,key,value}            */
  const float *ao_w, *ao_b, *ao_ln_g, *ao_ln_b;   /* attention.output.{dense,LayerNorm}          */
  const float *i_w, *i_b;                         /* intermediate.dense                          */
  const float *o_w, *o_b, *o_ln_g, *o_ln_b;       /* output.{dense,LayerNorm}                    */
} cpt_layer_weights;

/* The state_dict of BertImgForPreTraining (modeling_bert.py:927-1021); head pointers may be NULL when the
 * caller only runs the encoder.  Embedding tables are used IN PLACE (caller keeps them alive until the next
 * cpt_set_weights / cpt_destroy); every other tensor is copied (and the GEMM weights converted to 16-bit). */
typedef struct {
  const float *word_emb, *pos_emb, *type_emb, *emb_ln_g, *emb_ln_b; /* bert.embeddings.*                  */
  const float *img_w, *img_b, *img_ln_g, *img_ln_b;                 /* bert.img_embedding, bert.LayerNorm */
  const float *pooler_w, *pooler_b;                                 /* bert.pooler.dense                  */
  const float *mlm_dense_w, *mlm_dense_b, *mlm_ln_g, *mlm_ln_b;     /* cls.predictions.transform.*        */
  const float *mlm_bias;      /* cls.predictions.bias; decoder.weight IS word_emb (modeling_rec.py:130-135) */
  const float *nsp_w, *nsp_b; /* cls.seq_relationship                                                     */
  const cpt_layer_weights *layers; /* host array [num_hidden_layers]                                      */
} cpt_weights;

const char *cpt_last_error(void);
int cpt_abi_version(void);

/* Model(config) + .to(device)  — BertImgModel.__init__, modeling_bert.py:153-197 */
int cpt_create(const cpt_config *cfg, int device, cpt_handle **out);
int cpt_destroy(cpt_handle *h);
/* from_pretrained / load_state_dict — Oscar/oscar/modeling/modeling_utils.py:803-851 */
int cpt_set_weights(cpt_handle *h, const cpt_weights *w, void *stream);

size_t cpt_workspace_bytes(const cpt_handle *h, int B, int T, int R);

/* BertImgModel.forward — modeling_bert.py:199-279 (2-D attention_mask, head_mask=None,
 * encoder_history_states=None: everything the CPT callers use).
 *   input_ids, token_type_ids (nullable -> 0), position_ids (nullable -> arange(T)) : int64 [B,T]
 *   attention_mask : int64 [B,T+R] (nullable -> all ones)
 *   img_feats      : fp32 [B,R,img_feature_dim] (R may be 0)
 *   seq_out        : fp32 [B,T+R,H]  ("sequence_output")
 *   pooled         : fp32 [B,H], nullable (BertPooler, only the NSP path needs it)
 *   hidden_states  : fp32 [L+1,B,T+R,H], nullable (config.output_hidden_states) */
int cpt_encoder_forward(cpt_handle *h, void *stream, const int64_t *input_ids, const int64_t *token_type_ids,
                        const int64_t *attention_mask, const int64_t *position_ids, const float *img_feats,
                        int B, int T, int R, void *workspace, size_t workspace_bytes, float *seq_out,
                        float *pooled, float *hidden_states);

/* BertLMPredictionHead at the [MASK] rows and the caller's vocabulary gather, fused:
 *   logits[b,k] = scores[b, mask_pos[b], vocab_ids[k]]
 * — modeling_rec.py:143 + Oscar/oscar/zeroshot/refcoco_cpt.py:219,234-235, fewshot/gqa_cpt.py:597-600.
 * vocab_ids NULL => all V columns (K must be vocab_size).  mask_pos int64 [B], vocab_ids int64 [K]. */
int cpt_mlm_gather_forward(cpt_handle *h, void *stream, const float *seq_out, int B, int S,
                           const int64_t *mask_pos, const int64_t *vocab_ids, int K, void *workspace,
                           size_t workspace_bytes, float *logits);

/* The reference's full output: scores[rows, V] = cls(sequence_output) over ALL rows — modeling_rec.py:143. */
int cpt_mlm_scores_forward(cpt_handle *h, void *stream, const float *seq_out, long long rows, void *workspace,
                           size_t workspace_bytes, float *scores);
size_t cpt_mlm_scores_workspace_bytes(const cpt_handle *h, long long rows);

/* NSPCPT head: cls.seq_relationship(pooled) — Oscar/oscar/modeling/modeling_vcr.py:120-121.  out fp32 [B,C]. */
int cpt_nsp_forward(cpt_handle *h, void *stream, const float *pooled, int B, float *out);

/* ---- Peer-memory exchange of the per-rank logits (replaces the two pickle all_gathers of
 * Oscar/oscar/utils/comm.py:102-142 as called from zeroshot/refcoco_cpt.py:256,262 for the part that matters: the
 * [rows, K] scores).  One process per GPU on one node.  cpt_exchange_create allocates this rank's gather buffer and
 * returns its 64-byte CUDA IPC handle; the host exchanges the handles (torch.distributed.all_gather_object) and calls
 * cpt_exchange_connect with all of them, rank-ordered.  cpt_mlm_gather_exchange is cpt_mlm_gather_forward whose decoder
 * kernel stores its logits directly into EVERY rank's buffer over NVLink and raises one flag per peer; a second small
 * kernel waits for the peers' flags and writes gathered[world * rows_per_rank, K] (rank order).  Both launches can be
 * captured in a CUDA graph (the exchange count lives on the device).  cpt_exchange_rows does the same for rows some other
 * kernel produced (e.g. NSP scores).  Every rank must make the same sequence of exchange calls.  A peer that never
 * arrives sets error flag 7 after 10 s instead of hanging the GPU. */
typedef struct cpt_exchange cpt_exchange;
#define CPT_IPC_HANDLE_BYTES 64
int cpt_exchange_create(int device, int rank, int world, int rows_per_rank, int K, cpt_exchange **out,
                        unsigned char *handle /* [CPT_IPC_HANDLE_BYTES] */);
int cpt_exchange_connect(cpt_exchange *ex, const unsigned char *handles /* [world][CPT_IPC_HANDLE_BYTES] */);
int cpt_exchange_destroy(cpt_exchange *ex);
int cpt_exchange_rows(cpt_handle *h, cpt_exchange *ex, void *stream, const float *local, int rows, float *gathered);
int cpt_mlm_gather_exchange(cpt_handle *h, cpt_exchange *ex, void *stream, const float *seq_out, int B, int S,
                            const int64_t *mask_pos, const int64_t *vocab_ids, int K, void *workspace,
                            size_t workspace_bytes, float *gathered);

/* A classification head on the pooled vector with CALLER-held weights: out[B,C] = x[B,H] W[C,H]^T + bias, all fp32
 * device pointers — the two heads of VCRQAR_NSPCPT (`cls_ans` / `cls_rat`, Oscar/oscar/modeling/modeling_vcr.py:
 * 194-252, selected per call by `head=`) without re-registering the handle's weights at every switch. */
int cpt_head_linear(cpt_handle *h, void *stream, const float *x, int B, const float *W, const float *bias, int C,
                    float *out);

/* ---- training step ------------------------------------------------------------------
 * loss = CrossEntropyLoss(ignore_index=-1)(cls(bert(...)).view(-1, V), masked_lm_labels.view(-1)) and its gradient
 * with respect to every parameter — REC_MLM_CPT.forward with masked_lm_labels, modeling_rec.py:137-150, as the
 * few-shot loops call it (Oscar/oscar/fewshot/refcoco_cpt.py:231-250, gqa_cpt.py:428-462: forward, loss.backward(),
 * optimizer.step()).  Dropout (cpt_dropout) sits where the reference has it: on the embedding outputs, the attention
 * probabilities and the two dense outputs of every layer (modeling_bert.py:57,247,266 + BertSelfOutput / BertOutput).
 * Masks are a counter-based hash of (seed, site, element index) — regenerated by the backward, never stored — so they
 * are reproducible from the seed but are NOT the masks torch's Philox stream would draw.
 *
 * cpt_train_enable(h, 1) before cpt_set_weights: later cpt_set_weights calls refresh the handle's 16-bit copies in
 * stream order instead of reallocating (call it after every optimizer step). */
int cpt_train_enable(cpt_handle *h, int on);

/* Gradient buffers, fp32, same shapes as the cpt_weights tensors of the same name.  The backward ADDS into them
 * (zero them, or keep accumulating over micro-batches as `gradient_accumulation_steps` does).  All non-NULL except
 * the img_* group when the batch has no regions. */
typedef struct {
  float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *ao_w, *ao_b, *ao_ln_g, *ao_ln_b, *i_w, *i_b, *o_w, *o_b, *o_ln_g,
      *o_ln_b;
} cpt_layer_grads;
typedef struct {
  float *word_emb, *pos_emb, *type_emb, *emb_ln_g, *emb_ln_b;
  float *img_w, *img_b, *img_ln_g, *img_ln_b;
  float *mlm_dense_w, *mlm_dense_b, *mlm_ln_g, *mlm_ln_b, *mlm_bias; /* MLM loss only (else may be NULL) */
  float *pooler_w, *pooler_b, *nsp_w, *nsp_b;                        /* NSP loss only (else may be NULL) */
  const cpt_layer_grads *layers; /* host array [num_hidden_layers] */
} cpt_grads;

/* Dropout of one training step; NULL or all-zero probabilities = off.  The backward must get the same values. */
typedef struct {
  float p_hidden; /* config.hidden_dropout_prob            */
  float p_attn;   /* config.attention_probs_dropout_prob   */
  uint64_t seed;  /* fresh per forward (the binding draws it from torch's CPU generator) */
  const uint64_t *seed_dev; /* optional DEVICE copy of the seed, read by the kernels at run time instead of `seed`:
                               lets a captured CUDA graph of the step replay with a new seed */
} cpt_dropout;

/* Which loss head a training call runs. */
enum { CPT_HEAD_MLM = 0, CPT_HEAD_NSP = 1 };

/* Bytes of the tape: activations the forward saves for the backward plus the backward's scratch.  n_rows = number
 * of labelled positions (masked_lm_labels != -1) in the batch. */
size_t cpt_train_tape_bytes(const cpt_handle *h, int B, int T, int R, int n_rows);

/* Forward with a tape.  rows int64 [n_rows]: flat indices b*(T+R)+s of the labelled positions, ascending;
 * targets int64 [n_rows]: their labels.  loss: fp32 scalar (device).  Inputs as cpt_encoder_forward. */
int cpt_train_forward_mlm(cpt_handle *h, void *stream, const int64_t *input_ids, const int64_t *token_type_ids,
                          const int64_t *attention_mask, const int64_t *position_ids, const float *img_feats, int B,
                          int T, int R, const int64_t *rows, const int64_t *targets, int n_rows,
                          const cpt_dropout *dropout, void *tape, size_t tape_bytes, float *loss);
/* Backward of the forward that filled `tape` (same inputs, same weights, same dropout).  grad_loss: fp32 scalar
 * (device), d(objective)/d(loss) — autograd's grad_output. */
int cpt_train_backward_mlm(cpt_handle *h, void *stream, const int64_t *input_ids, const int64_t *token_type_ids,
                           const int64_t *position_ids, int B, int T, int R, const int64_t *rows,
                           const int64_t *targets, int n_rows, const cpt_dropout *dropout, const float *grad_loss,
                           void *tape, size_t tape_bytes, const cpt_grads *grads);

/* Progress notifications of cpt_train_backward_*: `fn(user, stage)` is called on the host, from inside the backward
 * call, right after the launches that COMPLETE a group of gradients have been enqueued on the stream:
 *   stage 0            the loss head's tensors (cls.* / pooler / seq_relationship; not the tied word embeddings)
 *   stage 1 .. L       encoder layer L-1 .. 0  (stage s completes layer L - s)
 *   stage L + 1        everything else (embedding tables, embedding LayerNorm, region embedding)
 * A data-parallel binding launches the gradient all-reduce of that group from the callback, so that the exchange
 * overlaps the rest of the backward.  fn == NULL removes the callback. */
typedef void (*cpt_progress_fn)(void *user, int stage);
int cpt_train_set_progress_callback(cpt_handle *h, cpt_progress_fn fn, void *user);

/* The VCR few-shot loss: CrossEntropyLoss(ignore_index=-1)(cls.seq_relationship(pooled), next_sentence_label) —
 * NSPCPT.forward, Oscar/oscar/modeling/modeling_vcr.py:115-129, as vcr_nsp_cpt.py:434-473 trains it.
 * rows int64 [n_rows]: b*(T+R) of the samples whose label is not -1 (the [CLS] rows); targets their labels. */
int cpt_train_forward_nsp(cpt_handle *h, void *stream, const int64_t *input_ids, const int64_t *token_type_ids,
                          const int64_t *attention_mask, const int64_t *position_ids, const float *img_feats, int B,
                          int T, int R, const int64_t *rows, const int64_t *targets, int n_rows,
                          const cpt_dropout *dropout, void *tape, size_t tape_bytes, float *loss);
int cpt_train_backward_nsp(cpt_handle *h, void *stream, const int64_t *input_ids, const int64_t *token_type_ids,
                           const int64_t *position_ids, int B, int T, int R, const int64_t *rows,
                           const int64_t *targets, int n_rows, const cpt_dropout *dropout, const float *grad_loss,
                           void *tape, size_t tape_bytes, const cpt_grads *grads);

/* ---- optimizer step (SURVEY.md 8f "optimizer") -----------------------------------------------------------------
 * One launch updates every parameter tensor.  mode 0 = torch.optim.AdamW (fewshot/refcoco_cpt.py:342); mode 1 =
 * pytorch-transformers 1.x AdamW (fewshot/gqa_cpt.py:342, vcr_nsp_cpt.py; weight decay applied after the Adam update,
 * eps added to sqrt(v) before the bias correction).  `tensors` and `chunks` are DEVICE arrays the binding builds:
 * chunk i covers elements [offset, offset+count) of tensor `tensor`, count <= 16384.  grad_scale: optional device
 * scalar every gradient is multiplied by (loss-scale / clipping coefficient), NULL = 1. */
typedef struct {
  float *p;       /* parameter (fp32 master weight), updated in place */
  const float *g; /* gradient                                          */
  float *m, *v;   /* exp_avg, exp_avg_sq                               */
  int64_t n;
  float lr, wd;   /* the group's learning rate and weight decay        */
  float bc1, bc2; /* 1 - beta1^t, 1 - beta2^t of this tensor (1, 1 when bias correction is off) */
} cpt_adam_tensor;
typedef struct {
  int32_t tensor, count;
  int64_t offset;
} cpt_adam_chunk;
int cpt_adamw_step(int device, void *stream, const cpt_adam_tensor *tensors, const cpt_adam_chunk *chunks,
                   int n_chunks, float beta1, float beta2, float eps, int mode, const float *grad_scale);

/* torch.nn.utils.clip_grad_norm_(parameters, max_norm) (Oscar/oscar/fewshot/gqa_cpt.py:454, vcr_nsp_cpt.py) fused with
 * the update: one launch over the same tables computes norm_out = ||grad_scale_in * g||_2 over ALL tensors and
 * scale_out = grad_scale_in * min(1, max_norm / (norm + 1e-6)); pass scale_out as cpt_adamw_step's grad_scale.  The
 * gradients themselves are not modified.  scratch: 16 bytes of device memory, zeroed once by the caller (the kernel
 * leaves it zeroed).  grad_scale_in may be NULL (= 1). */
int cpt_grad_clip_scale(int device, void *stream, const cpt_adam_tensor *tensors, const cpt_adam_chunk *chunks,
                        int n_chunks, float max_norm, const float *grad_scale_in, void *scratch, float *norm_out,
                        float *scale_out);

/* ---- host input assembly on the device (SURVEY.md 8f "input assembly") ---------------------------------------------
 * One padded CPT batch from token-id lists and a packed region-feature store: the reference's per-sample tokenize()
 * ([CLS] a [SEP] b [SEP], pair truncation to T - 3, segment ids, zero padding, mask over tokens and boxes —
 * Oscar/oscar/datasets/refcoco_zsl_cpt_dataset.py:191-302), feature padding to R rows (:119-120), [MASK] position
 * (:118) and the collate's stacking (Oscar/oscar/zeroshot/refcoco_cpt.py:159-172).  store: fp32 [rows, img_feature_dim]
 * on the device; row b takes n_boxes[b] <= R rows from feat_row0[b].  tok_a / tok_b: flat int32 token ids with CSR
 * offsets [B + 1]; has_b[b]: text_b was non-empty.  Outputs int64 [B,T], [B,T], [B,T+R], [B] and fp32 [B,R,F]. */
int cpt_assemble_inputs(cpt_handle *h, void *stream, int B, int T, int R, const float *store, const int64_t *feat_row0,
                        const int32_t *n_boxes, const int32_t *tok_a, const int32_t *a_off, const int32_t *tok_b,
                        const int32_t *b_off, const int32_t *has_b, int cls_id, int sep_id, int pad_id, int mask_id,
                        int64_t *input_ids, int64_t *segment_ids, int64_t *input_mask, int64_t *mask_pos,
                        float *img_feats);

/* ---- CPT decision per query on the device (SURVEY.md 8f "scoring") ----------------------------------------------
 * Replaces the per-image Python loops of Oscar/oscar/zeroshot/refcoco_cpt.py:222-254, fewshot/refcoco_cpt.py:273-297
 * and fewshot/vcr_nsp_cpt.py:600-604, and the IoU > 0.5 hit test of zeroshot/refcoco_cpt.py:268-276 +
 * Oscar/oscar/utils/iou.py:1-12.  logits fp32 [rows, ld]: for modes 0 (zero-shot: score = colour logit) and 1
 * (few-shot: colour / none) the K used columns are K - 1 palette colours followed by "none"; row r only counts its
 * first col_start[r+1] - col_start[r] colour columns (its own colour set; col_start NULL = all K - 1).  mode 2 (VCR):
 * logits are the NSP scores [rows, K], score = 1 - softmax[:, 1].  Query q owns rows [row_start[q], row_start[q+1]).
 * pick int32 [Q]: index into the query's concatenated scores (the reference's max_idx, torch.argmax tie rules).
 * rects double [total valid columns, 4] (x1 y1 x2 y2, collected order), gt double [Q,4] (x y w h): when given,
 * pick_rect [Q,4], iou [Q] and correct int32 [Q] (iou > 0.5) are written too (any of them may be NULL). */
int cpt_score_queries(cpt_handle *h, void *stream, const float *logits, long long ld, int K, int Q,
                      const int32_t *row_start, const int32_t *col_start, const double *rects, const double *gt,
                      int mode, int32_t *pick, double *pick_rect, double *iou, int32_t *correct);

/* Blocks until `stream` drains; reports device-side input errors (token id / position out of range, the
 * IndexError the reference's nn.Embedding would raise) and launch failures. */
int cpt_check_async_error(cpt_handle *h, void *stream);

/* ---- launch accounting and per-kernel-class timing (bench.py roofline leg) ------------------------------- */
enum {
  CPT_K_EXTMASK = 0, CPT_K_EMBED, CPT_K_CAST, CPT_K_GEMM_IMG, CPT_K_LN, CPT_K_GEMM_QKV, CPT_K_ATTN,
  CPT_K_GEMM_AO, CPT_K_GEMM_UP, CPT_K_GEMM_DOWN, CPT_K_HEAD, CPT_K_GEMM_HEAD, CPT_K_GEMM_OTHER,
  CPT_K_GEMM_DGRAD, CPT_K_GEMM_WGRAD, CPT_K_ATTN_BWD, CPT_K_TRAIN_ROWWISE, CPT_K_TRANSPOSE, CPT_K_COLSUM, CPT_K_LN_BWD,
  CPT_K_EMBED_BWD, CPT_K_CHAIN, CPT_K_COUNT
};
const char *cpt_kernel_name(int tag);
/* kernels launched by this handle since cpt_create */
long long cpt_launch_count(const cpt_handle *h);
/* on != 0: bracket every subsequent launch with CUDA events on its stream (a few us of host time per launch;
 * leave off in timed runs).  Resets the accumulators.  Synchronises the device. */
int cpt_profile_enable(cpt_handle *h, int on);
/* Synchronises, then returns per-class device milliseconds and launch counts accumulated since the last
 * enable/read (arrays of CPT_K_COUNT), and resets them. */
int cpt_profile_read(cpt_handle *h, double *ms, long long *launches);

/* ---- kernel-level entry points (unit tests, bench roofline leg) ----------------------------------------- */
/* out[M,N] = epi(A[M,K] . W[N,K]^T): A, W 16-bit (dtype as in cfg) with leading dims lda/ldw (elements, multiples
 * of 8); epi: 0 bias, 1 bias+erf-GELU, 2 bias+fp32 residual; out_fp32: 0 -> 16-bit out, 1 -> fp32 out.
 * epi | 0x100: operands are given transposed, A as [K,M] and W as [K,N] (out = A^T . W, the weight-gradient form;
 * lda/ldw are then the pitches of those); epi | 0x400: only W is given transposed, as [K,N] (out = A . W, the
 * data-gradient form on an nn.Linear weight); epi | 0x200: out += result (fp32 out only); epi | (s << 12): split K into
 * s pieces accumulated at the destination (with 0x200 and bias == NULL only; otherwise ignored).
 * tile_cfg: 0 = library default, else block_n (64/128/192/256) + 1000 * (CTAs per MMA: 1 or 2), e.g. 2256 =
 * 256-wide tiles computed by CTA pairs (tcgen05 cta_group::2). */
int cpt_gemm(cpt_handle *h, void *stream, const void *A, long long lda, const void *W, long long ldw, int M, int N,
             int K, const float *bias, const float *resid, long long ldr, int epi, int out_fp32, void *out,
             long long ldo, int tile_cfg);
/* Debug: per-CTA cycle counters of the LAST GEMM launch (needs CPT_B200_TRACE=1 in the environment at cpt_create):
 * out[cta][16] = {producer total, producer waiting for a free stage, MMA total, MMA waiting for operands,
 * MMA waiting for a drained accumulator, epilogue total, epilogue waiting for MMA, tiles, prologue cycles, prologue
 * + PDL wait, CTA lifetime cycles, globaltimer ns at entry, at exit, 0, 0, 0}. */
int cpt_gemm_trace(cpt_handle *h, long long *out, int max_ctas);
/* ctx[B*S,H] = softmax(QK^T/sqrt(dH) + (1-mask)*-1e4) V from packed qkv[B*S,3H] (16-bit); ext_mask fp32 [B,S].
 * impl: 0 = production (persistent ping-pong tcgen05 kernel, one thread per query row), 1 = CUDA-core cross-check
 * kernel, 2 = single-tile tcgen05 kernel (one CTA per (head, query tile, sample)), 3 = two-threads-per-row pipelined
 * tcgen05 kernel (earlier design, kept as a cross-check). */
int cpt_attention(cpt_handle *h, void *stream, const void *qkv, const float *ext_mask, int B, int S, void *ctx,
                  int impl);
/* d(qkv)[B*S,3H] from d(ctx)[B*S,H] (16-bit), probabilities recomputed from qkv.  impl: -1 = library choice, 0 =
 * tcgen05 kernels (one 128 x 128 block for S <= 128, 2 x 2 blocks with a statistics pre-pass above), 1 = CUDA-core
 * kernel. */
int cpt_attention_backward(cpt_handle *h, void *stream, const void *qkv, const void *dctx, const float *ext_mask,
                           int B, int S, void *dqkv, int impl);
/* Dataflow chain (cpt_b200/csrc/chain_sm100.cuh): up to 8 dependent stages over the same M rows as ONE persistent
 * launch — the encoder forward runs attention.output.dense + LayerNorm + intermediate.dense + output.dense + LayerNorm +
 * the next layer's query/key/value projection (modeling_bert.py:85,144-145, then :38-40) this way.  A stage is
 *   kind 0: out[M,N] = A[M,K] . W[N,K]^T + bias (16-bit operands; gelu: erf-GELU; out_fp32 = 0: 16-bit out,
 *           out_fp32 = 1: the fp32 tile is ADDED into `out` (which already holds the residual); ksplit: K pieces), or
 *   kind 0 with ln = 1: out32 / out16 [M,N] = LayerNorm(A . W^T + bias + resid[M,ldr]) * gamma + beta — the LayerNorm
 *           runs in the tile epilogues, row statistics are exchanged between the N tiles of a row through L2, or
 *   kind 1: out32 / out16 [M,N] = LayerNorm(ln_in[M,N]) * gamma + beta.
 * dep_stage: the earlier stage whose output this one reads (-1: data from an earlier launch); rows are handed from
 * stage to stage through readiness counters per 128-row tile, there is no grid-wide barrier. */
typedef struct {
  int32_t kind, M, N, K, gelu, out_fp32, ksplit, dep_stage;
  const void *A;
  int64_t lda;
  const void *W;
  int64_t ldw;
  const float *bias;
  void *out;
  int64_t ldo;
  const float *ln_in, *gamma, *beta;
  float eps;
  float *out32;
  void *out16;
  int32_t ln; /* 1: LayerNorm finished in the epilogue; 2: deferred — out32 / out16 are PRE-LayerNorm, `part` gets the
                 (mean, M2) partials of the rows ([2 * ceil(N / 256)][rows padded to 256] float2), the residual is
                 normalised on the fly from `rpart` (its own partials; NULL = used as it is) with gamma / beta / eps */
  const float *resid;
  int64_t ldr;
  void *part;
  const void *rpart;
  const void *apart; /* kind 0 consumer of a deferred LayerNorm: A holds raw rows, W carries gamma, bias = c vector,
                        gvec = g vector (cpt_b200/csrc/rowwise.cuh fold_weight_kernel); statistics from `apart` */
  const float *gvec;
} cpt_chain_stage;
int cpt_chain_run(cpt_handle *h, void *stream, const cpt_chain_stage *stages, int n_stages);
/* Debug: event log of the LAST chain launch (needs CPT_B200_CHAIN_TRACE=1 in the environment at cpt_create).
 * out = [pairs][2] {globaltimer ns, clock64 at CTA entry} followed by [pairs][pitch][16] SM-clock stamps per task of the
 * leader CTA's list (chain_sm100.cuh: ChainParams::trace). */
int cpt_chain_trace(cpt_handle *h, long long *out, long long max_words, int *pairs, int *pitch);
/* y = LayerNorm(x) rows: fp32 in, fp32 and/or 16-bit out (either may be NULL). */
int cpt_layernorm(cpt_handle *h, void *stream, const float *x, int M, const float *gamma, const float *beta,
                  float eps, float *out32, void *out16);
int cpt_cast16(cpt_handle *h, void *stream, const float *x, long long rows, int cols, int ld_out, void *out16);

#ifdef __cplusplus
}
#endif
#endif /* CPT_B200_H */
