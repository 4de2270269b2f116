/* modcr_b200 -- C ABI of the B200-native (sm_100a) ModCR training / scoring hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  The reference has no FFI
 * layer (it is pure PyTorch); each entry point therefore names the reference *operator site* it
 * replaces, and INTEGRATION.md shows the ctypes binding that the reference-side Python modules
 * (modeling/modeling_bert.py, modeling/modeling_vcr_chunkalign_v10.py, modeling/modeling_utils.py)
 * use to reach it.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (including workspaces); nothing is
 *     allocated or freed by the library; all work is enqueued on `stream` (a cudaStream_t);
 *   - activations are bf16, row-major, 16-byte aligned; statistics, losses, masks and parameter
 *     gradients are fp32; index tensors are int64 exactly as the reference batch contract
 *     (Data/VCRChunkAlign.py:690-741) delivers them;
 *   - return value 0 = success, negative = error (enum below); the text of the last error of the
 *     calling thread is returned by modcr_last_error_string().  Shapes are validated BEFORE launch.
 *     There is no CPU fallback: on a machine without an sm_100 device every compute entry fails.
 */
#ifndef MODCR_B200_H_
#define MODCR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MODCR_VERSION 100 /* 0.1.0 */

enum {
  MODCR_OK = 0,
  MODCR_ERR_SHAPE = -1,
  MODCR_ERR_ALIGN = -2,
  MODCR_ERR_ARCH = -3,
  MODCR_ERR_CUDA = -4,
  MODCR_ERR_DRIVER = -5,
  MODCR_ERR_WORKSPACE = -6,
  MODCR_ERR_UNSUPPORTED = -7
};

int modcr_version(void);
const char* modcr_last_error_string(void);
/* number of CUDA kernels this library has launched in the calling process (monotonic) */
uint64_t modcr_launch_count(void);

/* SM budget of the persistent kernels (GEMM, persistent attention forward): they launch one CTA (pair) per SM with a static
 * tile partition, so CTAs of a concurrent NCCL collective must find free SMs instead of queueing one of these kernels
 * behind them.  sms <= 0 restores "all SMs"; returns the previous value.  Takes effect for the launches enqueued after
 * the call (including launches captured into a CUDA graph). */
int modcr_set_sm_limit(int32_t sms);

/* Workspace query.  op: 1 = prefix_attn_bwd {N,S,heads}; 2 = img_embed_ln {rows, feat_dim, H} */
size_t modcr_workspace_bytes(int op, const int64_t* dims, int ndims);

/* ---------------------------------------------------------------------------------------------
 * tcgen05 GEMM with fused epilogues:  C[M,N] = epi(alpha * sum_k A[m,k] B[n,k])
 * Replaces every nn.Linear on the path:
 *   qkv_gemm        query/key/value        modeling/modeling_bert.py:38-44, v10:57-65
 *   gemm_bias_res   BertSelfOutput.dense / BertOutput.dense + residual   a_transformers/modeling_bert.py:369-373, 447-451
 *   gemm_bias_gelu  BertIntermediate       a_transformers/modeling_bert.py:434-437
 *   pooler          BertPooler (tanh)      a_transformers/modeling_bert.py:640-646
 *   img_embedding   modeling/modeling_transfomres.py:676-681
 * and their dgrad / wgrad in the backward pass (autograd of the reference).
 *   a_mn = 0: A stored [M,K];  a_mn = 1: A stored [K,M]      (same for B with N)
 *   act: 0 none, 1 GELU(erf), 2 tanh, 3 ReLU;  in_op: 0 none, 1 C = f(acc) + in, 2 C = f(acc) * GELU'(in),
 *   3 C = f(acc) + LayerNorm(in) (see in_mean .. in_beta below)
 *   aux: optional bf16 copy of the pre-activation;  c_fp32 + accumulate / split_k: wgrad
 *   a_colsum (a_mn = 1 only): a_colsum[m] += sum_k A[k,m], fp32 -- the bias gradient dY^T 1 of the Linear whose weight
 *   gradient dW = dY^T X this call computes, summed from the dY tiles the kernel stages anyway (replaces the reference's
 *   autograd bias reduction, a separate pass over dY)
 */
typedef struct modcr_gemm_desc {
  const void* A;
  const void* B;
  void* C;
  int64_t lda, ldb, ldc;
  int32_t a_mn, b_mn, c_fp32, accumulate, split_k;
  int32_t M, N, K;
  float alpha;
  const float* bias;
  int32_t act;
  void* aux;
  int64_t ldaux;
  const void* epi_in;
  int64_t ld_in;
  int32_t in_op;
  float* a_colsum;
  /* training-mode dropout of the dense output inside the epilogue: C = dropout(f(acc + bias)) (+ in), i.e.
   * LayerNorm(dropout(dense(x)) + residual) of BertSelfOutput / BertOutput without an extra pass; mask index m * N + n,
   * drop_p = 0: off */
  const uint32_t* drop_seed;
  uint32_t drop_site;
  float drop_p;
  /* in_op = 3 (fp32 C and fp32 epi_in only): the residual is LayerNorm(epi_in), rebuilt in the epilogue from the
   * PRE-LayerNorm tensor the previous block keeps for its backward pass, its row statistics and the LayerNorm's affine
   * parameters:  C = f(acc + bias) + (in[m,n] - in_mean[m]) * in_rstd[m] * in_gamma[n] + in_beta[n].
   * BertSelfOutput / BertOutput (a_transformers/modeling_bert.py:369-373, 447-451) add `input_tensor`, which IS the previous
   * LayerNorm's output: this keeps the residual stream in fp32 without ever writing an fp32 copy of a LayerNorm output. */
  const float* in_mean;
  const float* in_rstd;
  const float* in_gamma;
  const float* in_beta;
} modcr_gemm_desc;
int modcr_gemm(const modcr_gemm_desc* d, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Row maps: logical row m -> physical row (m / rows_per_seq) * seq_stride + row0 + m % rows_per_seq
 * (rows_per_seq <= 0: identity).  Lets the embedding kernels write text rows [0,T) and region rows
 * [T,S) of the concatenated [N,S,H] buffer directly (no torch.cat, modeling_transfomres.py:684). */
typedef struct modcr_rowmap {
  int32_t rows_per_seq, seq_stride, row0;
} modcr_rowmap;

/* LayerNorm over the last dim (H % 256 == 0, H <= 1024); BertSelfOutput/BertOutput/img LayerNorm.
 * x is bf16 (x_f32 = 0) or fp32 (x_f32 = 1: the residual stream is kept in fp32 -- pre-LayerNorm sums
 * and the residual operand never round to bf16; measured to cut the logit error 3x).  y is the bf16
 * tensor-core operand; y32 (optional) the fp32 copy consumed by the next residual add. */
int modcr_layernorm_fwd(const void* x, int32_t x_f32, modcr_rowmap xmap, const float* gamma, const float* beta,
                        void* y, float* y32, modcr_rowmap ymap, float* mean, float* rstd, int32_t M, int32_t H,
                        float eps, void* stream);
/* dbias (optional) receives sum_rows dx = bias gradient of the Linear that produced x. fp32 outputs accumulate.
 * dx_masked (optional, training-mode dropout of that Linear's output: LayerNorm(dropout(dense(.)) + residual)): second
 * bf16 [M,H] output dx * mask / (1 - p), the gradient w.r.t. the dense output, with the mask of the forward
 * modcr_dropout call (same seed / site, element index m * H + column); dbias then sums the masked gradient. */
int modcr_layernorm_bwd(const void* dy, modcr_rowmap dymap, const void* x, int32_t x_f32, modcr_rowmap xmap,
                        const float* mean, const float* rstd, const float* gamma, void* dx, modcr_rowmap dxmap,
                        float* dgamma, float* dbeta, float* dbias, int32_t M, int32_t H, void* dx_masked,
                        const uint32_t* drop_seed, uint32_t drop_site, float drop_p, void* stream);

/* BertEmbeddings.forward (a_transformers/modeling_bert.py:184-211): gather + sum + LayerNorm. */
int modcr_txt_embed_ln_fwd(const int64_t* input_ids, const int64_t* token_type_ids, const int64_t* position_ids,
                           const float* word, const float* type, const float* pos, const float* gamma,
                           const float* beta, void* y, float* y32, modcr_rowmap ymap, float* mean, float* rstd,
                           int32_t N, int32_t T, int32_t H, float eps, void* stream);
int modcr_txt_embed_ln_bwd(const void* dy, modcr_rowmap dymap, const int64_t* input_ids,
                           const int64_t* token_type_ids, const int64_t* position_ids, const float* word,
                           const float* type, const float* pos, const float* gamma, const float* mean,
                           const float* rstd, float* dword, float* dtype, float* dpos, float* dgamma, float* dbeta,
                           int32_t N, int32_t T, int32_t H, int64_t padding_idx, void* stream);

/* img_embedding + LayerNorm (modeling_transfomres.py:676-681, v10:339-343):
 *   packed = bf16(feats) zero-padded to Dpad (TMA needs 16-byte row strides; 2054 -> 2112)
 *   pre    = packed W^T + bias            (tcgen05 GEMM; W is bf16 [H, Dpad]; fp32 output)
 *   y, y32 = LN(pre)                      written through ymap (bf16 operand + fp32 residual copy)
 * Workspaces `packed` bf16 [rows, Dpad] and `pre` fp32 [rows, H] are kept for the backward pass. */
int modcr_img_embed_ln_fwd(const float* feats, int64_t rows, int32_t D, int32_t Dpad, const void* W,
                           const float* bias, const float* gamma, const float* beta, int32_t use_ln, float eps,
                           void* packed, float* pre, void* y, float* y32, modcr_rowmap ymap, float* mean,
                           float* rstd, int32_t H, const float* extra32, void* stream);
/* extra32 (fp32 [rows,H] or NULL): added to the projection inside the GEMM epilogue before the LayerNorm -- the
 * type + box-position term of BertImgModel_gtToken (modcr_pos_type_embed_fwd; modeling/modeling_bert.py:393). */
/* dW_pad: fp32 [H, Dpad] (overwritten); dbias/dgamma/dbeta accumulate; dpre: bf16 [rows, H] workspace. */
int modcr_img_embed_ln_bwd(const void* dy, modcr_rowmap dymap, const void* packed, const float* pre,
                           const float* mean, const float* rstd, const float* gamma, int32_t use_ln, void* dpre,
                           float* dW_pad, float* dbias, float* dgamma, float* dbeta, int64_t rows, int32_t Dpad,
                           int32_t H, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused prefix attention (CaptionBertSelfAttention.forward after the projections:
 * modeling/modeling_bert.py:46-72, v10:80-104).  K/V = [prefix rows ; sequence rows], never
 * concatenated in HBM.  mode: 0 key-bias only; 1 seq_enc layers 0-2; 2 seq_enc layers 9-11
 * (v10:171-232).  head_dim = 64. */
typedef struct modcr_attn_desc {
  int32_t N, S, P, heads, T, mode;
  float scale;
  const void* qkv;        /* bf16 [N,S,3H] */
  const void* pkv;        /* bf16 [N,P,2H] or NULL */
  const float* kbias;     /* fp32 [N,P+S] additive key bias = (1-mask)*-10000 */
  const uint32_t* cbits;  /* [N,T,cwords] chunk bit mask (modes 1/2) */
  int32_t cwords;
  void* ctx;              /* bf16 [N,S,H] */
  float* lse;             /* fp32 [N,heads,S] */
  float* probs_img;       /* fp32 [N,heads,T,S-T] or NULL */
  const void* dctx;       /* backward: bf16 [N,S,H] */
  const float* dalign;    /* backward: fp32 [N,T,S-T] or NULL */
  void* dqkv;             /* backward: bf16 [N,S,3H] */
  void* dpkv;             /* backward: bf16 [N,P,2H] */
  float* dq_ws;           /* backward workspace, modcr_workspace_bytes(1, {N,S,heads}) */
  /* training-mode dropout of the attention probabilities (modeling_bert.py:62, v10:94): softmax -> dropout -> P V, the
   * returned text->region probabilities are post-dropout.  mask = hash(*drop_seed, drop_site, (n, head, q, key));
   * drop_p = 0 disables it.  The backward pass must be given the same (seed, site, p). */
  const uint32_t* drop_seed;
  uint32_t drop_site;
  float drop_p;
  /* int64 [N,T] or NULL (align_pos of the batch contract, Data/VCRChunkAlign.py:733): when given, only text rows with a
   * non-zero entry are written to probs_img (forward) / read from it and given dalign (backward) -- the alignment loss
   * (v10:1071-1077) selects exactly those rows, so the other 97 % of the [N,heads,T,R] block never travel.  NULL: all rows. */
  const int64_t* row_sel;
} modcr_attn_desc;
int modcr_prefix_attn_fwd(const modcr_attn_desc* d, void* stream);
int modcr_prefix_attn_bwd(const modcr_attn_desc* d, void* stream);

/* Chunk-mean queries (v10:66-78), in place on columns [0,H) of the QKV buffer (row stride ld).
 * chunk_id: int32 [N,T], -1 for tokens outside every chunk.  Self-adjoint: backward = same call on dQ. */
int modcr_chunk_mean_q(void* q, int64_t ld, int32_t N, int32_t S, int32_t H, const int32_t* chunk_id, int32_t T,
                       int32_t max_chunks, void* stream);

/* Alignment loss (v10:982-987, 1071-1077).  probs: fp32 [nl,N,heads,T,R]; out[4] = {loss, M, correct, sum}.
 * block [N,T,R] = sum over the nl * heads slices (the reference's stack(...).sum(1).sum(1)); with align_pos / total_label /
 * row_loss / out given: masked_fill(block == 0, -1e5) -> softmax over R -> CrossEntropy(probabilities, total_label) over
 * the rows with align_pos == 1 (double softmax as written, v10:1073-1077).  row_sel (optional, = align_pos): rows without
 * a non-zero entry were never written by modcr_prefix_attn_fwd(row_sel) and get a zero block row. */
int modcr_align_softmax_ce_fwd(const float* probs, int32_t nl, int32_t heads, const int64_t* align_pos,
                               const int64_t* total_label, const int64_t* row_sel, float* block, float* row_loss,
                               float* out, int32_t N, int32_t T, int32_t R, void* stream);
int modcr_align_softmax_ce_bwd(const float* block, const int64_t* align_pos, const int64_t* total_label,
                               const float* out, float gout, float* dblock, int32_t N, int32_t T, int32_t R,
                               void* stream);

/* Training driver (run_PMR_ModCR.py:127-145, 211-227): torch.nn.utils.clip_grad_norm_ over all parameters followed by
 * transformers.AdamW.step() with two learning-rate groups (seq_enc x 0.1), eps 1e-5, bias correction, fused into two
 * launches over a block table (one block per 8192-element chunk of one tensor: blk_tensor / blk_chunk, int32 [nblocks]).
 *   meta  : device array, one entry per parameter tensor (fp32 master weight, AdamW moments, length, LR multiplier, wd)
 *   grads : device array of fp32 gradient pointers in the same order; NULL = no gradient this step (tensor skipped)
 *   hyper : device fp32[4] = {lr, 1 - beta1^t, 1 - beta2^t, max_grad_norm (<= 0: no clipping)}  (device memory, so the
 *           step can be replayed inside a CUDA graph while the schedule advances)
 * modcr_grad_sqnorm adds the sum of squares of every gradient to *out_sq (zeroed by the caller); modcr_adamw_step
 * scales the gradients by min(1, max_norm / (sqrt(*grad_sq) + 1e-6)) on the fly (grad_sq NULL: no clipping) and applies
 *   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr sqrt(1-b2^t)/(1-b1^t) m / (sqrt(v) + eps) ; p -= lr wd p
 * and, when `shadow` is set, rewrites the kernel-operand copy of the parameter in the same pass (shadow_kind 1: bf16,
 * 2: fp32; same element order as p) so the next forward does not re-cast the weights.
 * modcr_grad_clip is the in-place clip (g *= min(1, max_norm / (sqrt(*grad_sq) + 1e-6))) of the accumulation micro-steps
 * that do not end in optimizer.step() (run_PMR_ModCR.py:215 clips after every backward). */
typedef struct modcr_optim_tensor {
  float* p;
  float* m;
  float* v;
  void* shadow;
  int64_t n;
  float lr_mul;
  float wd;
  int32_t shadow_kind;
  int32_t reserved;
} modcr_optim_tensor;
int modcr_grad_clip(const modcr_optim_tensor* meta, float* const* grads, const int32_t* blk_tensor,
                    const int32_t* blk_chunk, int32_t nblocks, const float* grad_sq, float max_norm, void* stream);
int modcr_grad_sqnorm(const modcr_optim_tensor* meta, const float* const* grads, const int32_t* blk_tensor,
                       const int32_t* blk_chunk, int32_t nblocks, float* out_sq, void* stream);
int modcr_adamw_step(const modcr_optim_tensor* meta, const float* const* grads, const int32_t* blk_tensor,
                     const int32_t* blk_chunk, int32_t nblocks, const float* hyper, const float* grad_sq, float beta1,
                     float beta2, float eps, void* stream);

/* Gradient buckets of the data-parallel all-reduce (the reference's bucketing helper is
 * utils/distributed.py:46-97, flat 10 MiB buffers filled by torch copies): ONE launch copies up to MODCR_MULTI_MAX
 * tensors, dst[i][0..n[i]) = scale * src[i][0..n[i]), casting fp32 <-> bf16 on the way (src_bf16 / dst_bf16 give the two
 * element types).  src / dst / n are HOST arrays of `count` entries (the device pointers they hold travel by value in
 * the kernel parameters, so the call can be captured into a CUDA graph).  Used to gather a layer's fp32 gradients into
 * the bf16 bucket as soon as its wgrads have landed and to scatter the reduced bucket back. */
#define MODCR_MULTI_MAX 64
int modcr_multi_copy(const void* const* src, void* const* dst, const int64_t* n, int32_t count, int32_t src_bf16,
                     int32_t dst_bf16, float scale, void* stream);

/* Training-mode dropout of the hidden-state sites (a_transformers/modeling_bert.py:209, 371, 449;
 * modeling_transfomres.py:681; v10:343, 832, 866): out16 = in16 * m (bf16, may alias), io32 = io32 * m (+ res32),
 * m = keep / (1 - p) with keep(seed, site, element index) a counter-based hash -- the backward pass regenerates the mask
 * from the same (seed, site).  `seed` points to one uint32 in device memory (so a replayed CUDA graph can advance it). */
int modcr_dropout(const void* in16, void* out16, float* io32, const float* res32, int64_t n, const uint32_t* seed,
                  uint32_t site, float p, void* stream);

/* cross_attention_lyx with a single query per sequence (ClsLayer_lyx, modeling/modeling_vcr_chunkalign_v10.py:692-797,
 * 856-857): out[n] = concat_h softmax_l(q_h[n] . k_h[n,l] / sqrt(d_h)) v_h[n,l].  q, out [N,H]; k, v [N,L,H]; all bf16.
 * No mask (the reference never forwards word_mask), tau = 1, dropout p = 0.  lse (optional, fp32 [N,heads]) is the
 * per-head log-sum-exp the backward pass needs; the backward (autograd of v10:710-797) returns dq [N,H], dk, dv
 * [N,L,H] in bf16. */
int modcr_cross_attn_q1_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int32_t N, int32_t L,
                            int32_t H, int32_t heads, const uint32_t* drop_seed, uint32_t drop_site, float drop_p,
                            void* stream);
int modcr_cross_attn_q1_bwd(const void* q, const void* k, const void* v, const void* out, const float* lse,
                            const void* dout, void* dq, void* dk, void* dv, int32_t N, int32_t L, int32_t H,
                            int32_t heads, const uint32_t* drop_seed, uint32_t drop_site, float drop_p, void* stream);

/* ClsLayer2 attention pooling (v10:816-829). q [N,H], k [N,L,H] bf16 (qk_f32=0) or fp32 (qk_f32=1: the unscaled
 * 768-dim scores are the most rounding-sensitive spot of the path), mask fp32 [N,L] additive.  drop_*: training-mode
 * dropout of the attention weights (v10:828; drop_p = 0 disables it; w is returned undropped, the pooled vector uses
 * the dropped weights). */
int modcr_cls_pool_attn_fwd(const void* q, const void* k, int32_t qk_f32, const float* mask, float* w, void* pooled,
                            int32_t N, int32_t L, int32_t H, const uint32_t* drop_seed, uint32_t drop_site, float drop_p,
                            void* stream);
int modcr_cls_pool_attn_bwd(const void* dpooled, const void* q, const void* k, int32_t qk_f32, const float* w, void* dq,
                            void* dk, int32_t N, int32_t L, int32_t H, const uint32_t* drop_seed, uint32_t drop_site,
                            float drop_p, void* stream);

/* Fused multiple-choice classifier + cross-entropy (v10:1063-1069, 363-373;
 * modeling/modeling_ensemble.py:511-536; modeling/modeling_bert.py:553-567, 690-710). */
typedef struct modcr_scorer_desc {
  const void* x;          /* bf16 [rows,K] */
  const float* W;         /* fp32 [C,K] */
  const float* b;         /* fp32 [C] or NULL */
  int32_t rows, K, C, G;  /* CE over the view [rows*C/G, G] */
  const int64_t* hard;    /* int64 [rows*C/G] or NULL */
  const float* soft;      /* fp32 [rows*C/G, G] or NULL */
  float* logits;          /* fp32 [rows,C] */
  float* loss;            /* fp32 [1] */
  float* mp;              /* optional fp32 [rows] = softmax(logits)[:,1]  (binary_to_mp) */
  int64_t* pred;          /* optional int64 [rows/num_choice] */
  uint8_t* matched;       /* optional uint8 [rows/num_choice] */
  int32_t num_choice;
  float gout;
  void* dx;               /* backward: bf16 [rows,K] */
  float* dW;              /* backward: fp32 [C,K], accumulates */
  float* db;              /* backward: fp32 [C], accumulates */
  int32_t loss_kind;      /* 0: cross-entropy over [rows*C/G, G]; 1: instance_bce_with_logits (modeling/modeling_bert.py:416-421):
                             mean over all elements of BCE-with-logits against `soft`, times G (= labels.size(1)) */
} modcr_scorer_desc;
int modcr_mc_scorer_ce_fwd(const modcr_scorer_desc* d, void* stream);
int modcr_mc_scorer_ce_bwd(const modcr_scorer_desc* d, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Device-side batch packer.  Replaces the collate function SNLIGPT_gen_collate
 * (Data/VCRChunkAlign.py:690-741) and the per-choice tensor building of the dataset's __getitem__ (:655-686): the host
 * fills ONE pinned buffer, copies it once, and this call expands it into the padded batch of the reference's batch
 * contract plus the packed forms the kernels read (chunk-bit mask, chunk ids).
 *   header int32 [N,4]   : text length L, number of boxes, label (0/1), index of the question's feature block
 *   tokens int32 [N,T,2] : word 0 = input id; word 1 = (chunk id + 1) | segment << 10 | align_pos << 11 | total_label << 12
 *                          (chunk id of tokens 1..L-2 = position of the token's chunk in `offsets`, :667-670; 0 = none)
 *   feats  fp32 [U,R,D]  : one block per QUESTION; the four choices share it (4x fewer bytes over PCIe)
 * Outputs (caller-allocated; optional ones may be NULL): input_ids / token_type_ids / total_label / align_pos int64 [N,T]
 * (pad value 0, :698-705), label int64 [N] and label_f fp32 [N] (:706-708), input_mask fp32 [N,T+R] = text mask || box
 * mask (:709-714), chunk_id int32 [N,T] (-1 outside chunks), chunk_bits uint32 [N,T,cwords] with cwords = ceil(T/32),
 * bit k of row q = chunk mask [q,k] of utils/GetChunk_v4_vcr.py:104-140 zero-padded to [T,T] (:716-723), chunk_mask the
 * same as dense fp32 [N,T,T] (only for callers that want the reference tensor), img_feat fp32 [N,R,D] with rows beyond
 * the question's box count zero.  T <= 512, D even; tokens / feats / img_feat 8-byte aligned. */
typedef struct modcr_pack_desc {
  const int32_t* header;
  const int32_t* tokens;
  const float* feats;
  int32_t N, T, R, D;
  int64_t* input_ids;
  int64_t* token_type_ids;
  int64_t* total_label;
  int64_t* align_pos;
  int64_t* label;
  float* label_f;
  float* input_mask;
  int32_t* chunk_id;
  uint32_t* chunk_bits;
  int32_t cwords;
  float* chunk_mask;
  float* img_feat;
} modcr_pack_desc;
int modcr_pack_batch(const modcr_pack_desc* d, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Small utilities of the path. */
int modcr_cast_pad(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int64_t rows, int32_t cols,
                   int64_t row_step, void* stream);
int modcr_bf16_to_f32(const void* src, float* dst, int64_t n, int32_t accumulate, void* stream);
int modcr_colsum(const void* x, int64_t ld, int32_t M, int32_t N, float* out, void* stream);
int modcr_copy_rows(const void* src, int64_t src_seq_stride, int32_t srow0, void* dst, int64_t dst_seq_stride,
                    int32_t drow0, int32_t nseq, int32_t nrows, int32_t H, int32_t accumulate, void* stream);
int modcr_add_bf16(const void* a, const void* b, void* y, int64_t n, void* stream);
int modcr_tanh_bwd(const void* dy, const void* y, void* dx, int64_t n, void* stream);
/* 'mlp' classifier of ImageBertForMultipleChoice (modeling/modeling_bert.py:511-528: Linear -> ReLU [-> BertLayerNorm] ->
 * Linear): the ReLU runs in the modcr_gemm epilogue (act = 3); relu_bwd is dx = dy * (y > 0) on the bf16 ReLU output;
 * rowln_* is the LayerNorm over hidden*cls_hidden_scale columns of the few [questions] rows (any H % 8 == 0; bf16 in/out,
 * fp32 statistics; dgamma / dbeta accumulate). */
int modcr_relu_bwd(const void* dy, const void* y, void* dx, int64_t n, void* stream);
int modcr_rowln_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int32_t M,
                    int32_t H, float eps, void* stream);
int modcr_rowln_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, void* dx,
                    float* dgamma, float* dbeta, int32_t M, int32_t H, void* stream);
/* Box-position + token-type part of the region embedding of BertImgModel_gtToken (modeling/modeling_bert.py:312-316,
 * 390-394):  out[r,:] = img_token_type_embeddings[tt[r]] + LayerNorm_pos(img_pos_linear(pos[r,0:7])), fp32 [rows,H].
 * K = 7 is no tensor-core shape (one warp per row, fp32); `out` is then the fp32 epilogue input (in_op 1) of the 2054 -> H
 * feature-projection GEMM, so the three-way sum of :393 happens in that GEMM's epilogue.  mean / rstd [rows] are the
 * statistics of LayerNorm_pos for the backward pass, which takes d = gradient of the sum (bf16 [rows,H]) and accumulates
 * dW_pos [H,7], db_pos, dgamma_pos, dbeta_pos [H], dtype_emb [2,H] (all fp32, pre-zeroed by the caller). */
int modcr_pos_type_embed_fwd(const float* pos, const int64_t* tt, const float* W_pos, const float* b_pos,
                             const float* gamma_pos, const float* beta_pos, const float* type_emb, float* out,
                             float* mean, float* rstd, int32_t rows, int32_t H, float eps, void* stream);
int modcr_pos_type_embed_bwd(const void* d, const float* pos, const int64_t* tt, const float* W_pos, const float* b_pos,
                             const float* gamma_pos, const float* mean, const float* rstd, float* dW_pos, float* db_pos,
                             float* dgamma_pos, float* dbeta_pos, float* dtype_emb, int32_t rows, int32_t H, void* stream);

/* Rows that enter the encoder already embedded (the prompt embeddings of the language-model side,
 * modeling/modeling_ensemble.py:498-503 `prompt_embeddings=`): rows_place writes fp32 rows src [M,H] into the bf16
 * hidden-state buffer y (and its fp32 twin y32) at the physical rows given by ymap; rows_gather is its adjoint on the
 * bf16 gradient (dst [M,H] bf16). */
int modcr_rows_place(const float* src, void* y, float* y32, modcr_rowmap ymap, int32_t M, int32_t H, void* stream);
int modcr_rows_gather(const void* dy, modcr_rowmap dymap, void* dst, int32_t M, int32_t H, void* stream);
int modcr_key_bias(const float* mask, float* kbias, int64_t total, void* stream);
int modcr_chunk_bits(const float* chunk_mask, uint32_t* bits, int32_t N, int32_t T, int32_t cwords, void* stream);
int modcr_accum_f32_2d(float* dst, int64_t ld_dst, const float* src, int64_t ld_src, int64_t rows, int32_t cols,
                       int32_t accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MODCR_B200_H_ */
