// Internal C++ interface of the HBM-bound row kernels (rowops.cu) and head kernels (heads.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace modcr {

// Logical row m -> physical row (m / rows_per_seq) * seq_stride + row0 + m % rows_per_seq.
// rows_per_seq <= 0 means identity.
struct RowMap {
  int rows_per_seq = 0;
  int seq_stride = 0;
  int row0 = 0;
};

// x is bf16 (x_f32 = 0) or fp32 (x_f32 = 1, the fp32 residual stream); y32 (optional) receives an fp32 copy of y
int layernorm_fwd(const void* x, int x_f32, RowMap xmap, const float* gamma, const float* beta, void* y, float* y32,
                  RowMap ymap, float* mean, float* rstd, int M, int H, float eps, cudaStream_t st);
int layernorm_bwd(const void* dy, RowMap dymap, const void* x, int x_f32, RowMap xmap, const float* mean,
                  const float* rstd, const float* gamma, void* dx, RowMap dxmap, float* dgamma, float* dbeta, float* dbias,
                  int M, int H, void* dxm, const uint32_t* dseed, uint32_t dsite, float dp, cudaStream_t st);
int txt_embed_ln_fwd(const int64_t* ids, const int64_t* tts, const int64_t* pos_ids, const float* word, const float* type,
                     const float* pos, const float* gamma, const float* beta, void* y, float* y32, RowMap ymap,
                     float* mean, float* rstd, int N, int T, int H, float eps, cudaStream_t st);
int txt_embed_ln_bwd(const void* dy, RowMap dymap, const int64_t* ids, const int64_t* tts, const int64_t* pos_ids,
                     const float* word, const float* type, const float* pos, const float* gamma, const float* mean,
                     const float* rstd, float* dword, float* dtype, float* dpos, float* dgamma, float* dbeta, int N, int T,
                     int H, int64_t padding_idx, cudaStream_t st);
int cast_pad(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int64_t rows, int cols, int64_t row_step,
             cudaStream_t st);
int bf16_to_f32(const void* src, float* dst, int64_t n, int accumulate, cudaStream_t st);
int accum_f32_2d(float* dst, int64_t ld_dst, const float* src, int64_t ld_src, int64_t rows, int cols, int accumulate,
                 cudaStream_t st);
int colsum(const void* x, int64_t ld, int M, int N, float* out, cudaStream_t st);
int copy_rows(const void* src, int64_t src_seq_stride, int srow0, void* dst, int64_t dst_seq_stride, int drow0, int nseq,
              int nrows, int H, int accumulate, cudaStream_t st);
int add_bf16(const void* a, const void* b, void* y, int64_t n, cudaStream_t st);
int tanh_bwd(const void* dy, const void* y, void* dx, int64_t n, cudaStream_t st);
// box-position projection (7 -> H) + LayerNorm + type embedding of BertImgModel_gtToken (modeling_bert.py:312-316, 390-394)
int pos_type_fwd(const float* pos, const int64_t* tt, const float* Wp, const float* bp, const float* gam, const float* bet,
                 const float* type_emb, float* out, float* mean, float* rstd, int rows, int H, float eps, cudaStream_t st);
int pos_type_bwd(const void* d, const float* pos, const int64_t* tt, const float* Wp, const float* bp, const float* gam,
                 const float* mean, const float* rstd, float* dWp, float* dbp, float* dgam, float* dbet, float* dtype,
                 int rows, int H, cudaStream_t st);
int rows_place(const float* src, void* y, float* y32, RowMap ymap, int M, int H, cudaStream_t st);
int rows_gather(const void* dy, RowMap dymap, void* dst, int M, int H, cudaStream_t st);
int relu_bwd(const void* dy, const void* y, void* dx, int64_t n, cudaStream_t st);
int rowln_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int M, int H,
              float eps, cudaStream_t st);
int rowln_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, void* dx,
              float* dgamma, float* dbeta, int M, int H, cudaStream_t st);
int key_bias(const float* mask, float* kbias, int64_t total, cudaStream_t st);
int chunk_bits(const float* cmask, uint32_t* bits, int N, int T, int cwords, cudaStream_t st);
int chunk_mean_q(void* q, int64_t ld, int N, int S, int H, const int32_t* chunk_id, int T, int max_chunks,
                 cudaStream_t st);

// ---- heads.cu
// Alignment loss (v10:982-987): A = sum over `nslices` slices of probs_img, masked_fill(A==0,-1e5),
// softmax over R, rows with align_pos==1, CE(probabilities, total_label) -- double softmax as written.
//   probs: fp32 [nl, N, heads, T, R]; block: [N,T,R]; row_loss: [N*T, 2]; out: [4] = {loss, M, correct, loss sum}
int align_softmax_ce_fwd(const float* probs, int nl, int heads, const int64_t* align_pos, const int64_t* total_label,
                         const int64_t* row_sel, float* block, float* row_loss, float* out, int N, int T, int R,
                         cudaStream_t st);
int align_softmax_ce_bwd(const float* block, const int64_t* align_pos, const int64_t* total_label, const float* out,
                         float gout, float* dblock, int N, int T, int R, cudaStream_t st);

// ClsLayer2 attention pooling (v10:816-829): w = softmax(q k^T + mask) (no 1/sqrt(d)), pooled = w k
// q [N,H] and k [N,L,H] are bf16 (qk_f32 = 0) or fp32 (qk_f32 = 1)
// fused multi-tensor gradient-norm + AdamW (csrc/optim.cu); `meta` = device array of OptTensor, `grads` = device array
// of gradient pointers (NULL entries are skipped), one block per (tensor, 8192-element chunk) of the block table
int grad_sqnorm(const void* meta, const void* grads, const int32_t* blk_tensor, const int32_t* blk_chunk, int nblocks,
                float* out, cudaStream_t st);
int grad_clip(const void* meta, const void* grads, const int32_t* blk_tensor, const int32_t* blk_chunk, int nblocks,
              const float* grad_sq, float max_norm, cudaStream_t st);
int adamw_step(const void* meta, const void* grads, const int32_t* blk_tensor, const int32_t* blk_chunk, int nblocks,
               const float* hyper, const float* grad_sq, float b1, float b2, float eps, cudaStream_t st);
// dst[i][0..n[i]) = scale * src[i][0..n[i]) for up to 64 tensors per launch, fp32 <-> bf16 (gradient buckets, csrc/optim.cu);
// src / dst / n are host arrays
int multi_copy(const void* const* src, void* const* dst, const int64_t* n, int count, int src_bf16, int dst_bf16, float scale,
               cudaStream_t st);
// training-mode dropout with a counter-based mask (seed read from device memory): out16 = in16 * m, io32 = io32 * m (+ res32)
int dropout(const void* in16, void* out16, float* io32, const float* res32, int64_t n, const uint32_t* seed, uint32_t site,
            float p, cudaStream_t st);
// multi-head attention of one query per sequence over L key/value rows (cross_attention_lyx, v10:692-797)
int cross_attn_q1_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int N, int L, int H, int heads,
                      const uint32_t* dseed, uint32_t dsite, float dp, cudaStream_t st);
int cross_attn_q1_bwd(const void* q, const void* k, const void* v, const void* out, const float* lse, const void* dout,
                      void* dq, void* dk, void* dv, int N, int L, int H, int heads, const uint32_t* dseed, uint32_t dsite,
                      float dp, cudaStream_t st);
int cls_pool_attn_fwd(const void* q, const void* k, int qk_f32, const float* mask, float* w, void* pooled, int N, int L,
                      int H, const uint32_t* dseed, uint32_t dsite, float dp, cudaStream_t st);
int cls_pool_attn_bwd(const void* dpooled, const void* q, const void* k, int qk_f32, const float* w, void* dq, void* dk,
                      int N, int L, int H, const uint32_t* dseed, uint32_t dsite, float dp, cudaStream_t st);

// Fused small classifier + cross-entropy (+ binary_to_mp / argmax / matched):
//   logits[r, c] = x[r,:] . W[c,:] + b[c]            x bf16 [rows,K], W fp32 [C,K], C <= 8
//   CE over the view [rows*C/G, G]; hard labels (int64) or soft labels (fp32 [rows*C/G, G]).
struct ScorerDesc {
  const void* x = nullptr;
  const float* W = nullptr;
  const float* b = nullptr;
  int rows = 0, K = 0, C = 0, G = 0;
  const int64_t* hard = nullptr;
  const float* soft = nullptr;
  float* logits = nullptr;     // [rows, C]
  float* loss = nullptr;       // [1]
  float* mp = nullptr;         // optional [rows/num_choice, num_choice] = softmax(logits)[:,1] (C == 2)
  int64_t* pred = nullptr;     // optional [rows/num_choice] argmax of mp
  uint8_t* matched = nullptr;  // optional [rows/num_choice] pred == argmax(label group)
  int num_choice = 4;
  // backward
  float gout = 1.f;
  void* dx = nullptr;          // bf16 [rows, K]
  float* dW = nullptr;         // fp32 [C, K] accumulate
  float* db = nullptr;         // fp32 [C] accumulate
  int loss_kind = 0;           // 0 CE, 1 BCE-with-logits x G (instance_bce_with_logits)
};
int mc_scorer_ce_fwd(const ScorerDesc& d, cudaStream_t st);
int mc_scorer_ce_bwd(const ScorerDesc& d, cudaStream_t st);

// device-side batch packer (pack.cu; SURVEY 8 f2, Data/VCRChunkAlign.py:655-741)
struct PackDesc {
  const int32_t* header = nullptr;
  const int32_t* tokens = nullptr;
  const float* feats = nullptr;
  int N = 0, T = 0, R = 0, D = 0;
  int64_t* input_ids = nullptr;
  int64_t* token_type_ids = nullptr;
  int64_t* total_label = nullptr;
  int64_t* align_pos = nullptr;
  int64_t* label = nullptr;
  float* label_f = nullptr;
  float* input_mask = nullptr;
  int32_t* chunk_id = nullptr;
  uint32_t* chunk_bits = nullptr;
  int cwords = 0;
  float* chunk_mask = nullptr;
  float* img_feat = nullptr;
};
int pack_batch(const PackDesc& d, cudaStream_t st);

}  // namespace modcr
