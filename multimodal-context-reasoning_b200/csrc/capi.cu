// C ABI of libmodcr_b200.so (declared in include/modcr_b200.h): thin, validating wrappers that
// forward to the kernels' C++ launchers.  No torch types, no allocation, no exceptions.
#include "../../include/modcr_b200.h"

#include <cuda_runtime.h>

#include "attn.h"
#include "gemm.h"
#include "rowops.h"
#include "tma_host.h"

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline modcr::RowMap RM(modcr_rowmap m) {
  modcr::RowMap r;
  r.rows_per_seq = m.rows_per_seq;
  r.seq_stride = m.seq_stride;
  r.row0 = m.row0;
  return r;
}

// every compute entry point refuses to run on anything but an sm_100 device (no fallback)
static int check_arch() {
  static int cached = 1;  // 1 = unknown, 0 = ok, <0 = error
  if (cached <= 0 && cached != 1) return cached;
  if (cached == 0) return 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    modcr::set_error("no CUDA device: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return modcr::MODCR_ERR_ARCH;
  }
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) {
    modcr::set_error("libmodcr_b200 needs an sm_100 (Blackwell B200) device, found compute capability %d.x", major);
    return modcr::MODCR_ERR_ARCH;
  }
  cached = 0;
  return 0;
}
#define ARCH_OR_RETURN()       \
  do {                         \
    int rc_ = check_arch();    \
    if (rc_) return rc_;       \
  } while (0)

extern "C" {

int modcr_version(void) { return MODCR_VERSION; }
const char* modcr_last_error_string(void) { return modcr::last_error(); }
uint64_t modcr_launch_count(void) { return (uint64_t)modcr::launch_count(); }

size_t modcr_workspace_bytes(int op, const int64_t* dims, int ndims) {
  if (op == 1 && ndims >= 3) return modcr::attn_bwd_workspace_bytes((int)dims[0], (int)dims[1], (int)dims[2]);
  if (op == 2 && ndims >= 3) {
    const int64_t rows = dims[0], D = dims[1], H = dims[2];
    const int64_t Dpad = (D + 63) / 64 * 64;
    return (size_t)(rows * Dpad * 2 + rows * H * 2);
  }
  return 0;
}

int modcr_gemm(const modcr_gemm_desc* d, void* stream) {
  ARCH_OR_RETURN();
  if (d == nullptr) {
    modcr::set_error("modcr_gemm: null descriptor");
    return modcr::MODCR_ERR_SHAPE;
  }
  modcr::GemmDesc g;
  g.A = d->A; g.B = d->B; g.C = d->C;
  g.lda = d->lda; g.ldb = d->ldb; g.ldc = d->ldc;
  g.a_mn = d->a_mn; g.b_mn = d->b_mn; g.c_fp32 = d->c_fp32; g.accumulate = d->accumulate; g.split_k = d->split_k;
  g.M = d->M; g.N = d->N; g.K = d->K;
  g.alpha = d->alpha; g.bias = d->bias; g.act = d->act;
  g.aux = d->aux; g.ldaux = d->ldaux; g.epi_in = d->epi_in; g.ld_in = d->ld_in; g.in_op = d->in_op;
  g.a_colsum = d->a_colsum;
  g.drop_seed = d->drop_seed; g.drop_site = d->drop_site; g.drop_p = d->drop_p;
  g.in_mean = d->in_mean; g.in_rstd = d->in_rstd; g.in_gamma = d->in_gamma; g.in_beta = d->in_beta;
  return modcr::gemm_launch(g, S(stream));
}

int modcr_layernorm_fwd(const void* x, int32_t x_f32, modcr_rowmap xmap, const float* gamma, const float* beta, void* y,
                        float* y32, modcr_rowmap ymap, float* mean, float* rstd, int32_t M, int32_t H, float eps,
                        void* stream) {
  ARCH_OR_RETURN();
  return modcr::layernorm_fwd(x, x_f32, RM(xmap), gamma, beta, y, y32, RM(ymap), mean, rstd, M, H, eps, S(stream));
}
int modcr_layernorm_bwd(const void* dy, modcr_rowmap dymap, const void* x, int32_t x_f32, modcr_rowmap xmap,
                        const float* mean, const float* rstd, const float* gamma, void* dx, modcr_rowmap dxmap,
                        float* dgamma, float* dbeta, float* dbias, int32_t M, int32_t H, void* dx_masked,
                        const uint32_t* drop_seed, uint32_t drop_site, float drop_p, void* stream) {
  ARCH_OR_RETURN();
  return modcr::layernorm_bwd(dy, RM(dymap), x, x_f32, RM(xmap), mean, rstd, gamma, dx, RM(dxmap), dgamma, dbeta, dbias,
                              M, H, dx_masked, drop_seed, drop_site, drop_p, S(stream));
}

int modcr_txt_embed_ln_fwd(const int64_t* input_ids, const int64_t* token_type_ids, const int64_t* position_ids,
                           const float* word, const float* type, const float* pos, const float* gamma,
                           const float* beta, void* y, float* y32, modcr_rowmap ymap, float* mean, float* rstd,
                           int32_t N, int32_t T, int32_t H, float eps, void* stream) {
  ARCH_OR_RETURN();
  return modcr::txt_embed_ln_fwd(input_ids, token_type_ids, position_ids, word, type, pos, gamma, beta, y, y32, RM(ymap),
                                 mean, rstd, N, T, H, eps, S(stream));
}
int modcr_txt_embed_ln_bwd(const void* dy, modcr_rowmap dymap, const int64_t* input_ids,
                           const int64_t* token_type_ids, const int64_t* position_ids, const float* word,
                           const float* type, const float* pos, const float* gamma, const float* mean,
                           const float* rstd, float* dword, float* dtype, float* dpos, float* dgamma, float* dbeta,
                           int32_t N, int32_t T, int32_t H, int64_t padding_idx, void* stream) {
  ARCH_OR_RETURN();
  return modcr::txt_embed_ln_bwd(dy, RM(dymap), input_ids, token_type_ids, position_ids, word, type, pos, gamma, mean,
                                 rstd, dword, dtype, dpos, dgamma, dbeta, N, T, H, padding_idx, S(stream));
}

int modcr_img_embed_ln_fwd(const float* feats, int64_t rows, int32_t D, int32_t Dpad, const void* W,
                           const float* bias, const float* gamma, const float* beta, int32_t use_ln, float eps,
                           void* packed, float* pre, void* y, float* y32, modcr_rowmap ymap, float* mean, float* rstd,
                           int32_t H, const float* extra32, void* stream) {
  ARCH_OR_RETURN();
  if (rows <= 0 || D <= 0 || Dpad < D || Dpad % 8 != 0 || rows > 0x7fffffff) {
    modcr::set_error("img_embed_ln_fwd: bad shape rows=%lld D=%d Dpad=%d", (long long)rows, D, Dpad);
    return modcr::MODCR_ERR_SHAPE;
  }
  int rc = modcr::cast_pad(feats, D, packed, Dpad, rows, D, 1, S(stream));
  if (rc) return rc;
  modcr::GemmDesc g;
  g.A = packed; g.lda = Dpad;
  g.B = W; g.ldb = Dpad;
  g.M = (int)rows; g.N = H; g.K = Dpad;
  g.bias = bias;
  g.C = pre; g.ldc = H; g.c_fp32 = 1;   // fp32 pre-LayerNorm activations
  if (extra32 != nullptr) {             // + type / box-position term, summed in the epilogue
    g.epi_in = extra32; g.ld_in = H; g.in_op = 1;
  }
  rc = modcr::gemm_launch(g, S(stream));
  if (rc) return rc;
  if (!use_ln)  // config without use_img_layernorm (modeling_transfomres.py:676-680): the projection is the embedding
    return modcr::rows_place(pre, y, y32, RM(ymap), (int)rows, H, S(stream));
  modcr::RowMap id;
  return modcr::layernorm_fwd(pre, 1, id, gamma, beta, y, y32, RM(ymap), mean, rstd, (int)rows, H, eps, S(stream));
}

int modcr_img_embed_ln_bwd(const void* dy, modcr_rowmap dymap, const void* packed, const float* pre,
                           const float* mean, const float* rstd, const float* gamma, int32_t use_ln, void* dpre,
                           float* dW_pad, float* dbias, float* dgamma, float* dbeta, int64_t rows, int32_t Dpad,
                           int32_t H, void* stream) {
  ARCH_OR_RETURN();
  if (rows <= 0 || Dpad % 8 != 0 || rows > 0x7fffffff) {
    modcr::set_error("img_embed_ln_bwd: bad shape rows=%lld Dpad=%d", (long long)rows, Dpad);
    return modcr::MODCR_ERR_SHAPE;
  }
  modcr::RowMap id;
  int rc;
  if (!use_ln) {
    rc = modcr::rows_gather(dy, RM(dymap), dpre, (int)rows, H, S(stream));
    if (rc) return rc;
    rc = modcr::colsum(dpre, H, (int)rows, H, dbias, S(stream));
  } else {
    rc = modcr::layernorm_bwd(dy, RM(dymap), pre, 1, id, mean, rstd, gamma, dpre, id, dgamma, dbeta, dbias, (int)rows, H,
                              nullptr, nullptr, 0, 0.f, S(stream));
  }
  if (rc) return rc;
  // dW[H, Dpad] = dpre^T packed      (A = dpre stored [K=rows, M=H], B = packed stored [K=rows, N=Dpad])
  modcr::GemmDesc g;
  g.A = dpre; g.lda = H; g.a_mn = 1;
  g.B = packed; g.ldb = Dpad; g.b_mn = 1;
  g.C = dW_pad; g.ldc = Dpad; g.c_fp32 = 1;
  g.M = H; g.N = Dpad; g.K = (int)rows;
  return modcr::gemm_launch(g, S(stream));
}

static void to_attn(const modcr_attn_desc* d, modcr::AttnDesc& a) {
  a.N = d->N; a.S = d->S; a.P = d->P; a.heads = d->heads; a.T = d->T; a.mode = d->mode; a.scale = d->scale;
  a.qkv = d->qkv; a.pkv = d->pkv; a.kbias = d->kbias; a.cbits = d->cbits; a.cwords = d->cwords;
  a.ctx = d->ctx; a.lse = d->lse; a.probs_img = d->probs_img; a.row_sel = d->row_sel;
  a.dctx = d->dctx; a.dalign = d->dalign; a.dqkv = d->dqkv; a.dpkv = d->dpkv; a.dq_ws = d->dq_ws;
  a.drop_seed = d->drop_seed; a.drop_site = d->drop_site; a.drop_p = d->drop_p;
}
int modcr_prefix_attn_fwd(const modcr_attn_desc* d, void* stream) {
  ARCH_OR_RETURN();
  if (d == nullptr) {
    modcr::set_error("modcr_prefix_attn_fwd: null descriptor");
    return modcr::MODCR_ERR_SHAPE;
  }
  modcr::AttnDesc a;
  to_attn(d, a);
  return modcr::attn_fwd_launch(a, S(stream));
}
int modcr_prefix_attn_bwd(const modcr_attn_desc* d, void* stream) {
  ARCH_OR_RETURN();
  if (d == nullptr) {
    modcr::set_error("modcr_prefix_attn_bwd: null descriptor");
    return modcr::MODCR_ERR_SHAPE;
  }
  modcr::AttnDesc a;
  to_attn(d, a);
  return modcr::attn_bwd_launch(a, S(stream));
}

int modcr_chunk_mean_q(void* q, int64_t ld, int32_t N, int32_t S_, int32_t H, const int32_t* chunk_id, int32_t T,
                       int32_t max_chunks, void* stream) {
  ARCH_OR_RETURN();
  return modcr::chunk_mean_q(q, ld, N, S_, H, chunk_id, T, max_chunks, S(stream));
}

int modcr_align_softmax_ce_fwd(const float* probs, int32_t nl, int32_t heads, const int64_t* align_pos,
                               const int64_t* total_label, const int64_t* row_sel, float* block, float* row_loss,
                               float* out, int32_t N, int32_t T, int32_t R, void* stream) {
  ARCH_OR_RETURN();
  return modcr::align_softmax_ce_fwd(probs, nl, heads, align_pos, total_label, row_sel, block, row_loss, out, N, T, R,
                                     S(stream));
}
int modcr_align_softmax_ce_bwd(const float* block, const int64_t* align_pos, const int64_t* total_label,
                               const float* out, float gout, float* dblock, int32_t N, int32_t T, int32_t R,
                               void* stream) {
  ARCH_OR_RETURN();
  return modcr::align_softmax_ce_bwd(block, align_pos, total_label, out, gout, dblock, N, T, R, S(stream));
}

int modcr_grad_sqnorm(const modcr_optim_tensor* meta, const float* const* grads, const int32_t* blk_tensor,
                       const int32_t* blk_chunk, int32_t nblocks, float* out_sq, void* stream) {
  ARCH_OR_RETURN();
  return modcr::grad_sqnorm(meta, grads, blk_tensor, blk_chunk, nblocks, out_sq, S(stream));
}
int modcr_grad_clip(const modcr_optim_tensor* meta, float* const* grads, const int32_t* blk_tensor,
                    const int32_t* blk_chunk, int32_t nblocks, const float* grad_sq, float max_norm, void* stream) {
  ARCH_OR_RETURN();
  return modcr::grad_clip(meta, grads, blk_tensor, blk_chunk, nblocks, grad_sq, max_norm, S(stream));
}
int modcr_adamw_step(const modcr_optim_tensor* meta, const float* const* grads, const int32_t* blk_tensor,
                     const int32_t* blk_chunk, int32_t nblocks, const float* hyper, const float* grad_sq, float beta1,
                     float beta2, float eps, void* stream) {
  ARCH_OR_RETURN();
  return modcr::adamw_step(meta, grads, blk_tensor, blk_chunk, nblocks, hyper, grad_sq, beta1, beta2, eps, S(stream));
}

int modcr_set_sm_limit(int32_t sms) { return modcr::set_sm_limit(sms); }

int modcr_multi_copy(const void* const* src, void* const* dst, const int64_t* n, int32_t count, int32_t src_bf16,
                     int32_t dst_bf16, float scale, void* stream) {
  ARCH_OR_RETURN();
  return modcr::multi_copy(src, dst, n, count, src_bf16, dst_bf16, scale, S(stream));
}

int modcr_dropout(const void* in16, void* out16, float* io32, const float* res32, int64_t n, const uint32_t* seed,
                  uint32_t site, float p, void* stream) {
  ARCH_OR_RETURN();
  return modcr::dropout(in16, out16, io32, res32, n, seed, site, p, S(stream));
}

int modcr_cross_attn_q1_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int32_t N, int32_t L,
                            int32_t H, int32_t heads, const uint32_t* drop_seed, uint32_t drop_site, float drop_p,
                            void* stream) {
  ARCH_OR_RETURN();
  return modcr::cross_attn_q1_fwd(q, k, v, out, lse, N, L, H, heads, drop_seed, drop_site, drop_p, S(stream));
}
int modcr_cross_attn_q1_bwd(const void* q, const void* k, const void* v, const void* out, const float* lse,
                            const void* dout, void* dq, void* dk, void* dv, int32_t N, int32_t L, int32_t H,
                            int32_t heads, const uint32_t* drop_seed, uint32_t drop_site, float drop_p, void* stream) {
  ARCH_OR_RETURN();
  return modcr::cross_attn_q1_bwd(q, k, v, out, lse, dout, dq, dk, dv, N, L, H, heads, drop_seed, drop_site, drop_p,
                                  S(stream));
}

int modcr_cls_pool_attn_fwd(const void* q, const void* k, int32_t qk_f32, const float* mask, float* w, void* pooled,
                            int32_t N, int32_t L, int32_t H, const uint32_t* drop_seed, uint32_t drop_site, float drop_p,
                            void* stream) {
  ARCH_OR_RETURN();
  return modcr::cls_pool_attn_fwd(q, k, qk_f32, mask, w, pooled, N, L, H, drop_seed, drop_site, drop_p, S(stream));
}
int modcr_cls_pool_attn_bwd(const void* dpooled, const void* q, const void* k, int32_t qk_f32, const float* w, void* dq,
                            void* dk, int32_t N, int32_t L, int32_t H, const uint32_t* drop_seed, uint32_t drop_site,
                            float drop_p, void* stream) {
  ARCH_OR_RETURN();
  return modcr::cls_pool_attn_bwd(dpooled, q, k, qk_f32, w, dq, dk, N, L, H, drop_seed, drop_site, drop_p, S(stream));
}

static void to_scorer(const modcr_scorer_desc* d, modcr::ScorerDesc& s) {
  s.x = d->x; s.W = d->W; s.b = d->b; s.rows = d->rows; s.K = d->K; s.C = d->C; s.G = d->G;
  s.hard = d->hard; s.soft = d->soft; s.logits = d->logits; s.loss = d->loss; s.mp = d->mp; s.pred = d->pred;
  s.matched = d->matched; s.num_choice = d->num_choice; s.gout = d->gout; s.dx = d->dx; s.dW = d->dW; s.db = d->db;
  s.loss_kind = d->loss_kind;
}
int modcr_mc_scorer_ce_fwd(const modcr_scorer_desc* d, void* stream) {
  ARCH_OR_RETURN();
  if (d == nullptr) {
    modcr::set_error("modcr_mc_scorer_ce_fwd: null descriptor");
    return modcr::MODCR_ERR_SHAPE;
  }
  modcr::ScorerDesc s;
  to_scorer(d, s);
  return modcr::mc_scorer_ce_fwd(s, S(stream));
}
int modcr_mc_scorer_ce_bwd(const modcr_scorer_desc* d, void* stream) {
  ARCH_OR_RETURN();
  if (d == nullptr) {
    modcr::set_error("modcr_mc_scorer_ce_bwd: null descriptor");
    return modcr::MODCR_ERR_SHAPE;
  }
  modcr::ScorerDesc s;
  to_scorer(d, s);
  return modcr::mc_scorer_ce_bwd(s, S(stream));
}

int modcr_pack_batch(const modcr_pack_desc* d, void* stream) {
  ARCH_OR_RETURN();
  if (d == nullptr) {
    modcr::set_error("modcr_pack_batch: null descriptor");
    return modcr::MODCR_ERR_SHAPE;
  }
  modcr::PackDesc s;
  s.header = d->header; s.tokens = d->tokens; s.feats = d->feats; s.N = d->N; s.T = d->T; s.R = d->R; s.D = d->D;
  s.input_ids = d->input_ids; s.token_type_ids = d->token_type_ids; s.total_label = d->total_label;
  s.align_pos = d->align_pos; s.label = d->label; s.label_f = d->label_f; s.input_mask = d->input_mask;
  s.chunk_id = d->chunk_id; s.chunk_bits = d->chunk_bits; s.cwords = d->cwords; s.chunk_mask = d->chunk_mask;
  s.img_feat = d->img_feat;
  return modcr::pack_batch(s, S(stream));
}

int modcr_cast_pad(const float* src, int64_t ld_src, void* dst, int64_t ld_dst, int64_t rows, int32_t cols,
                   int64_t row_step, void* stream) {
  ARCH_OR_RETURN();
  return modcr::cast_pad(src, ld_src, dst, ld_dst, rows, cols, row_step, S(stream));
}
int modcr_bf16_to_f32(const void* src, float* dst, int64_t n, int32_t accumulate, void* stream) {
  ARCH_OR_RETURN();
  return modcr::bf16_to_f32(src, dst, n, accumulate, S(stream));
}
int modcr_colsum(const void* x, int64_t ld, int32_t M, int32_t N, float* out, void* stream) {
  ARCH_OR_RETURN();
  return modcr::colsum(x, ld, M, N, out, S(stream));
}
int modcr_copy_rows(const void* src, int64_t src_seq_stride, int32_t srow0, void* dst, int64_t dst_seq_stride,
                    int32_t drow0, int32_t nseq, int32_t nrows, int32_t H, int32_t accumulate, void* stream) {
  ARCH_OR_RETURN();
  return modcr::copy_rows(src, src_seq_stride, srow0, dst, dst_seq_stride, drow0, nseq, nrows, H, accumulate, S(stream));
}
int modcr_add_bf16(const void* a, const void* b, void* y, int64_t n, void* stream) {
  ARCH_OR_RETURN();
  return modcr::add_bf16(a, b, y, n, S(stream));
}
int modcr_relu_bwd(const void* dy, const void* y, void* dx, int64_t n, void* stream) {
  ARCH_OR_RETURN();
  return modcr::relu_bwd(dy, y, dx, n, S(stream));
}
int modcr_rowln_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int32_t M,
                    int32_t H, float eps, void* stream) {
  ARCH_OR_RETURN();
  return modcr::rowln_fwd(x, gamma, beta, y, mean, rstd, M, H, eps, S(stream));
}
int modcr_rowln_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, void* dx,
                    float* dgamma, float* dbeta, int32_t M, int32_t H, void* stream) {
  ARCH_OR_RETURN();
  return modcr::rowln_bwd(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, M, H, S(stream));
}
int modcr_tanh_bwd(const void* dy, const void* y, void* dx, int64_t n, void* stream) {
  ARCH_OR_RETURN();
  return modcr::tanh_bwd(dy, y, dx, n, S(stream));
}
int modcr_pos_type_embed_fwd(const float* pos, const int64_t* tt, const float* W_pos, const float* b_pos,
                             const float* gamma_pos, const float* beta_pos, const float* type_emb, float* out,
                             float* mean, float* rstd, int32_t rows, int32_t H, float eps, void* stream) {
  ARCH_OR_RETURN();
  return modcr::pos_type_fwd(pos, tt, W_pos, b_pos, gamma_pos, beta_pos, type_emb, out, mean, rstd, rows, H, eps, S(stream));
}
int modcr_pos_type_embed_bwd(const void* d, const float* pos, const int64_t* tt, const float* W_pos, const float* b_pos,
                             const float* gamma_pos, const float* mean, const float* rstd, float* dW_pos, float* db_pos,
                             float* dgamma_pos, float* dbeta_pos, float* dtype_emb, int32_t rows, int32_t H, void* stream) {
  ARCH_OR_RETURN();
  return modcr::pos_type_bwd(d, pos, tt, W_pos, b_pos, gamma_pos, mean, rstd, dW_pos, db_pos, dgamma_pos, dbeta_pos, dtype_emb,
                             rows, H, S(stream));
}
int modcr_rows_place(const float* src, void* y, float* y32, modcr_rowmap ymap, int32_t M, int32_t H, void* stream) {
  ARCH_OR_RETURN();
  return modcr::rows_place(src, y, y32, RM(ymap), M, H, S(stream));
}
int modcr_rows_gather(const void* dy, modcr_rowmap dymap, void* dst, int32_t M, int32_t H, void* stream) {
  ARCH_OR_RETURN();
  return modcr::rows_gather(dy, RM(dymap), dst, M, H, S(stream));
}
int modcr_key_bias(const float* mask, float* kbias, int64_t total, void* stream) {
  ARCH_OR_RETURN();
  return modcr::key_bias(mask, kbias, total, S(stream));
}
int modcr_chunk_bits(const float* chunk_mask, uint32_t* bits, int32_t N, int32_t T, int32_t cwords, void* stream) {
  ARCH_OR_RETURN();
  return modcr::chunk_bits(chunk_mask, bits, N, T, cwords, S(stream));
}
int modcr_accum_f32_2d(float* dst, int64_t ld_dst, const float* src, int64_t ld_src, int64_t rows, int32_t cols,
                       int32_t accumulate, void* stream) {
  ARCH_OR_RETURN();
  return modcr::accum_f32_2d(dst, ld_dst, src, ld_src, rows, cols, accumulate, S(stream));
}

#ifdef MODCR_ATTN_TRACE
/* developer build only (make trace): not part of include/modcr_b200.h */
int modcr_debug_read_attn_trace(long long* host, int32_t max_entries) { return modcr::attn_trace_read(host, max_entries); }
#endif

}  // extern "C"
