"""Tensor-level wrappers over the C ABI (``include/modcr_b200.h``).

Each function takes CUDA tensors that the caller owns (PyTorch is only the memory
and stream plumbing), validates what the C side cannot see (dtype / device /
contiguity) and enqueues the kernel on the current stream.  There is no fallback
path: without the shared library or an sm_100 device these raise ``RuntimeError``.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import AttnDesc, GemmDesc, PackDesc, RowMap, ScorerDesc

BF16, F32, I64 = torch.bfloat16, torch.float32, torch.int64
IDENT = RowMap(0, 0, 0)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _chk(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError("modcr_b200: %s must be a CUDA tensor (no CPU path)" % name)
    if t.dtype != dtype:
        raise RuntimeError("modcr_b200: %s must be %s, got %s" % (name, dtype, t.dtype))


def rowmap(rows_per_seq=0, seq_stride=0, row0=0):
    return RowMap(rows_per_seq, seq_stride, row0)


# --------------------------------------------------------------------------- GEMM
def gemm(A, B, C_, M, N, K, lda=None, ldb=None, ldc=None, a_mn=0, b_mn=0, bias=None, act=0, aux=None,
         epi_in=None, in_op=0, accumulate=0, split_k=1, alpha=1.0, ld_in=None, ldaux=None, a_colsum=None, drop=None,
         in_ln=None):
    """C[M,N] = epi(alpha * A B^T); operands are raw 2-D views described by (ptr, ld).
    ``a_colsum`` (a_mn=1): fp32 [M] that receives += sum_k A[k,m] (the bias gradient of a wgrad).
    ``in_ln`` = (mean [M], rstd [M], gamma [N], beta [N]) with fp32 ``epi_in`` / ``C``: the residual is
    LayerNorm(epi_in), rebuilt in the epilogue (in_op 3)."""
    d = GemmDesc()
    d.A, d.B, d.C = A.data_ptr(), B.data_ptr(), C_.data_ptr()
    d.lda = lda if lda is not None else (M if a_mn else K)
    d.ldb = ldb if ldb is not None else (N if b_mn else K)
    d.ldc = ldc if ldc is not None else N
    d.a_mn, d.b_mn = a_mn, b_mn
    d.c_fp32 = 1 if C_.dtype == F32 else 0
    d.accumulate, d.split_k = accumulate, split_k
    d.M, d.N, d.K = M, N, K
    d.alpha = alpha
    d.bias = _p(bias)
    d.act = act
    d.aux = _p(aux)
    d.ldaux = ldaux if ldaux is not None else N
    d.epi_in = _p(epi_in)
    d.ld_in = ld_in if ld_in is not None else N
    d.in_op = in_op if epi_in is not None else 0
    if in_ln is not None:
        if epi_in is None or epi_in.dtype != F32 or C_.dtype != F32:
            raise RuntimeError("modcr_b200: gemm in_ln needs an fp32 epilogue input and an fp32 output")
        for t_, nm in zip(in_ln, ("in_mean", "in_rstd", "in_gamma", "in_beta")):
            _chk(t_, F32, nm)
        d.in_op = 3
        d.in_mean, d.in_rstd, d.in_gamma, d.in_beta = (_p(t_) for t_ in in_ln)
    d.a_colsum = _p(a_colsum)
    if drop is not None:                      # (seed tensor, site, p): dropout of f(acc + bias) before `epi_in`
        d.drop_seed, d.drop_site, d.drop_p = _p(drop[0]), int(drop[1]), float(drop[2])
    _lib.check(_lib.load().modcr_gemm(C.byref(d), _stream()), "gemm")
    return C_


# --------------------------------------------------------------------------- LayerNorm / embeddings
def layernorm_fwd(x, gamma, beta, y, mean, rstd, M, H, eps, xmap=IDENT, ymap=IDENT, y32=None):
    """x may be bf16 or fp32 (the fp32 residual stream); y is bf16, y32 an optional fp32 copy."""
    _lib.check(_lib.load().modcr_layernorm_fwd(_p(x), int(x.dtype == F32), xmap, _p(gamma), _p(beta), _p(y), _p(y32),
                                               ymap, _p(mean), _p(rstd), M, H, eps, _stream()), "layernorm_fwd")


def layernorm_bwd(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, dbias, M, H, dymap=IDENT, xmap=IDENT, dxmap=IDENT,
                  dx_masked=None, drop=None):
    """``dx_masked`` + ``drop`` = (seed, site, p): second output dx * mask / (1 - p) for the dense layer whose dropped
    output fed this LayerNorm; ``dbias`` then sums the masked gradient."""
    ds, dsite, dp = (None, 0, 0.0) if drop is None else (_p(drop[0]), int(drop[1]), float(drop[2]))
    _lib.check(_lib.load().modcr_layernorm_bwd(_p(dy), dymap, _p(x), int(x.dtype == F32), xmap, _p(mean), _p(rstd),
                                               _p(gamma), _p(dx), dxmap, _p(dgamma), _p(dbeta), _p(dbias), M, H,
                                               _p(dx_masked), ds, dsite, dp, _stream()), "layernorm_bwd")


def txt_embed_ln_fwd(ids, tts, pos_ids, word, type_, pos, gamma, beta, y, ymap, mean, rstd, N, T, H, eps, y32=None):
    _chk(ids, I64, "input_ids")
    _chk(tts, I64, "token_type_ids")
    _chk(pos_ids, I64, "position_ids")
    _lib.check(_lib.load().modcr_txt_embed_ln_fwd(_p(ids), _p(tts), _p(pos_ids), _p(word), _p(type_), _p(pos),
                                                  _p(gamma), _p(beta), _p(y), _p(y32), ymap, _p(mean), _p(rstd), N, T,
                                                  H, eps, _stream()), "txt_embed_ln_fwd")


def txt_embed_ln_bwd(dy, dymap, ids, tts, pos_ids, word, type_, pos, gamma, mean, rstd, dword, dtype, dpos, dgamma,
                     dbeta, N, T, H, padding_idx):
    _lib.check(_lib.load().modcr_txt_embed_ln_bwd(_p(dy), dymap, _p(ids), _p(tts), _p(pos_ids), _p(word), _p(type_),
                                                  _p(pos), _p(gamma), _p(mean), _p(rstd), _p(dword), _p(dtype),
                                                  _p(dpos), _p(dgamma), _p(dbeta), N, T, H, padding_idx, _stream()),
               "txt_embed_ln_bwd")


def img_embed_ln_fwd(feats, rows, D, Dpad, W, bias, gamma, beta, use_ln, eps, packed, pre, y, ymap, mean, rstd, H,
                     y32=None, extra32=None):
    _chk(feats, F32, "img_feats")
    _chk(pre, F32, "img_embed pre-LN workspace")
    _chk(extra32, F32, "img_embed extra term")
    _lib.check(_lib.load().modcr_img_embed_ln_fwd(_p(feats), rows, D, Dpad, _p(W), _p(bias), _p(gamma), _p(beta),
                                                  int(use_ln), eps, _p(packed), _p(pre), _p(y), _p(y32), ymap,
                                                  _p(mean), _p(rstd), H, _p(extra32), _stream()), "img_embed_ln_fwd")


def img_embed_ln_bwd(dy, dymap, packed, pre, mean, rstd, gamma, use_ln, dpre, dW_pad, dbias, dgamma, dbeta, rows,
                     Dpad, H):
    _lib.check(_lib.load().modcr_img_embed_ln_bwd(_p(dy), dymap, _p(packed), _p(pre), _p(mean), _p(rstd), _p(gamma),
                                                  int(use_ln), _p(dpre), _p(dW_pad), _p(dbias), _p(dgamma),
                                                  _p(dbeta), rows, Dpad, H, _stream()), "img_embed_ln_bwd")


# --------------------------------------------------------------------------- attention
def _attn_desc(N, S, P, heads, T, mode, qkv, pkv, kbias, cbits, cwords, ctx, lse, probs_img=None, dctx=None,
               dalign=None, dqkv=None, dpkv=None, dq_ws=None, drop=None, row_sel=None):
    d = AttnDesc()
    d.N, d.S, d.P, d.heads, d.T, d.mode = N, S, P, heads, T, mode
    d.scale = 0.125
    d.qkv, d.pkv, d.kbias, d.cbits, d.cwords = _p(qkv), _p(pkv), _p(kbias), _p(cbits), cwords
    d.ctx, d.lse, d.probs_img = _p(ctx), _p(lse), _p(probs_img)
    d.dctx, d.dalign, d.dqkv, d.dpkv, d.dq_ws = _p(dctx), _p(dalign), _p(dqkv), _p(dpkv), _p(dq_ws)
    if drop is not None:                      # (seed tensor, site, p): attention-probability dropout
        d.drop_seed, d.drop_site, d.drop_p = _p(drop[0]), int(drop[1]), float(drop[2])
    _chk(row_sel, I64, "row_sel (align_pos)")
    d.row_sel = _p(row_sel)
    return d


def prefix_attn_fwd(N, S, P, heads, T, mode, qkv, pkv, kbias, cbits, cwords, ctx, lse, probs_img=None, drop=None,
                    row_sel=None):
    """``row_sel`` (int64 [N,T], the batch's ``align_pos``): only rows with a non-zero entry are written to ``probs_img``."""
    d = _attn_desc(N, S, P, heads, T, mode, qkv, pkv, kbias, cbits, cwords, ctx, lse, probs_img, drop=drop,
                   row_sel=row_sel)
    _lib.check(_lib.load().modcr_prefix_attn_fwd(C.byref(d), _stream()), "prefix_attn_fwd")


def prefix_attn_bwd(N, S, P, heads, T, mode, qkv, pkv, kbias, cbits, cwords, ctx, lse, probs_img, dctx, dalign, dqkv,
                    dpkv, dq_ws, drop=None, row_sel=None):
    d = _attn_desc(N, S, P, heads, T, mode, qkv, pkv, kbias, cbits, cwords, ctx, lse, probs_img, dctx, dalign, dqkv,
                   dpkv, dq_ws, drop, row_sel)
    _lib.check(_lib.load().modcr_prefix_attn_bwd(C.byref(d), _stream()), "prefix_attn_bwd")


def attn_bwd_workspace_bytes(N, S, heads):
    dims = (C.c_int64 * 3)(N, S, heads)
    return int(_lib.load().modcr_workspace_bytes(1, dims, 3))


def chunk_mean_q(q, ld, N, S, H, chunk_id, T, max_chunks):
    _chk(chunk_id, torch.int32, "chunk_id")
    _lib.check(_lib.load().modcr_chunk_mean_q(_p(q), ld, N, S, H, _p(chunk_id), T, max_chunks, _stream()),
               "chunk_mean_q")


# --------------------------------------------------------------------------- heads
def align_softmax_ce_fwd(probs, nl, heads, align_pos, total_label, block, row_loss, out, N, T, R, row_sel=None):
    _chk(align_pos, I64, "align_pos")
    _chk(total_label, I64, "total_label")
    _chk(row_sel, I64, "row_sel (align_pos)")
    _lib.check(_lib.load().modcr_align_softmax_ce_fwd(_p(probs), nl, heads, _p(align_pos), _p(total_label), _p(row_sel),
                                                      _p(block), _p(row_loss), _p(out), N, T, R, _stream()),
               "align_softmax_ce_fwd")


def align_softmax_ce_bwd(block, align_pos, total_label, out, gout, dblock, N, T, R):
    _lib.check(_lib.load().modcr_align_softmax_ce_bwd(_p(block), _p(align_pos), _p(total_label), _p(out), gout,
                                                      _p(dblock), N, T, R, _stream()), "align_softmax_ce_bwd")


def grad_sqnorm(meta, grads, blk_tensor, blk_chunk, nblocks, out_sq):
    _lib.check(_lib.load().modcr_grad_sqnorm(_p(meta), _p(grads), _p(blk_tensor), _p(blk_chunk), nblocks, _p(out_sq),
                                             _stream()), "grad_sqnorm")


def grad_clip(meta, grads, blk_tensor, blk_chunk, nblocks, grad_sq, max_norm):
    _lib.check(_lib.load().modcr_grad_clip(_p(meta), _p(grads), _p(blk_tensor), _p(blk_chunk), nblocks, _p(grad_sq),
                                           float(max_norm), _stream()), "grad_clip")


def adamw_step(meta, grads, blk_tensor, blk_chunk, nblocks, hyper, grad_sq, beta1, beta2, eps):
    _lib.check(_lib.load().modcr_adamw_step(_p(meta), _p(grads), _p(blk_tensor), _p(blk_chunk), nblocks, _p(hyper),
                                            _p(grad_sq), beta1, beta2, eps, _stream()), "adamw_step")


def multi_copy(srcs, dsts, scale=1.0):
    """dst[i] = scale * src[i] for lists of contiguous CUDA tensors (fp32 or bf16, any mix of the two between the lists
    but uniform inside each); 64 tensors per launch."""
    if len(srcs) != len(dsts):
        raise ValueError("multi_copy: %d sources for %d destinations" % (len(srcs), len(dsts)))
    if not srcs:
        return
    sdt, ddt = srcs[0].dtype, dsts[0].dtype
    for s_, d_ in zip(srcs, dsts):
        if s_.dtype != sdt or d_.dtype != ddt or s_.dtype not in (F32, BF16) or d_.dtype not in (F32, BF16):
            raise RuntimeError("modcr_b200: multi_copy takes fp32 / bf16 tensors, one dtype per list")
        if not (s_.is_cuda and d_.is_cuda):
            raise RuntimeError("modcr_b200: multi_copy needs CUDA tensors (no CPU path)")
        if s_.numel() != d_.numel() or not s_.is_contiguous() or not d_.is_contiguous():
            raise RuntimeError("modcr_b200: multi_copy needs contiguous tensors of equal length")
    lib, st = _lib.load(), _stream()
    for i in range(0, len(srcs), 64):
        ss, dd = srcs[i:i + 64], dsts[i:i + 64]
        k = len(ss)
        sp = (C.c_void_p * k)(*[t.data_ptr() for t in ss])
        dp = (C.c_void_p * k)(*[t.data_ptr() for t in dd])
        nn = (C.c_int64 * k)(*[t.numel() for t in ss])
        _lib.check(lib.modcr_multi_copy(sp, dp, nn, k, int(sdt == BF16), int(ddt == BF16), float(scale), st), "multi_copy")


def dropout(in16, out16, io32, res32, n, seed, site, p):
    """out16 = in16 * m (bf16), io32 = io32 * m (+ res32) (fp32, in place); m = keep / (1 - p), mask = f(seed, site, i)."""
    _lib.check(_lib.load().modcr_dropout(_p(in16), _p(out16), _p(io32), _p(res32), n, _p(seed), site, float(p), _stream()),
               "dropout")


def _drop3(drop):
    """(seed tensor, site, p) -> ctypes arguments; ``None`` disables dropout."""
    if drop is None:
        return None, 0, 0.0
    return _p(drop[0]), int(drop[1]), float(drop[2])


def cross_attn_q1_fwd(q, k, v, out, N, L, H, heads, lse=None, drop=None):
    """cross_attention_lyx with one query per sequence (v10:692-797): bf16 q/out [N,H], k/v [N,L,H];
    ``lse`` fp32 [N,heads] (optional) feeds the backward pass; ``drop`` = (seed, site, p) of the probability dropout."""
    ds, dsite, dp = _drop3(drop)
    _lib.check(_lib.load().modcr_cross_attn_q1_fwd(_p(q), _p(k), _p(v), _p(out), _p(lse), N, L, H, heads, ds, dsite, dp,
                                                   _stream()), "cross_attn_q1_fwd")


def cross_attn_q1_bwd(q, k, v, out, lse, dout, dq, dk, dv, N, L, H, heads, drop=None):
    ds, dsite, dp = _drop3(drop)
    _lib.check(_lib.load().modcr_cross_attn_q1_bwd(_p(q), _p(k), _p(v), _p(out), _p(lse), _p(dout), _p(dq), _p(dk), _p(dv),
                                                   N, L, H, heads, ds, dsite, dp, _stream()), "cross_attn_q1_bwd")


def cls_pool_attn_fwd(q, k, mask, w, pooled, N, L, H, drop=None):
    if q.dtype != k.dtype:
        raise RuntimeError("modcr_b200: cls_pool q and k must share a dtype")
    ds, dsite, dp = _drop3(drop)
    _lib.check(_lib.load().modcr_cls_pool_attn_fwd(_p(q), _p(k), int(k.dtype == F32), _p(mask), _p(w), _p(pooled), N, L,
                                                   H, ds, dsite, dp, _stream()), "cls_pool_attn_fwd")


def cls_pool_attn_bwd(dpooled, q, k, w, dq, dk, N, L, H, drop=None):
    ds, dsite, dp = _drop3(drop)
    _lib.check(_lib.load().modcr_cls_pool_attn_bwd(_p(dpooled), _p(q), _p(k), int(k.dtype == F32), _p(w), _p(dq),
                                                   _p(dk), N, L, H, ds, dsite, dp, _stream()), "cls_pool_attn_bwd")


def _scorer_desc(x, W, b, rows, K, Cc, G, hard, soft, logits, loss, mp=None, pred=None, matched=None, num_choice=4,
                 gout=1.0, dx=None, dW=None, db=None, loss_kind=0):
    d = ScorerDesc()
    d.x, d.W, d.b = _p(x), _p(W), _p(b)
    d.rows, d.K, d.C, d.G = rows, K, Cc, G
    d.hard, d.soft, d.logits, d.loss = _p(hard), _p(soft), _p(logits), _p(loss)
    d.mp, d.pred, d.matched, d.num_choice = _p(mp), _p(pred), _p(matched), num_choice
    d.gout, d.dx, d.dW, d.db = gout, _p(dx), _p(dW), _p(db)
    d.loss_kind = loss_kind
    return d


def mc_scorer_ce_fwd(x, W, b, rows, K, Cc, G, hard, soft, logits, loss, mp=None, pred=None, matched=None,
                     num_choice=4, loss_kind=0):
    """``loss_kind`` 0: cross-entropy over the view [rows*C/G, G]; 1: ``instance_bce_with_logits`` against ``soft``."""
    _chk(hard, I64, "labels")
    _chk(soft, F32, "soft labels")
    d = _scorer_desc(x, W, b, rows, K, Cc, G, hard, soft, logits, loss, mp, pred, matched, num_choice,
                     loss_kind=loss_kind)
    _lib.check(_lib.load().modcr_mc_scorer_ce_fwd(C.byref(d), _stream()), "mc_scorer_ce_fwd")


def mc_scorer_ce_bwd(x, W, b, rows, K, Cc, G, hard, soft, logits, gout, dx, dW, db, loss_kind=0):
    d = _scorer_desc(x, W, b, rows, K, Cc, G, hard, soft, logits, None, gout=gout, dx=dx, dW=dW, db=db,
                     loss_kind=loss_kind)
    _lib.check(_lib.load().modcr_mc_scorer_ce_bwd(C.byref(d), _stream()), "mc_scorer_ce_bwd")


def pack_batch(header, tokens, feats, N, T, R, D, input_ids, input_mask, token_type_ids=None, total_label=None,
               align_pos=None, label=None, label_f=None, chunk_id=None, chunk_bits=None, chunk_mask=None, img_feat=None):
    """Device-side collate (Data/VCRChunkAlign.py:690-741): expands the packed descriptors into the padded batch."""
    for t, dt, nm in ((header, torch.int32, "header"), (tokens, torch.int32, "tokens"), (feats, F32, "feats"),
                      (input_ids, I64, "input_ids"), (token_type_ids, I64, "token_type_ids"),
                      (total_label, I64, "total_label"), (align_pos, I64, "align_pos"), (label, I64, "label"),
                      (label_f, F32, "label_f"), (input_mask, F32, "input_mask"), (chunk_id, torch.int32, "chunk_id"),
                      (chunk_bits, torch.int32, "chunk_bits"), (chunk_mask, F32, "chunk_mask"), (img_feat, F32, "img_feat")):
        _chk(t, dt, nm)
    d = PackDesc()
    d.header, d.tokens, d.feats = _p(header), _p(tokens), _p(feats)
    d.N, d.T, d.R, d.D = N, T, R, D
    d.input_ids, d.token_type_ids, d.total_label, d.align_pos = _p(input_ids), _p(token_type_ids), _p(total_label), _p(align_pos)
    d.label, d.label_f, d.input_mask, d.chunk_id = _p(label), _p(label_f), _p(input_mask), _p(chunk_id)
    d.chunk_bits, d.cwords, d.chunk_mask, d.img_feat = _p(chunk_bits), (T + 31) // 32, _p(chunk_mask), _p(img_feat)
    _lib.check(_lib.load().modcr_pack_batch(C.byref(d), _stream()), "pack_batch")


# --------------------------------------------------------------------------- utilities
def cast_pad(src, ld_src, dst, ld_dst, rows, cols, row_step=1):
    _chk(src, F32, "cast_pad src")
    _lib.check(_lib.load().modcr_cast_pad(_p(src), ld_src, _p(dst), ld_dst, rows, cols, row_step, _stream()),
               "cast_pad")


def to_bf16(src):
    """fp32 [rows, cols] (contiguous) -> new bf16 tensor of the same shape (cols % 8 == 0)."""
    src = src.contiguous()
    cols = src.shape[-1]
    rows = src.numel() // cols
    dst = torch.empty(src.shape, dtype=BF16, device=src.device)
    if cols % 8 == 0:
        cast_pad(src, cols, dst, cols, rows, cols)
    else:  # 1-D / odd shapes: flat view, numel % 8 == 0 required
        cast_pad(src, src.numel(), dst, src.numel(), 1, src.numel())
    return dst


def bf16_to_f32(src, dst, accumulate=False):
    _lib.check(_lib.load().modcr_bf16_to_f32(_p(src), _p(dst), src.numel(), int(accumulate), _stream()),
               "bf16_to_f32")


def colsum(x, ld, M, N, out):
    _lib.check(_lib.load().modcr_colsum(_p(x), ld, M, N, _p(out), _stream()), "colsum")


def copy_rows(src, src_seq_stride, srow0, dst, dst_seq_stride, drow0, nseq, nrows, H, accumulate=False):
    _lib.check(_lib.load().modcr_copy_rows(_p(src), src_seq_stride, srow0, _p(dst), dst_seq_stride, drow0, nseq,
                                           nrows, H, int(accumulate), _stream()), "copy_rows")


def add_bf16(a, b, y):
    _lib.check(_lib.load().modcr_add_bf16(_p(a), _p(b), _p(y), a.numel(), _stream()), "add_bf16")


def tanh_bwd(dy, y, dx):
    _lib.check(_lib.load().modcr_tanh_bwd(_p(dy), _p(y), _p(dx), dy.numel(), _stream()), "tanh_bwd")


def relu_bwd(dy, y, dx):
    _lib.check(_lib.load().modcr_relu_bwd(_p(dy), _p(y), _p(dx), dy.numel(), _stream()), "relu_bwd")


def rowln_fwd(x, gamma, beta, y, mean, rstd, M, H, eps):
    _lib.check(_lib.load().modcr_rowln_fwd(_p(x), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd), M, H, eps, _stream()),
               "rowln_fwd")


def rowln_bwd(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, M, H):
    _lib.check(_lib.load().modcr_rowln_bwd(_p(dy), _p(x), _p(mean), _p(rstd), _p(gamma), _p(dx), _p(dgamma), _p(dbeta),
                                           M, H, _stream()), "rowln_bwd")


def pos_type_embed_fwd(pos, tt, Wp, bp, gam, bet, type_emb, out, mean, rstd, rows, H, eps):
    _chk(pos, F32, "img_pos_feat")
    _chk(tt, I64, "img_token_type_ids")
    _lib.check(_lib.load().modcr_pos_type_embed_fwd(_p(pos), _p(tt), _p(Wp), _p(bp), _p(gam), _p(bet), _p(type_emb), _p(out),
                                                    _p(mean), _p(rstd), rows, H, eps, _stream()), "pos_type_embed_fwd")


def pos_type_embed_bwd(d, pos, tt, Wp, bp, gam, mean, rstd, dWp, dbp, dgam, dbet, dtype, rows, H):
    _lib.check(_lib.load().modcr_pos_type_embed_bwd(_p(d), _p(pos), _p(tt), _p(Wp), _p(bp), _p(gam), _p(mean), _p(rstd),
                                                    _p(dWp), _p(dbp), _p(dgam), _p(dbet), _p(dtype), rows, H, _stream()),
               "pos_type_embed_bwd")


def rows_place(src32, y, y32, ymap, M, H):
    """fp32 rows [M,H] -> bf16 y (+ fp32 y32) at the physical rows of ``ymap``."""
    _chk(src32, F32, "rows_place src")
    _lib.check(_lib.load().modcr_rows_place(_p(src32), _p(y), _p(y32), ymap, M, H, _stream()), "rows_place")


def rows_gather(dy, dymap, dst, M, H):
    """bf16 gradient rows at ``dymap`` -> contiguous bf16 [M,H]."""
    _lib.check(_lib.load().modcr_rows_gather(_p(dy), dymap, _p(dst), M, H, _stream()), "rows_gather")


def key_bias(mask, kbias):
    _chk(mask, F32, "attention mask")
    _lib.check(_lib.load().modcr_key_bias(_p(mask), _p(kbias), mask.numel(), _stream()), "key_bias")


def chunk_bits(cmask, bits, N, T, cwords):
    _chk(cmask, F32, "chunk_attention_mask")
    _lib.check(_lib.load().modcr_chunk_bits(_p(cmask), _p(bits), N, T, cwords, _stream()), "chunk_bits")


def accum_f32_2d(dst, ld_dst, src, ld_src, rows, cols, accumulate=False):
    _lib.check(_lib.load().modcr_accum_f32_2d(_p(dst), ld_dst, _p(src), ld_src, rows, cols, int(accumulate),
                                              _stream()), "accum_f32_2d")
