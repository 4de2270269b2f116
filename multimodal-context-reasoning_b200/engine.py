"""Host-side execution engine of the hot path: explicit forward / backward programs over the
CUDA kernels (``ops``), with the activations each backward needs kept in caller-owned buffers.

The reference runs this path as eager PyTorch autograd over cuBLAS / ATen kernels
(modeling/modeling_bert.py:25-155, modeling/modeling_transfomres.py:471-727,
modeling/modeling_vcr_chunkalign_v10.py:45-373, 801-837, 1016-1165).  Here every arithmetic step is
one of our sm_100a kernels; PyTorch only owns memory, streams and the autograd *edges* between
the few coarse ``torch.autograd.Function`` nodes defined in ``functions.py``.

Layout in HBM (all row-major, bf16 unless noted):
  hidden states      [N, S, H]      S = T text rows followed by R region rows (no torch.cat)
  fused QKV          [N, S, 3H]     Q | K | V column blocks, heads are 64-column slices
  prefix K/V         [N, P, 2H]     K | V of the per-layer ``history_state`` rows
  FFN activations    [N*S, I]       pre-GELU (kept for GELU') and post-GELU
  LN statistics      fp32 [N*S]     mean / rstd per row
  parameter grads    fp32, produced directly by the tcgen05 wgrad GEMM (TN, fp32 out)
"""
import torch

from . import ops

BF16, F32 = torch.bfloat16, torch.float32
NUM_SMS = 148


def _empty(shape, dtype, dev):
    return torch.empty(shape, dtype=dtype, device=dev)


def _zeros(shape, dtype, dev):
    return torch.zeros(shape, dtype=dtype, device=dev)


# -------------------------------------------------------------------------------------------- dropout state
class Drop:
    """Training-mode dropout of one forward/backward pair: probabilities, the step's seed (a device int32 tensor the
    kernels read, so a replayed CUDA graph draws fresh masks) and the site-id base of the module.  Masks are a pure
    function of (seed, site, element index) and are regenerated in the backward pass."""
    __slots__ = ("p_hidden", "p_attn", "seed", "base")

    def __init__(self, p_hidden, p_attn, seed, base):
        self.p_hidden, self.p_attn, self.seed, self.base = float(p_hidden), float(p_attn), seed, int(base)

    def site(self, layer, k):
        return (self.base + layer * 8 + k) & 0x7FFFFFFF


_RNG_STATE = {}


def manual_seed(seed):
    """Seed the dropout masks of every device (the state advances by an LCG step per encoder / head forward).
    Existing per-device state tensors are reseeded IN PLACE: a captured CUDA graph has their addresses baked in."""
    seed = int(seed) & 0x7FFFFFFF
    _RNG_STATE["seed"] = seed
    for key, st in _RNG_STATE.items():
        if key != "seed":
            st.fill_(seed)


def next_seed(dev):
    """Advance the per-device dropout state (on the device: capturable) and return a private copy for this pass."""
    key = str(dev)
    st = _RNG_STATE.get(key)
    if st is None:
        st = torch.full((1,), _RNG_STATE.get("seed", 0x2545F491), dtype=torch.int32, device=dev)
        _RNG_STATE[key] = st
    st.mul_(1664525).add_(1013904223)      # int32 wrap-around LCG
    return st.clone()


def make_drop(module_training, config_hidden_p, config_attn_p, dev, base):
    """``None`` (no dropout arithmetic at all) unless the module is training with a non-zero probability."""
    ph, pa = float(config_hidden_p or 0.0), float(config_attn_p or 0.0)
    if not module_training or (ph <= 0.0 and pa <= 0.0):
        return None
    return Drop(ph, pa, next_seed(dev), base)


def _zeros_many(shapes, dev):
    """fp32 tensors that must start at zero (atomically reduced bias / LayerNorm gradients, split-K wgrad outputs),
    cleared by ONE multi-tensor fill instead of one fill kernel each.  They stay separate allocations so that
    autograd can adopt them as ``.grad`` without a copy."""
    ts = [torch.empty(sh, dtype=F32, device=dev) for sh in shapes]
    torch._foreach_zero_(ts)
    return ts


# --------------------------------------------------------------------------------------------
class ParamStore:
    """fp32 master parameters (the module's ``nn.Parameter`` objects, reference state_dict names)
    plus lazily refreshed bf16 operand copies for the tensor-core kernels."""

    def __init__(self, named):
        self.named = dict(named)
        self._cache = {}

    def f32(self, name):
        return self.named[name]

    def has(self, name):
        return name in self.named

    def _key(self, names):
        return tuple((self.named[n].data_ptr(), self.named[n]._version) for n in names)

    def bf16(self, name):
        ent = self._cache.get(name)
        key = self._key((name,))
        if ent is None or ent[0] != key:
            ent = (key, ops.to_bf16(self.named[name].detach()))
            self._cache[name] = ent
        return ent[1]

    def qkv(self, pfx):
        """Concatenated [3H,H] bf16 weight and [3H] fp32 bias of attention.self.{query,key,value}."""
        names = [pfx + n + s for n in ("query", "key", "value") for s in (".weight", ".bias")]
        ent = self._cache.get(pfx + "#qkv")
        key = self._key(names)
        if ent is None or ent[0] != key:
            ws = [self.named[pfx + n + ".weight"].detach() for n in ("query", "key", "value")]
            H = ws[0].shape[1]
            W = _empty((3 * H, H), BF16, ws[0].device)
            for i, w in enumerate(ws):
                ops.cast_pad(w.contiguous(), H, W[i * H:], H, H, H)
            b = torch.cat([self.named[pfx + n + ".bias"].detach() for n in ("query", "key", "value")]).contiguous()
            ent = (key, W, b)
            self._cache[pfx + "#qkv"] = ent
        return ent[1], ent[2]

    def padded(self, name, Dpad):
        ent = self._cache.get(name + "#pad")
        key = self._key((name,))
        if ent is None or ent[0] != key:
            w = self.named[name].detach().contiguous()
            out = _empty((w.shape[0], Dpad), BF16, w.device)
            ops.cast_pad(w, w.shape[1], out, Dpad, w.shape[0], w.shape[1])
            ent = (key, out)
            self._cache[name + "#pad"] = ent
        return ent[1]


def _split_k(M, N, K):
    """Split-K factor of the wgrad GEMM: the tile grid (256 x 256 tiles on CTA pairs) of a weight gradient is much
    smaller than the machine, so K (= rows of the activations) is split until the pairs are busy.  Cost model: waves of
    work units over the pairs x (k-blocks per unit + a fixed epilogue / reduce-add cost in k-block units)."""
    pairs = NUM_SMS // 2
    tiles = ((M + 255) // 256) * ((N + 255) // 256)
    kb = (K + 63) // 64
    best, best_cost = 1, None
    for s in range(1, 17):
        if s > 1 and kb // s < 4:
            break
        waves = (tiles * s + pairs - 1) // pairs
        cost = waves * ((kb + s - 1) // s + 8)
        if best_cost is None or cost < best_cost:
            best, best_cost = s, cost
    return best


def wgrad(dY, X, out_f, in_f, rows, ldy=None, ldx=None, out=None, zeroed=None, bias_out=None):
    """dW[out_f, in_f] = dY^T X  (fp32; split-K through TMA reduce-add when the tile grid is small).
    With ``out`` given the product is accumulated into it; ``zeroed`` is a pre-cleared destination.
    ``bias_out`` (fp32 [out_f], pre-zeroed or accumulating) receives += dY^T 1, the bias gradient, summed inside
    the same kernel from the dY tiles it stages (no separate pass over dY)."""
    dev = dY.device
    if zeroed is not None:
        out = zeroed
    if out is not None:
        ops.gemm(dY, X, out, out_f, in_f, rows, lda=ldy, ldb=ldx, a_mn=1, b_mn=1, accumulate=1,
                 split_k=_split_k(out_f, in_f, rows), a_colsum=bias_out)
        return out
    sk = _split_k(out_f, in_f, rows)
    if sk > 1:
        dW = _zeros((out_f, in_f), F32, dev)
        ops.gemm(dY, X, dW, out_f, in_f, rows, lda=ldy, ldb=ldx, a_mn=1, b_mn=1, accumulate=1, split_k=sk,
                 a_colsum=bias_out)
    else:
        dW = _empty((out_f, in_f), F32, dev)
        ops.gemm(dY, X, dW, out_f, in_f, rows, lda=ldy, ldb=ldx, a_mn=1, b_mn=1, a_colsum=bias_out)
    return dW


# -------------------------------------------------------------------------------------------- fp32 residual stream
class LnRes:
    """An fp32 residual that is the OUTPUT OF A LAYERNORM, kept as what its backward pass stores anyway -- the pre-LayerNorm
    tensor, its row statistics and the affine parameters -- instead of a second, fp32 copy of the output.  The dense-output
    GEMM that adds it (BertSelfOutput / BertOutput: ``LayerNorm(dense(x) + input_tensor)``,
    a_transformers/modeling_bert.py:369-373, 447-451) rebuilds ``(pre - mean) * rstd * gamma + beta`` in its epilogue
    (``modcr_gemm`` in_op 3): 70 MB less written per LayerNorm at cfg-2, 48 LayerNorms per step."""
    __slots__ = ("pre", "mean", "rstd", "gamma", "beta", "eps")

    def __init__(self, pre, mean, rstd, gamma, beta, eps):
        self.pre, self.mean, self.rstd, self.gamma, self.beta, self.eps = pre, mean, rstd, gamma, beta, eps


def res_kw(res):
    """Keyword arguments of ``ops.gemm`` that add the fp32 residual ``res`` (a dense fp32 tensor or an ``LnRes``)."""
    if isinstance(res, LnRes):
        return dict(epi_in=res.pre, in_ln=(res.mean, res.rstd, res.gamma.detach(), res.beta.detach()))
    return dict(epi_in=res, in_op=1)


def res_dense(res):
    """The residual as a dense fp32 tensor (only the optional-residual paths of the chunk-aware encoder need it)."""
    if not isinstance(res, LnRes):
        return res
    M, H = res.pre.shape
    dev = res.pre.device
    y, y32 = _empty((M, H), BF16, dev), _empty((M, H), F32, dev)
    ops.layernorm_fwd(res.pre, res.gamma.detach(), res.beta.detach(), y, _empty((M,), F32, dev), _empty((M,), F32, dev), M, H,
                      res.eps, y32=y32)
    return y32


# -------------------------------------------------------------------------------------------- FFN block
def ffn_fwd(ps, pfx, a, a32, M, H, I, eps, save, drop=None, site=0):
    """BertIntermediate + BertOutput (a_transformers/modeling_bert.py:434-437, 447-451).
    ``a`` is the bf16 tensor-core operand, ``a32`` the fp32 residual (dense tensor or ``LnRes``); returns ``(y, LnRes)``."""
    dev = a.device
    W1, W2 = ps.bf16(pfx + "intermediate.dense.weight"), ps.bf16(pfx + "output.dense.weight")
    gact = _empty((M, I), BF16, dev)
    u = _empty((M, I), BF16, dev) if save is not None else None
    ops.gemm(a, W1, gact, M, I, H, bias=ps.f32(pfx + "intermediate.dense.bias"), act=1, aux=u)
    pre2 = _empty((M, H), F32, dev)                                         # fp32 residual stream
    if drop is not None and drop.p_hidden > 0.0:
        # BertOutput: LayerNorm(dropout(dense(y)) + x)  (a_transformers/modeling_bert.py:447-451)
        ops.gemm(gact, W2, pre2, M, H, I, bias=ps.f32(pfx + "output.dense.bias"),
                 drop=(drop.seed, site, drop.p_hidden), **res_kw(a32))
    else:
        ops.gemm(gact, W2, pre2, M, H, I, bias=ps.f32(pfx + "output.dense.bias"), **res_kw(a32))
    y = _empty((M, H), BF16, dev)
    mean2, rstd2 = _empty((M,), F32, dev), _empty((M,), F32, dev)
    g2_, b2_ = ps.f32(pfx + "output.LayerNorm.weight"), ps.f32(pfx + "output.LayerNorm.bias")
    ops.layernorm_fwd(pre2, g2_, b2_, y, mean2, rstd2, M, H, eps)
    if save is not None:
        save.update(a=a, u=u, gact=gact, pre2=pre2, mean2=mean2, rstd2=rstd2)
    return y, LnRes(pre2, mean2, rstd2, g2_, b2_, eps)


def ffn_bwd(ps, pfx, sv, dy, M, H, I, grads, z=None, drop=None, site=0):
    """Returns d(a) given d(y).  ``z``: pre-zeroed (dgam, dbet, db2, db1, dW2, dW1) from ``_zeros_many``."""
    dev = dy.device
    W1, W2 = ps.bf16(pfx + "intermediate.dense.weight"), ps.bf16(pfx + "output.dense.weight")
    g2 = _empty((M, H), BF16, dev)
    if z is None:
        z = _zeros_many([(H,), (H,), (H,), (I,), (H, I), (I, H)], dev)
    dgam, dbet, db2, db1, dW2, dW1 = z
    hdrop = drop is not None and drop.p_hidden > 0.0
    g2d = _empty((M, H), BF16, dev) if hdrop else g2      # gradient w.r.t. the dense output (masked under dropout)
    ops.layernorm_bwd(dy, sv["pre2"], sv["mean2"], sv["rstd2"], ps.f32(pfx + "output.LayerNorm.weight"), g2, dgam, dbet,
                      db2, M, H, dx_masked=g2d if hdrop else None,
                      drop=(drop.seed, site, drop.p_hidden) if hdrop else None)
    grads[pfx + "output.LayerNorm.weight"], grads[pfx + "output.LayerNorm.bias"] = dgam, dbet
    grads[pfx + "output.dense.bias"] = db2
    grads[pfx + "output.dense.weight"] = wgrad(g2d, sv["gact"], H, I, M, zeroed=dW2)
    du = _empty((M, I), BF16, dev)
    ops.gemm(g2d, W2, du, M, I, H, b_mn=1, epi_in=sv["u"], in_op=2)          # (g2 W2) * GELU'(u)
    grads[pfx + "intermediate.dense.bias"] = db1
    grads[pfx + "intermediate.dense.weight"] = wgrad(du, sv["a"], I, H, M, zeroed=dW1, bias_out=db1)
    da = _empty((M, H), BF16, dev)
    ops.gemm(du, W1, da, M, H, I, b_mn=1, epi_in=g2, in_op=1)                # + residual branch
    return da


# -------------------------------------------------------------------------------------------- attention block
class AttnCfg:
    """Static description of one self-attention call."""
    __slots__ = ("N", "S", "H", "heads", "P", "T", "mode", "kbias", "cbits", "cwords", "chunk_id", "max_chunks",
                 "row_sel")

    def __init__(self, N, S, H, heads, P=0, T=0, mode=0, kbias=None, cbits=None, cwords=0, chunk_id=None,
                 max_chunks=0, row_sel=None):
        self.N, self.S, self.H, self.heads, self.P, self.T, self.mode = N, S, H, heads, P, T, mode
        self.kbias, self.cbits, self.cwords, self.chunk_id, self.max_chunks = kbias, cbits, cwords, chunk_id, max_chunks
        self.row_sel = row_sel      # int64 [N,T] (align_pos): the only text rows whose text->region probabilities travel


def attention_fwd(ps, pfx, x, x32, hist, ac, eps, save, do_chunk_cross=False, probs_img=None, drop=None, site=0):
    """CaptionBertAttention: self-attention with prefix (modeling/modeling_bert.py:34-93, v10:55-129)
    followed by BertSelfOutput.  ``x`` [N*S,H]; ``hist`` [N*P,H] or None."""
    dev = x.device
    N, S, H, heads, P = ac.N, ac.S, ac.H, ac.heads, ac.P
    M = N * S
    Wqkv, bqkv = ps.qkv(pfx + "self.")
    qkv = _empty((M, 3 * H), BF16, dev)
    ops.gemm(x, Wqkv, qkv, M, 3 * H, H, bias=bqkv)
    pkv = None
    if P > 0:
        pkv = _empty((N * P, 2 * H), BF16, dev)
        ops.gemm(hist, Wqkv[H:], pkv, N * P, 2 * H, H, bias=bqkv[H:])
    if do_chunk_cross:
        ops.chunk_mean_q(qkv, 3 * H, N, S, H, ac.chunk_id, ac.T, ac.max_chunks)
    ctx = _empty((M, H), BF16, dev)
    lse = _empty((N, heads, S), F32, dev)
    adrop = (drop.seed, site, drop.p_attn) if (drop is not None and drop.p_attn > 0.0) else None
    ops.prefix_attn_fwd(N, S, P, heads, ac.T, ac.mode, qkv, pkv, ac.kbias, ac.cbits, ac.cwords, ctx, lse, probs_img,
                        drop=adrop, row_sel=ac.row_sel if probs_img is not None else None)
    pre1 = _empty((M, H), F32, dev)                                         # fp32 residual stream
    if drop is not None and drop.p_hidden > 0.0:
        # BertSelfOutput: LayerNorm(dropout(dense(ctx)) + x)  (a_transformers/modeling_bert.py:369-373)
        ops.gemm(ctx, ps.bf16(pfx + "output.dense.weight"), pre1, M, H, H, bias=ps.f32(pfx + "output.dense.bias"),
                 drop=(drop.seed, site + 1, drop.p_hidden), **res_kw(x32))
    else:
        ops.gemm(ctx, ps.bf16(pfx + "output.dense.weight"), pre1, M, H, H, bias=ps.f32(pfx + "output.dense.bias"),
                 **res_kw(x32))
    a = _empty((M, H), BF16, dev)
    mean1, rstd1 = _empty((M,), F32, dev), _empty((M,), F32, dev)
    g1_, b1_ = ps.f32(pfx + "output.LayerNorm.weight"), ps.f32(pfx + "output.LayerNorm.bias")
    ops.layernorm_fwd(pre1, g1_, b1_, a, mean1, rstd1, M, H, eps)
    if save is not None:
        save.update(x=x, hist=hist, qkv=qkv, pkv=pkv, ctx=ctx, lse=lse, pre1=pre1, mean1=mean1, rstd1=rstd1,
                    chunk=do_chunk_cross, probs_img=probs_img)
    return a, LnRes(pre1, mean1, rstd1, g1_, b1_, eps)


def attention_bwd(ps, pfx, sv, da, ac, grads, dalign=None, ws=None, z=None, drop=None, site=0, before_core=None):
    """Returns (dx, dhist).  ``z``: pre-zeroed (dgam, dbet, dbo, dbqkv, dWo, dWqkv) from ``_zeros_many``.
    ``before_core()`` is called right before the attention-core kernel is launched (the data-parallel reducer issues its
    collectives there, ``parallel.GradReducer.open_window``)."""
    dev = da.device
    N, S, H, heads, P = ac.N, ac.S, ac.H, ac.heads, ac.P
    M = N * S
    Wqkv, _ = ps.qkv(pfx + "self.")
    g1 = _empty((M, H), BF16, dev)
    if z is None:
        z = _zeros_many([(H,), (H,), (H,), (3 * H,), (H, H), (3 * H, H)], dev)
    dgam, dbet, dbo, dbqkv, dWo, dWqkv = z
    hdrop = drop is not None and drop.p_hidden > 0.0
    g1d = _empty((M, H), BF16, dev) if hdrop else g1      # gradient w.r.t. the dense output (masked under dropout)
    ops.layernorm_bwd(da, sv["pre1"], sv["mean1"], sv["rstd1"], ps.f32(pfx + "output.LayerNorm.weight"), g1, dgam, dbet,
                      dbo, M, H, dx_masked=g1d if hdrop else None,
                      drop=(drop.seed, site + 1, drop.p_hidden) if hdrop else None)
    grads[pfx + "output.LayerNorm.weight"], grads[pfx + "output.LayerNorm.bias"] = dgam, dbet
    grads[pfx + "output.dense.bias"] = dbo
    grads[pfx + "output.dense.weight"] = wgrad(g1d, sv["ctx"], H, H, M, zeroed=dWo)
    dctx = _empty((M, H), BF16, dev)
    ops.gemm(g1d, ps.bf16(pfx + "output.dense.weight"), dctx, M, H, H, b_mn=1)
    dqkv = _empty((M, 3 * H), BF16, dev)
    dpkv = _empty((N * P, 2 * H), BF16, dev) if P > 0 else None
    if ws is None:
        ws = _empty((ops.attn_bwd_workspace_bytes(N, S, heads) // 4,), F32, dev)
    use_align = dalign if sv["probs_img"] is not None else None
    if before_core is not None:
        before_core()
    ops.prefix_attn_bwd(N, S, P, heads, ac.T, ac.mode, sv["qkv"], sv["pkv"], ac.kbias, ac.cbits, ac.cwords, sv["ctx"],
                        sv["lse"], sv["probs_img"] if use_align is not None else None, dctx, use_align, dqkv, dpkv, ws,
                        drop=(drop.seed, site, drop.p_attn) if (drop is not None and drop.p_attn > 0.0) else None,
                        row_sel=ac.row_sel if use_align is not None else None)
    if sv["chunk"]:
        ops.chunk_mean_q(dqkv, 3 * H, N, S, H, ac.chunk_id, ac.T, ac.max_chunks)   # self-adjoint operator
    wgrad(dqkv, sv["x"], 3 * H, H, M, zeroed=dWqkv, bias_out=dbqkv)
    dhist = None
    if P > 0:
        wgrad(dpkv, sv["hist"], 2 * H, H, N * P, out=dWqkv[H:], bias_out=dbqkv[H:])
        dhist = _empty((N * P, H), BF16, dev)
        ops.gemm(dpkv, Wqkv[H:], dhist, N * P, H, 2 * H, b_mn=1)
    for i, n in enumerate(("query", "key", "value")):
        grads[pfx + "self." + n + ".weight"] = dWqkv[i * H:(i + 1) * H]
        grads[pfx + "self." + n + ".bias"] = dbqkv[i * H:(i + 1) * H]
    dx = _empty((M, H), BF16, dev)
    ops.gemm(dqkv, Wqkv, dx, M, H, 3 * H, b_mn=1, epi_in=g1, in_op=1)
    return dx, dhist


def layer_fwd(ps, pfx, x, x32, hist, ac, I, eps, save, do_chunk_cross=False, probs_img=None, drop=None, layer=0):
    """CaptionBertLayer.forward (modeling/modeling_transfomres.py:481-489, v10:140-150); returns (y, y32).
    Dropout sites of layer ``layer``: +0 attention probabilities, +1 attention output, +2 FFN output."""
    site = drop.site(layer + 1, 0) if drop is not None else 0
    a, a32 = attention_fwd(ps, pfx + "attention.", x, x32, hist, ac, eps, save, do_chunk_cross, probs_img, drop, site)
    return ffn_fwd(ps, pfx, a, a32, ac.N * ac.S, ac.H, I, eps, save, drop, site + 2)


def layer_bwd(ps, pfx, sv, dy, ac, I, grads, dalign=None, ws=None, drop=None, layer=0, before_core=None):
    H = ac.H
    z = _zeros_many([(H,), (H,), (H,), (I,), (H, I), (I, H), (H,), (H,), (H,), (3 * H,), (H, H), (3 * H, H)], dy.device)
    site = drop.site(layer + 1, 0) if drop is not None else 0
    da = ffn_bwd(ps, pfx, sv, dy, ac.N * ac.S, H, I, grads, z[:6], drop, site + 2)
    return attention_bwd(ps, pfx + "attention.", sv, da, ac, grads, dalign, ws, z[6:], drop, site, before_core)


# -------------------------------------------------------------------------------------------- embeddings
def embeddings_fwd(ps, cfg, input_ids, token_type_ids, position_ids, img_feats, save, pre_rows=None, img_extra=None):
    """BertEmbeddings + img_embedding (+LayerNorm) written into one [N,S,H] buffer
    (a_transformers/modeling_bert.py:184-211; modeling_transfomres.py:676-684; v10:339-345).
    ``pre_rows`` (fp32 [N,Pn,H]): rows that arrive already embedded and are placed IN FRONT of the text rows -- the
    prompt embeddings of the language-model side (modeling/modeling_ensemble.py:498-503).
    ``img_extra`` = (img_pos_feat fp32 [N,R,7], img_token_type_ids int64 [N,R]): the region embedding of
    ``BertImgModel_gtToken`` (modeling/modeling_bert.py:390-397): type embedding + LayerNorm(Linear(7 -> H)(box)) + feature
    projection, summed in the projection GEMM's epilogue, then the image LayerNorm."""
    dev = input_ids.device
    N, T = input_ids.shape
    H = cfg["H"]
    R = 0 if img_feats is None else img_feats.shape[1]
    Pn = 0 if pre_rows is None else pre_rows.shape[1]
    S = Pn + T + R
    x0, x0_32 = _empty((N, S, H), BF16, dev), _empty((N, S, H), F32, dev)
    tmean, trstd = _empty((N * T,), F32, dev), _empty((N * T,), F32, dev)
    tmap = ops.rowmap(T, S, Pn)
    if Pn:
        ops.rows_place(pre_rows.to(F32).contiguous().view(N * Pn, H), x0, x0_32, ops.rowmap(Pn, S, 0), N * Pn, H)
    ids = input_ids.contiguous()
    tts = None if token_type_ids is None else token_type_ids.contiguous()
    pos = None if position_ids is None else position_ids.expand(N, T).contiguous()
    ops.txt_embed_ln_fwd(ids, tts, pos, ps.f32("embeddings.word_embeddings.weight"),
                         ps.f32("embeddings.token_type_embeddings.weight"),
                         ps.f32("embeddings.position_embeddings.weight"), ps.f32("embeddings.LayerNorm.weight"),
                         ps.f32("embeddings.LayerNorm.bias"), x0, tmap, tmean, trstd, N, T, H, cfg["eps"], y32=x0_32)
    sv = dict(ids=ids, tts=tts, pos=pos, tmean=tmean, trstd=trstd, N=N, T=T, R=R, S=S, Pn=Pn)
    if R > 0:
        D = img_feats.shape[2]
        Dpad = (D + 63) // 64 * 64
        rows = N * R
        feats = img_feats.contiguous()
        packed = _empty((rows, Dpad), BF16, dev)
        pre = _empty((rows, H), F32, dev)
        imean, irstd = _empty((rows,), F32, dev), _empty((rows,), F32, dev)
        use_ln = cfg["use_img_layernorm"]
        extra32 = None
        if img_extra is not None:
            posf = img_extra[0].to(F32).contiguous().view(rows, 7)
            tti = None if img_extra[1] is None else img_extra[1].to(torch.int64).contiguous().view(rows)
            extra32 = _empty((rows, H), F32, dev)
            pmean, prstd = _empty((rows,), F32, dev), _empty((rows,), F32, dev)
            ops.pos_type_embed_fwd(posf, tti, ps.f32("img_pos_linear.weight"), ps.f32("img_pos_linear.bias"),
                                   ps.f32("img_pos_layer_norm.weight"), ps.f32("img_pos_layer_norm.bias"),
                                   ps.f32("img_token_type_embeddings.weight"), extra32, pmean, prstd, rows, H, 1e-12)
            sv.update(posf=posf, tti=tti, pmean=pmean, prstd=prstd)
        ops.img_embed_ln_fwd(feats, rows, D, Dpad, ps.padded("img_embedding.weight", Dpad),
                             ps.f32("img_embedding.bias"),
                             ps.f32("LayerNorm.weight") if use_ln else None,
                             ps.f32("LayerNorm.bias") if use_ln else None, use_ln, cfg["img_eps"], packed, pre, x0,
                             ops.rowmap(R, S, Pn + T), imean, irstd, H, y32=x0_32, extra32=extra32)
        sv.update(packed=packed, pre=pre, imean=imean, irstd=irstd, D=D, Dpad=Dpad)
    if save is not None:
        save.update(sv)
    return x0, x0_32


def embeddings_bwd(ps, cfg, sv, dx0, grads):
    """Returns the bf16 gradient [N,Pn,H] of the pre-embedded rows (None without them)."""
    dev = dx0.device
    N, T, R, S, H = sv["N"], sv["T"], sv["R"], sv["S"], cfg["H"]
    Pn = sv.get("Pn", 0)
    word = ps.f32("embeddings.word_embeddings.weight")
    type_ = ps.f32("embeddings.token_type_embeddings.weight")
    pos = ps.f32("embeddings.position_embeddings.weight")
    dw, dt, dp = torch.zeros_like(word), torch.zeros_like(type_), torch.zeros_like(pos)
    dg, db = _zeros((H,), F32, dev), _zeros((H,), F32, dev)
    ops.txt_embed_ln_bwd(dx0, ops.rowmap(T, S, Pn), sv["ids"], sv["tts"], sv["pos"], word, type_, pos,
                         ps.f32("embeddings.LayerNorm.weight"), sv["tmean"], sv["trstd"], dw, dt, dp, dg, db, N, T, H,
                         cfg["pad_token_id"])
    grads["embeddings.word_embeddings.weight"] = dw
    grads["embeddings.token_type_embeddings.weight"] = dt
    grads["embeddings.position_embeddings.weight"] = dp
    grads["embeddings.LayerNorm.weight"], grads["embeddings.LayerNorm.bias"] = dg, db
    if R > 0:
        rows, D, Dpad = N * R, sv["D"], sv["Dpad"]
        use_ln = cfg["use_img_layernorm"]
        dpre = _empty((rows, H), BF16, dev)
        dWp = _empty((H, Dpad), F32, dev)
        dbias = _zeros((H,), F32, dev)
        dgi, dbi = (_zeros((H,), F32, dev), _zeros((H,), F32, dev)) if use_ln else (None, None)
        ops.img_embed_ln_bwd(dx0, ops.rowmap(R, S, Pn + T), sv["packed"], sv["pre"], sv["imean"], sv["irstd"],
                             ps.f32("LayerNorm.weight") if use_ln else None, use_ln, dpre, dWp, dbias, dgi, dbi, rows,
                             Dpad, H)
        dW = _empty((H, D), F32, dev)
        ops.accum_f32_2d(dW, D, dWp, Dpad, H, D)
        grads["img_embedding.weight"], grads["img_embedding.bias"] = dW, dbias
        if use_ln:
            grads["LayerNorm.weight"], grads["LayerNorm.bias"] = dgi, dbi
        if "posf" in sv:    # type + box-position term of BertImgModel_gtToken: d(sum) = dpre
            z = _zeros_many([(H, 7), (H,), (H,), (H,), (2, H)], dev)
            ops.pos_type_embed_bwd(dpre, sv["posf"], sv["tti"], ps.f32("img_pos_linear.weight"), ps.f32("img_pos_linear.bias"),
                                   ps.f32("img_pos_layer_norm.weight"), sv["pmean"], sv["prstd"], z[0], z[1], z[2], z[3], z[4],
                                   rows, H)
            grads["img_pos_linear.weight"], grads["img_pos_linear.bias"] = z[0], z[1]
            grads["img_pos_layer_norm.weight"], grads["img_pos_layer_norm.bias"] = z[2], z[3]
            grads["img_token_type_embeddings.weight"] = z[4]
    if Pn:
        dpre_rows = _empty((N, Pn, H), BF16, dev)
        ops.rows_gather(dx0, ops.rowmap(Pn, S, 0), dpre_rows, N * Pn, H)
        return dpre_rows
    return None


# -------------------------------------------------------------------------------------------- pooler
def pooler_fwd(ps, h, N, S, H, out=None, ldc=None):
    """BertPooler (a_transformers/modeling_bert.py:640-646): tanh(W h[:,0] + b); reads row 0 of every
    sequence in place (lda = S*H) and can write into a column block of a wider buffer."""
    if out is None:
        out = _empty((N, H), BF16, h.device)
    ops.gemm(h, ps.bf16("pooler.dense.weight"), out, N, H, H, lda=S * H, ldc=ldc, bias=ps.f32("pooler.dense.bias"),
             act=2)
    return out


def pooler_bwd(ps, h, pooled, dpooled, N, S, H, dh, grads):
    """Accumulates the gradient of row 0 into ``dh`` [N,S,H]."""
    dev = h.device
    dpre = _empty((N, H), BF16, dev)
    ops.tanh_bwd(dpooled, pooled, dpre)
    db = _zeros((H,), F32, dev)
    ops.colsum(dpre, H, N, H, db)
    grads["pooler.dense.bias"] = db
    grads["pooler.dense.weight"] = wgrad(dpre, h, H, H, N, ldx=S * H)
    d0 = _empty((N, H), BF16, dev)
    ops.gemm(dpre, ps.bf16("pooler.dense.weight"), d0, N, H, H, b_mn=1)
    ops.copy_rows(d0, 1, 0, dh, S, 0, N, 1, H, accumulate=True)


# -------------------------------------------------------------------------------------------- masks
def make_key_bias(mask2d):
    """(1 - mask) * -10000 (modeling_transfomres.py:628-641; no dtype cast)."""
    m = mask2d.to(F32).contiguous()
    kb = torch.empty_like(m)
    ops.key_bias(m, kb)
    return kb


def make_chunk_bits(chunk_mask):
    N, T, _ = chunk_mask.shape
    cwords = (T + 31) // 32
    bits = _empty((N, T, cwords), torch.int32, chunk_mask.device)
    ops.chunk_bits(chunk_mask.to(F32).contiguous(), bits, N, T, cwords)
    return bits, cwords


def pack_chunk_ids(gather_index, T, device):
    """Ragged ``gather_index`` (list of N int64 tensors, chunk id of tokens 1..L-2,
    Data/VCRChunkAlign.py:666-670) -> int32 [N,T] with -1 outside the chunked span, plus the
    largest chunk count.  One host-side pack and one H2D copy per batch."""
    N = len(gather_index)
    lens = [int(g.numel()) for g in gather_index]
    flat = torch.cat([g.reshape(-1) for g in gather_index]) if N else torch.zeros(0, dtype=torch.int64)
    flat = flat.to("cpu") if flat.is_cuda else flat
    out = torch.full((N, T), -1, dtype=torch.int32)
    off, mx = 0, 0
    for n, L in enumerate(lens):
        if L:
            seg = flat[off:off + L]
            out[n, 1:1 + L] = seg.to(torch.int32)
            mx = max(mx, int(seg.max()) + 1)
        off += L
    return out.to(device, non_blocking=True), max(mx, 1)
