"""GPU parity tests of every CUDA kernel, called through the C ABI (ctypes), against plain
PyTorch fp32 references of the same op evaluated on the same (bf16-rounded) inputs.

Tolerances: bf16 outputs are compared with a normalised max error (max|a-b| / max|b|) of 2e-2
(1.5e-2 typical bf16 rounding of a long fp32-accumulated contraction); fp32 statistics /
losses with 1e-4..1e-3; index / mask work bit-exactly."""
import math

import pytest
import torch

import modcr_b200  # noqa: F401
from modcr_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"
NEG = -10000.0


def nerr(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))


def rnd(*shape, scale=1.0, seed=0, dtype=torch.bfloat16):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(DEV)


# ------------------------------------------------------------------------------- GEMM via ctypes
def test_gemm_forward_variants():
    M, N, K = 712, 768, 2304
    A, B = rnd(M, K, seed=1), rnd(N, K, scale=0.05, seed=2)
    bias = rnd(N, seed=3, dtype=torch.float32)
    res = rnd(M, N, seed=4)
    ref = A.float() @ B.float().t() + bias
    C = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    ops.gemm(A, B, C, M, N, K, bias=bias)
    assert nerr(C, ref) < 1e-2
    ops.gemm(A, B, C, M, N, K, bias=bias, epi_in=res, in_op=1)
    assert nerr(C, ref + res.float()) < 1e-2
    aux = torch.empty_like(C)
    ops.gemm(A, B, C, M, N, K, bias=bias, act=1, aux=aux)
    assert nerr(aux, ref) < 1e-2 and nerr(C, torch.nn.functional.gelu(ref)) < 1e-2
    # strided A (pooler reads row 0 of every sequence: lda = S*H)
    S_, H = 5, 256
    X = rnd(64, S_, H, seed=5)
    Wp = rnd(H, H, scale=0.05, seed=6)
    out = torch.empty(64, H, dtype=torch.bfloat16, device=DEV)
    ops.gemm(X, Wp, out, 64, H, H, lda=S_ * H, act=2)
    assert nerr(out, torch.tanh(X[:, 0].float() @ Wp.float().t())) < 1e-2


def test_gemm_epilogue_input_slot_is_not_overwritten_early():
    """Regression test of a write-after-read race: the epilogue-input (residual) slot used to be handed back to the TMA
    loader right behind the shared-memory reads, without a proxy fence; about once per 1300 launches of this shape the
    next load overwrote the slot under a slow reader and the residual of the chunk 64 columns ahead was added.  The
    BertOutput.dense shape of cfg-2 (fp32 residual stream), repeated: every launch must be bit-identical."""
    M, N, K = 128 * 178, 768, 3072
    A, B = rnd(M, K, seed=1), rnd(N, K, scale=0.02, seed=2)
    bias = rnd(N, seed=3, dtype=torch.float32)
    res = rnd(M, N, seed=4, dtype=torch.float32)
    C = torch.empty(M, N, dtype=torch.float32, device=DEV)
    ops.gemm(A, B, C, M, N, K, bias=bias, epi_in=res, in_op=1)
    ref = C.clone()
    assert nerr(ref, A.float() @ B.float().t() + bias + res) < 1e-2
    bad = 0
    for _ in range(4000):
        ops.gemm(A, B, C, M, N, K, bias=bias, epi_in=res, in_op=1)
        bad += int(not torch.equal(C, ref))
    assert bad == 0, "%d of 4000 launches differed" % bad
    # bf16 epilogue input (dgrad + residual)
    res16 = rnd(M, N, seed=5)
    C16 = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
    ops.gemm(A, B, C16, M, N, K, epi_in=res16, in_op=1)
    ref16 = C16.clone()
    for _ in range(1500):
        ops.gemm(A, B, C16, M, N, K, epi_in=res16, in_op=1)
        bad += int(not torch.equal(C16, ref16))
    assert bad == 0, "%d bf16-input launches differed" % bad


def test_gemm_backward_variants():
    M, N, K = 712, 768, 1024   # dY [M,N], W [N,K], X [M,K]
    dY, W, X = rnd(M, N, seed=1), rnd(N, K, scale=0.05, seed=2), rnd(M, K, seed=3)
    dX = torch.empty(M, K, dtype=torch.bfloat16, device=DEV)
    ops.gemm(dY, W, dX, M, K, N, b_mn=1)                               # dgrad: dX = dY W
    assert nerr(dX, dY.float() @ W.float()) < 1e-2
    dW = torch.zeros(N, K, dtype=torch.float32, device=DEV)
    ops.gemm(dY, X, dW, N, K, M, a_mn=1, b_mn=1)                       # wgrad: dW = dY^T X
    assert nerr(dW, dY.float().t() @ X.float()) < 1e-3
    ops.gemm(dY, X, dW, N, K, M, a_mn=1, b_mn=1, accumulate=1)         # accumulate (TMA reduce-add)
    assert nerr(dW, 2 * (dY.float().t() @ X.float())) < 1e-3


def test_gemm_wgrad_fused_bias_gradient():
    """a_colsum: the bias gradient dY^T 1 summed inside the wgrad kernel from the staged dY tiles (all splits, ragged
    K and M), against a separate column sum."""
    for (rows, out_f, in_f, sk) in ((1424, 768, 768, 4), (3048, 2304, 768, 3), (400, 2056, 768, 1), (128 * 9 + 5, 3072, 768, 2)):
        dY, X = rnd(rows, out_f, seed=rows), rnd(rows, in_f, seed=rows + 1)
        dW = torch.zeros(out_f, in_f, device=DEV)
        db = torch.zeros(out_f, device=DEV)
        ops.gemm(dY, X, dW, out_f, in_f, rows, a_mn=1, b_mn=1, accumulate=1, split_k=sk, a_colsum=db)
        assert nerr(dW, dY.float().t() @ X.float()) < 1e-4
        assert nerr(db, dY.float().sum(0)) < 1e-5, (rows, out_f, sk)
        ops.gemm(dY, X, dW, out_f, in_f, rows, a_mn=1, b_mn=1, accumulate=1, split_k=sk, a_colsum=db)   # accumulates
        assert nerr(db, 2 * dY.float().sum(0)) < 1e-5


def test_gemm_fp32_residual_stream():
    M, N, K = 712, 768, 3072
    A, B = rnd(M, K, seed=1), rnd(N, K, scale=0.05, seed=2)
    bias = rnd(N, seed=3, dtype=torch.float32)
    res = rnd(M, N, seed=4, dtype=torch.float32)
    C = torch.empty(M, N, dtype=torch.float32, device=DEV)
    ops.gemm(A, B, C, M, N, K, bias=bias, epi_in=res, in_op=1)
    ref = A.float() @ B.float().t() + bias + res
    assert nerr(C, ref) < 1e-4


def test_gemm_fp32_residual_rebuilt_from_the_previous_layernorm():
    """modcr_gemm in_op 3: C = A B^T + bias + LayerNorm(in) with the LayerNorm output rebuilt in the epilogue from the
    pre-LayerNorm tensor, its row statistics and gamma / beta (BertSelfOutput / BertOutput add the previous LayerNorm's
    output, a_transformers/modeling_bert.py:369-373, 447-451) -- against the two-kernel form (LayerNorm with an fp32 copy,
    then the fp32 residual add) and against torch; a ragged M exercises the row guard."""
    for (M, N, K) in ((712, 768, 3072), (128 * 178, 768, 768), (130, 1024, 1024)):
        A, B = rnd(M, K, seed=1), rnd(N, K, scale=0.03, seed=2)
        bias = rnd(N, seed=3, dtype=torch.float32)
        pre = rnd(M, N, seed=4, dtype=torch.float32, scale=1.5)
        gam = 1.0 + 0.1 * rnd(N, seed=5, dtype=torch.float32)
        bet = 0.05 * rnd(N, seed=6, dtype=torch.float32)
        y = torch.empty(M, N, dtype=torch.bfloat16, device=DEV)
        y32 = torch.empty(M, N, dtype=torch.float32, device=DEV)
        mean, rstd = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
        ops.layernorm_fwd(pre, gam, bet, y, mean, rstd, M, N, 1e-12, y32=y32)
        C2 = torch.empty(M, N, dtype=torch.float32, device=DEV)
        ops.gemm(A, B, C2, M, N, K, bias=bias, epi_in=y32, in_op=1)
        C3 = torch.full((M, N), float("nan"), dtype=torch.float32, device=DEV)
        ops.gemm(A, B, C3, M, N, K, bias=bias, epi_in=pre, in_ln=(mean, rstd, gam, bet))
        ref = A.float() @ B.float().t() + bias + torch.nn.functional.layer_norm(pre, (N,), gam, bet, 1e-12)
        assert float((C3 - C2).abs().max()) < 2e-5, (M, N, K)          # same arithmetic, two routes
        assert nerr(C3, ref) < 2e-3, (M, N, K)


def test_layernorm_fp32_input_dual_output():
    M, H = 333, 768
    x = rnd(M, H, scale=2.0, seed=1, dtype=torch.float32)
    gamma, beta = (1 + 0.1 * torch.randn(H)).to(DEV), (0.1 * torch.randn(H)).to(DEV)
    y = torch.empty(M, H, dtype=torch.bfloat16, device=DEV)
    y32 = torch.empty(M, H, device=DEV)
    mean, rstd = torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    ops.layernorm_fwd(x, gamma, beta, y, mean, rstd, M, H, 1e-12, y32=y32)
    xr = x.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (H,), gamma, beta, 1e-12)
    torch.testing.assert_close(y32, yr.detach(), atol=2e-5, rtol=1e-5)
    assert torch.equal(y, y32.to(torch.bfloat16))
    dy = rnd(M, H, seed=2)
    yr.backward(dy.float())
    dx = torch.empty(M, H, dtype=torch.bfloat16, device=DEV)
    dg, db = torch.zeros(H, device=DEV), torch.zeros(H, device=DEV)
    ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx, dg, db, None, M, H)
    assert nerr(dx, xr.grad) < 1e-2


# ------------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("H", [768, 1024])
def test_layernorm_fwd_bwd(H):
    M = 1000
    x = rnd(M, H, scale=2.0, seed=1)
    gamma = (1 + 0.1 * torch.randn(H)).to(DEV)
    beta = (0.1 * torch.randn(H)).to(DEV)
    y = torch.empty_like(x)
    mean = torch.empty(M, device=DEV)
    rstd = torch.empty(M, device=DEV)
    ops.layernorm_fwd(x, gamma, beta, y, mean, rstd, M, H, 1e-12)
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(xr, (H,), gr, br, 1e-12)
    assert nerr(y, yr) < 1e-2
    torch.testing.assert_close(mean, xr.mean(-1), atol=1e-5, rtol=1e-5)
    dy = rnd(M, H, seed=2)
    yr.backward(dy.float())
    dx = torch.empty_like(x)
    dg, db, dbias = (torch.zeros(H, device=DEV) for _ in range(3))
    ops.layernorm_bwd(dy, x, mean, rstd, gamma, dx, dg, db, dbias, M, H)
    assert nerr(dx, xr.grad) < 1e-2
    assert nerr(dg, gr.grad) < 1e-3 and nerr(db, br.grad) < 1e-3
    assert nerr(dbias, xr.grad.sum(0)) < 2e-2   # sum of bf16-rounded... compared to fp32 sum


def test_layernorm_rowmap_writes_into_concat_buffer():
    N, T, R, H = 3, 5, 4, 256
    S = T + R
    x = rnd(N * R, H, seed=1)
    gamma, beta = torch.ones(H, device=DEV), torch.zeros(H, device=DEV)
    buf = torch.zeros(N, S, H, dtype=torch.bfloat16, device=DEV)
    mean, rstd = torch.empty(N * R, device=DEV), torch.empty(N * R, device=DEV)
    ops.layernorm_fwd(x, gamma, beta, buf, mean, rstd, N * R, H, 1e-12, ymap=ops.rowmap(R, S, T))
    ref = torch.nn.functional.layer_norm(x.float(), (H,)).view(N, R, H)
    assert nerr(buf[:, T:], ref) < 1e-2
    assert float(buf[:, :T].abs().max()) == 0.0


# ------------------------------------------------------------------------------- embeddings
def test_txt_embed_ln_fwd_bwd():
    N, T, R, H, V = 4, 9, 3, 768, 500
    S = T + R
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, V, (N, T), generator=g).to(DEV)
    ids[0, -2:] = 0   # padding rows (padding_idx = 0 receives no gradient)
    tts = torch.randint(0, 2, (N, T), generator=g).to(DEV)
    word = (0.02 * torch.randn(V, H, generator=g)).to(DEV)
    type_ = (0.02 * torch.randn(2, H, generator=g)).to(DEV)
    pos = (0.02 * torch.randn(512, H, generator=g)).to(DEV)
    gamma = (1 + 0.1 * torch.randn(H, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(H, generator=g)).to(DEV)
    buf = torch.zeros(N, S, H, dtype=torch.bfloat16, device=DEV)
    mean, rstd = torch.empty(N * T, device=DEV), torch.empty(N * T, device=DEV)
    ymap = ops.rowmap(T, S, 0)
    ops.txt_embed_ln_fwd(ids, tts, None, word, type_, pos, gamma, beta, buf, ymap, mean, rstd, N, T, H, 1e-12)
    wr, tr, pr = (t.clone().requires_grad_(True) for t in (word, type_, pos))
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    e = torch.nn.functional.embedding(ids, wr, padding_idx=0) + tr[tts] + pr[torch.arange(T, device=DEV)][None]
    yr = torch.nn.functional.layer_norm(e, (H,), gr, br, 1e-12)
    assert nerr(buf[:, :T], yr) < 1e-2
    dy = torch.zeros(N, S, H, dtype=torch.bfloat16, device=DEV)
    dy[:, :T] = rnd(N, T, H, seed=5)
    yr.backward(dy[:, :T].float())
    dw, dt, dp = torch.zeros_like(word), torch.zeros_like(type_), torch.zeros_like(pos)
    dg, db = torch.zeros(H, device=DEV), torch.zeros(H, device=DEV)
    ops.txt_embed_ln_bwd(dy, ymap, ids, tts, None, word, type_, pos, gamma, mean, rstd, dw, dt, dp, dg, db, N, T, H, 0)
    assert nerr(dw, wr.grad) < 1e-3 and nerr(dt, tr.grad) < 1e-3 and nerr(dp, pr.grad) < 1e-3
    assert float(dw[0].abs().max()) == 0.0
    assert nerr(dg, gr.grad) < 1e-3 and nerr(db, br.grad) < 1e-3


def test_img_embed_ln_fwd_bwd():
    N, T, R, H, D = 3, 4, 50, 768, 2054
    Dpad, S = 2112, 4 + 50
    feats = torch.randn(N, R, D, generator=torch.Generator().manual_seed(1)).to(DEV)
    W = (0.02 * torch.randn(H, D, generator=torch.Generator().manual_seed(2))).to(DEV)
    bias = (0.02 * torch.randn(H, generator=torch.Generator().manual_seed(3))).to(DEV)
    gamma, beta = (1 + 0.1 * torch.randn(H)).to(DEV), (0.1 * torch.randn(H)).to(DEV)
    Wb = torch.zeros(H, Dpad, dtype=torch.bfloat16, device=DEV)
    ops.cast_pad(W, D, Wb, Dpad, H, D)
    assert torch.equal(Wb[:, :D], W.to(torch.bfloat16)) and float(Wb[:, D:].abs().max()) == 0.0
    rows = N * R
    packed = torch.empty(rows, Dpad, dtype=torch.bfloat16, device=DEV)
    pre = torch.empty(rows, H, dtype=torch.float32, device=DEV)
    buf = torch.zeros(N, S, H, dtype=torch.bfloat16, device=DEV)
    buf32 = torch.zeros(N, S, H, dtype=torch.float32, device=DEV)
    mean, rstd = torch.empty(rows, device=DEV), torch.empty(rows, device=DEV)
    ymap = ops.rowmap(R, S, T)
    ops.img_embed_ln_fwd(feats, rows, D, Dpad, Wb, bias, gamma, beta, 1, 1e-12, packed, pre, buf, ymap, mean, rstd, H,
                         y32=buf32)
    assert torch.equal(buf32.to(torch.bfloat16), buf)
    fr = feats.to(torch.bfloat16).float()
    Wr = W.to(torch.bfloat16).float().requires_grad_(True)
    br_ = bias.clone().requires_grad_(True)
    gr, ber = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(fr @ Wr.t() + br_, (H,), gr, ber, 1e-12)
    assert nerr(buf[:, T:], yr) < 1.5e-2
    dy = torch.zeros(N, S, H, dtype=torch.bfloat16, device=DEV)
    dy[:, T:] = rnd(N, R, H, seed=7)
    yr.backward(dy[:, T:].float())
    dpre = torch.empty(rows, H, dtype=torch.bfloat16, device=DEV)
    dWp = torch.zeros(H, Dpad, dtype=torch.float32, device=DEV)
    dbias, dg, db = (torch.zeros(H, device=DEV) for _ in range(3))
    ops.img_embed_ln_bwd(dy, ymap, packed, pre, mean, rstd, gamma, 1, dpre, dWp, dbias, dg, db, rows, Dpad, H)
    assert nerr(dWp[:, :D], Wr.grad) < 2e-2
    assert nerr(dbias, br_.grad) < 2e-2 and nerr(dg, gr.grad) < 1e-2 and nerr(db, ber.grad) < 1e-2
    dW = torch.zeros(H, D, device=DEV)
    ops.accum_f32_2d(dW, D, dWp, Dpad, H, D)
    assert torch.equal(dW, dWp[:, :D])


# ------------------------------------------------------------------------------- attention
def _chunk_mask(N, T, lens, seed):
    g = torch.Generator().manual_seed(seed)
    cm = torch.zeros(N, T, T)
    for n in range(N):
        L = int(lens[n])
        m = torch.eye(L)
        m[0], m[L - 1] = 1.0, 1.0
        t = 1
        while t <= L - 2:
            run = min(int(torch.randint(1, 5, (1,), generator=g)), L - 1 - t)
            m[t:t + run, t:t + run] = 1.0
            t += run
        cm[n, :L, :L] = m
    return cm


def _dense_mask(mode, N, S, T, P, key_mask, cm):
    """additive [N,1,S,P+S] mask exactly as the reference builds it (v10:171-232, modeling_transfomres.py:628-641)."""
    ext = ((1.0 - key_mask) * NEG)[:, None, None, :]
    if mode == 0:
        return ext.expand(N, 1, S, P + S).clone()
    R = S - T
    am = ext.repeat(1, 1, S, 1)
    am[:, :, :T, :T] = ((1.0 - cm) * NEG)[:, None]
    am[:, :, T:, :T] = NEG
    if mode == 2:
        im = torch.cat((torch.zeros(N, R, T, device=cm.device), torch.eye(R, device=cm.device)[None].repeat(N, 1, 1)), -1)
        am[:, :, T:, :] = ((1.0 - im) * NEG)[:, None]
    return am


def _attn_ref(qkv, pkv, am, heads, T, dctx, dalign):
    N, S, H3 = qkv.shape
    H = H3 // 3
    q, k, v = qkv.float().split(H, dim=-1)
    if pkv is not None:
        pk, pv = pkv.float().split(H, dim=-1)
        k, v = torch.cat([pk, k], 1), torch.cat([pv, v], 1)

    def heads_(t):
        return t.view(N, -1, heads, 64).permute(0, 2, 1, 3)
    scores = heads_(q) @ heads_(k).transpose(-1, -2) / 8.0 + am
    probs = torch.softmax(scores, -1)
    ctx = (probs @ heads_(v)).permute(0, 2, 1, 3).reshape(N, S, H)
    loss = (ctx * dctx.float()).sum()
    if dalign is not None:
        loss = loss + (probs[:, :, :T, T:].sum(1) * dalign).sum()
    return ctx, probs, loss, torch.logsumexp(scores.detach(), -1)


@pytest.mark.parametrize("mode,P,S,T", [(0, 0, 178, 128), (0, 10, 178, 128), (1, 0, 178, 128), (2, 0, 178, 128),
                                        (0, 64, 256, 0), (2, 0, 300, 200), (0, 10, 300, 200), (0, 0, 40, 0),
                                        (0, 3, 130, 0),
                                        # BASELINE.json configs[3]: prefix-length sweep 0/10/32/64 at sequence 256
                                        (0, 0, 256, 0), (0, 10, 256, 0), (0, 32, 256, 0)])
def test_prefix_attention_fwd_bwd(mode, P, S, T):
    N, heads = 3, 4
    H = heads * 64
    g = torch.Generator().manual_seed(mode * 100 + P + S)
    qkv = rnd(N, S, 3 * H, seed=11 + mode)
    pkv = rnd(N, P, 2 * H, seed=12) if P > 0 else None
    key_mask = torch.ones(N, P + S)
    lens = [S] * N
    if T > 0:
        lens = [T, max(T // 2, 3), max(T - 7, 3)]
        for n in range(N):
            key_mask[n, P + lens[n]:P + T] = 0.0
        key_mask[1, P + S - 5:] = 0.0       # some padded regions on one sequence
    else:
        key_mask[1, P + S - 9:] = 0.0
    key_mask = key_mask.to(DEV)
    cm = _chunk_mask(N, T, lens, 5).to(DEV) if mode != 0 else None
    cwords = (T + 31) // 32 if mode != 0 else 0
    kbias = torch.empty(N, P + S, device=DEV)
    ops.key_bias(key_mask, kbias)
    assert torch.equal(kbias, (1.0 - key_mask) * NEG)
    cbits = None
    if mode != 0:
        cbits = torch.zeros(N, T, cwords, dtype=torch.int32, device=DEV)
        ops.chunk_bits(cm, cbits, N, T, cwords)
        # bit-exact packing check
        unpacked = ((cbits.unsqueeze(-1) >> torch.arange(32, device=DEV)) & 1).reshape(N, T, cwords * 32)[:, :, :T]
        assert torch.equal(unpacked.float(), cm)
    want_align = mode == 2
    R = S - T
    ctx = torch.empty(N, S, H, dtype=torch.bfloat16, device=DEV)
    lse = torch.empty(N, heads, S, device=DEV)
    probs_img = torch.zeros(N, heads, T, R, device=DEV) if want_align else None
    ops.prefix_attn_fwd(N, S, P, heads, T if (mode != 0 or want_align) else 0, mode, qkv, pkv, kbias, cbits, cwords,
                        ctx, lse, probs_img)
    dctx = rnd(N, S, H, seed=13)
    dalign = (torch.randn(N, T, R, generator=g) * 0.5).to(DEV) if want_align else None
    qkv_r = qkv.float().requires_grad_(True)
    pkv_r = pkv.float().requires_grad_(True) if P > 0 else None
    am = _dense_mask(mode, N, S, T, P, key_mask, cm)
    ctx_r, probs_r, loss, lse_r = _attn_ref(qkv_r, pkv_r, am, heads, T, dctx, dalign)
    assert nerr(ctx, ctx_r) < 1.5e-2
    torch.testing.assert_close(lse, lse_r, atol=2e-3, rtol=1e-4)
    if want_align:
        torch.testing.assert_close(probs_img, probs_r[:, :, :T, T:].detach(), atol=2e-5, rtol=1e-3)
        # exact zeros where the reference has exact zeros (masked_fill(A == 0) relies on it)
        assert torch.equal(probs_img == 0, probs_r[:, :, :T, T:] == 0)
    loss.backward()
    dqkv = torch.zeros(N, S, 3 * H, dtype=torch.bfloat16, device=DEV)
    dpkv = torch.zeros(N, P, 2 * H, dtype=torch.bfloat16, device=DEV) if P > 0 else None
    ws = torch.empty(ops.attn_bwd_workspace_bytes(N, S, heads) // 4, device=DEV)
    ops.prefix_attn_bwd(N, S, P, heads, T if (mode != 0 or want_align) else 0, mode, qkv, pkv, kbias, cbits, cwords,
                        ctx, lse, probs_img, dctx, dalign, dqkv, dpkv, ws)
    gq, gk, gv = qkv_r.grad.split(H, -1)
    dq, dk, dv = dqkv.split(H, -1)
    assert nerr(dv, gv) < 2e-2, "dV"
    assert nerr(dk, gk) < 2e-2, "dK"
    assert nerr(dq, gq) < 2e-2, "dQ"
    if P > 0:
        assert nerr(dpkv, pkv_r.grad) < 2e-2, "dprefix"


def test_prefix_attention_row_selection_of_the_alignment_block():
    """``row_sel`` (= align_pos): the mode-2 kernels emit / consume the per-head text->region probabilities of the selected
    rows only.  Selected rows are bit-identical to the dense run, unselected rows of the buffer are never written, and the
    backward pass equals the dense one fed a gradient that is zero on the unselected rows."""
    N, heads, S, T = 3, 4, 178, 128
    H, R = heads * 64, S - T
    qkv = rnd(N, S, 3 * H, seed=21)
    key_mask = torch.ones(N, S, device=DEV)
    lens = [T, T // 2, T - 7]
    for n in range(N):
        key_mask[n, lens[n]:T] = 0.0
    cm = _chunk_mask(N, T, lens, 5).to(DEV)
    cwords = (T + 31) // 32
    kbias = torch.empty(N, S, device=DEV)
    ops.key_bias(key_mask, kbias)
    cbits = torch.zeros(N, T, cwords, dtype=torch.int32, device=DEV)
    ops.chunk_bits(cm, cbits, N, T, cwords)
    g = torch.Generator().manual_seed(5)
    sel = (torch.rand(N, T, generator=g) < 0.04).to(torch.int64)
    sel[0, 3] = 1
    sel = sel.to(DEV)
    outs = {}
    for name, rs in (("dense", None), ("sparse", sel)):
        ctx = torch.empty(N, S, H, dtype=torch.bfloat16, device=DEV)
        lse = torch.empty(N, heads, S, device=DEV)
        probs = torch.full((N, heads, T, R), -7.0, device=DEV)
        ops.prefix_attn_fwd(N, S, 0, heads, T, 2, qkv, None, kbias, cbits, cwords, ctx, lse, probs, row_sel=rs)
        dctx = rnd(N, S, H, seed=13)
        dalign = (torch.randn(N, T, R, generator=g) * 0.5).to(DEV) * sel.unsqueeze(-1).float()
        dqkv = torch.zeros(N, S, 3 * H, dtype=torch.bfloat16, device=DEV)
        ws = torch.empty(ops.attn_bwd_workspace_bytes(N, S, heads) // 4, device=DEV)
        if name == "sparse":
            dalign_in = outs["dense"][4]
        else:
            dalign_in = dalign
        ops.prefix_attn_bwd(N, S, 0, heads, T, 2, qkv, None, kbias, cbits, cwords, ctx, lse, probs, dctx, dalign_in, dqkv,
                            None, ws, row_sel=rs)
        outs[name] = (ctx, lse, probs, dqkv, dalign_in)
    d, s_ = outs["dense"], outs["sparse"]
    assert torch.equal(d[0], s_[0]) and torch.equal(d[1], s_[1])
    m4 = sel.bool().unsqueeze(1).expand(N, heads, T)
    assert torch.equal(d[2][m4], s_[2][m4])                       # selected rows: identical
    assert bool((s_[2][~m4] == -7.0).all())                       # the others were never written
    assert nerr(s_[3], d[3]) < 1e-6 or torch.equal(s_[3], d[3])   # same gradient
    # and the head-summed block of the selected rows only
    blk_d = torch.empty(N, T, R, device=DEV)
    blk_s = torch.empty(N, T, R, device=DEV)
    ops.align_softmax_ce_fwd(d[2], 1, heads, None, None, blk_d, None, None, N, T, R)
    ops.align_softmax_ce_fwd(s_[2], 1, heads, None, None, blk_s, None, None, N, T, R, row_sel=sel)
    assert torch.equal(blk_d[sel.bool()], blk_s[sel.bool()]) and float(blk_s[~sel.bool()].abs().max()) == 0.0
    torch.testing.assert_close(blk_d, d[2].sum(1), atol=1e-6, rtol=1e-6)


def test_chunk_mean_q_matches_index_add_gather():
    N, T, S, H = 4, 37, 50, 256
    g = torch.Generator().manual_seed(3)
    qkv = rnd(N, S, 3 * H, seed=4)
    ref = qkv.clone().float()
    chunk_id = torch.full((N, T), -1, dtype=torch.int32)
    max_chunks = 1
    for n in range(N):
        L = int(torch.randint(6, T + 1, (1,), generator=g))
        ids, t, c = [], 1, 0
        while t <= L - 2:
            run = min(int(torch.randint(1, 5, (1,), generator=g)), L - 1 - t)
            ids += [c] * run
            t += run
            c += 1
        gi = torch.tensor(ids)
        chunk_id[n, 1:1 + len(ids)] = gi.int()
        max_chunks = max(max_chunks, c)
        q = ref[n, 1:1 + len(ids), :H]
        cs = torch.zeros(c, H, device=DEV).index_add(0, gi.to(DEV), q)
        cnt = torch.bincount(gi, minlength=c).to(DEV).float()
        ref[n, 1:1 + len(ids), :H] = (cs / cnt[:, None])[gi.to(DEV)]
    before = qkv.clone()
    ops.chunk_mean_q(qkv, 3 * H, N, S, H, chunk_id.to(DEV), T, max_chunks)
    assert nerr(qkv[..., :H], ref[..., :H]) < 1e-2
    assert torch.equal(qkv[..., H:], before[..., H:])                     # K / V untouched
    untouched = (chunk_id < 0).to(DEV)
    assert torch.equal(qkv[:, :T, :H][untouched], before[:, :T, :H][untouched])   # CLS / SEP / pad rows bit-exact
    assert torch.equal(qkv[:, T:], before[:, T:])                         # region rows bit-exact


# ------------------------------------------------------------------------------- heads
def test_align_softmax_ce_fwd_bwd():
    nl, N, heads, T, R = 3, 5, 12, 20, 50
    g = torch.Generator().manual_seed(0)
    probs = torch.rand(nl, N, heads, T, R, generator=g)
    probs[:, :, :, :, 40:] = 0.0          # padded regions: exact zeros -> masked_fill branch
    probs = (probs / probs.sum(-1, keepdim=True).clamp_min(1e-6) * 0.3).to(DEV)
    total_label = torch.zeros(N, T, dtype=torch.int64)
    for n in range(N):
        pos = torch.randperm(T - 2, generator=g)[:4] + 1
        total_label[n, pos] = torch.randint(1, 40, (4,), generator=g)
    align_pos = (total_label != 0).long()
    total_label, align_pos = total_label.to(DEV), align_pos.to(DEV)
    block = torch.empty(N, T, R, device=DEV)
    row_loss = torch.empty(N * T, 2, device=DEV)
    out = torch.empty(4, device=DEV)
    ops.align_softmax_ce_fwd(probs, nl, heads, align_pos, total_label, block, row_loss, out, N, T, R)
    pr = probs.clone().requires_grad_(True)
    blk = pr.permute(1, 0, 2, 3, 4).sum(1).sum(1)
    w = torch.softmax(blk.masked_fill(blk == 0, -1e5), -1)
    wa, tl = w[align_pos == 1], total_label[align_pos == 1]
    loss = torch.nn.functional.cross_entropy(wa, tl)
    torch.testing.assert_close(block, blk.detach(), atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(out[0], loss.detach(), atol=1e-5, rtol=1e-5)
    assert int(out[1]) == int(tl.numel()) and int(out[2]) == int((wa.argmax(-1) == tl).sum())
    loss.backward()
    dblock = torch.empty(N, T, R, device=DEV)
    ops.align_softmax_ce_bwd(block, align_pos, total_label, out, 1.0, dblock, N, T, R)
    torch.testing.assert_close(dblock, pr.grad[0, :, 0], atol=1e-7, rtol=1e-3)


@pytest.mark.parametrize("L,H,dt", [(381, 768, torch.bfloat16), (57, 1024, torch.bfloat16), (381, 768, torch.float32)])
def test_cls_pool_attn_fwd_bwd(L, H, dt):
    N = 6
    q, k = rnd(N, H, scale=0.3, seed=1, dtype=dt), rnd(N, L, H, scale=0.3, seed=2, dtype=dt)
    mask = torch.zeros(N, L)
    mask[:, L - 11:] = NEG
    mask = mask.to(DEV)
    w = torch.empty(N, L, device=DEV)
    pooled = torch.empty(N, H, dtype=torch.bfloat16, device=DEV)
    ops.cls_pool_attn_fwd(q, k, mask, w, pooled, N, L, H)
    qr, kr = q.float().requires_grad_(True), k.float().requires_grad_(True)
    wr = torch.softmax((qr[:, None] @ kr.transpose(1, 2)) + mask[:, None], -1)
    pr = (wr @ kr).squeeze(1)
    torch.testing.assert_close(w, wr.squeeze(1).detach(), atol=1e-5, rtol=1e-3)
    assert nerr(pooled, pr) < 1e-2
    dp = rnd(N, H, seed=3)
    pr.backward(dp.float())
    dq = torch.empty(N, H, dtype=torch.bfloat16, device=DEV)
    dk = torch.empty(N, L, H, dtype=torch.bfloat16, device=DEV)
    ops.cls_pool_attn_bwd(dp, q, k, w, dq, dk, N, L, H)
    assert nerr(dq, qr.grad) < 1.5e-2 and nerr(dk, kr.grad) < 1.5e-2


@pytest.mark.parametrize("L,H,heads", [(381, 768, 8), (69, 768, 8), (7, 768, 8), (130, 1024, 8)])
def test_cross_attn_q1_matches_cross_attention_lyx(L, H, heads):
    """cross_attention_lyx with a single query (v10:692-797): q scaled by head_dim^-0.5, softmax over L per head."""
    N = 5
    q, k, v = rnd(N, H, seed=21), rnd(N, L, H, seed=22), rnd(N, L, H, seed=23)
    out = torch.empty(N, H, dtype=torch.bfloat16, device=DEV)
    ops.cross_attn_q1_fwd(q, k, v, out, N, L, H, heads)
    hd = H // heads
    qf = (q.float() * hd ** -0.5).view(N, heads, 1, hd)
    kf = k.float().view(N, L, heads, hd).permute(0, 2, 1, 3)
    vf = v.float().view(N, L, heads, hd).permute(0, 2, 1, 3)
    w = torch.softmax(qf @ kf.transpose(-1, -2), dim=-1)
    ref = (w @ vf).permute(0, 2, 1, 3).reshape(N, H)
    assert nerr(out, ref) < 1e-2


def test_mc_scorer_ce_hard_labels_binary_to_mp():
    B, nc, H = 8, 4, 768
    rows = B * nc
    x = rnd(rows, H, seed=1)
    W = (0.05 * torch.randn(2, H)).to(DEV)
    b = (0.05 * torch.randn(2)).to(DEV)
    correct = torch.randint(0, nc, (B,))
    label = torch.zeros(rows, dtype=torch.int64)
    label[torch.arange(B) * nc + correct] = 1
    label = label.to(DEV)
    logits = torch.empty(rows, 2, device=DEV)
    loss = torch.empty(1, device=DEV)
    mp = torch.empty(rows, device=DEV)
    pred = torch.empty(B, dtype=torch.int64, device=DEV)
    matched = torch.empty(B, dtype=torch.uint8, device=DEV)
    ops.mc_scorer_ce_fwd(x, W, b, rows, H, 2, 2, label, None, logits, loss, mp, pred, matched, nc)
    xr = x.float().requires_grad_(True)
    Wr, br = W.clone().requires_grad_(True), b.clone().requires_grad_(True)
    lr = xr @ Wr.t() + br
    loss_r = torch.nn.functional.cross_entropy(lr, label)
    mp_r = torch.softmax(lr, 1)[:, 1].reshape(-1, nc)
    torch.testing.assert_close(logits, lr.detach(), atol=1e-4, rtol=1e-4)
    torch.testing.assert_close(loss[0], loss_r.detach(), atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(mp.view(-1, nc), mp_r.detach(), atol=1e-5, rtol=1e-4)
    assert torch.equal(pred, mp_r.argmax(-1))
    assert torch.equal(matched.bool(), mp_r.argmax(-1) == label.view(-1, nc).argmax(-1))
    loss_r.backward()
    dx = torch.empty(rows, H, dtype=torch.bfloat16, device=DEV)
    dW, db = torch.zeros(2, H, device=DEV), torch.zeros(2, device=DEV)
    ops.mc_scorer_ce_bwd(x, W, b, rows, H, 2, 2, label, None, logits, 1.0, dx, dW, db)
    assert nerr(dx, xr.grad) < 1e-2 and nerr(dW, Wr.grad) < 1e-4 and nerr(db, br.grad) < 1e-4


def test_mc_scorer_ce_soft_labels_and_choice_flattened():
    B, nc, H = 8, 4, 768
    rows = B * nc
    x = rnd(rows, H, seed=1)
    # Abstract_Specific: Linear(H -> 1).view(-1, 4), CE with a float one-hot target
    W = (0.05 * torch.randn(1, H)).to(DEV)
    b = (0.05 * torch.randn(1)).to(DEV)
    onehot = torch.zeros(B, nc)
    onehot[torch.arange(B), torch.randint(0, nc, (B,))] = 1.0
    onehot = onehot.to(DEV)
    logits = torch.empty(rows, 1, device=DEV)
    loss = torch.empty(1, device=DEV)
    ops.mc_scorer_ce_fwd(x, W, b, rows, H, 1, nc, None, onehot, logits, loss)
    xr = x.float().requires_grad_(True)
    lr = (xr @ W.t() + b).view(-1, nc)
    loss_r = torch.nn.functional.cross_entropy(lr, onehot)
    torch.testing.assert_close(loss[0], loss_r.detach(), atol=1e-5, rtol=1e-5)
    loss_r.backward()
    dx = torch.empty(rows, H, dtype=torch.bfloat16, device=DEV)
    dW, db = torch.zeros(1, H, device=DEV), torch.zeros(1, device=DEV)
    ops.mc_scorer_ce_bwd(x, W, b, rows, H, 1, nc, None, onehot, logits, 1.0, dx, dW, db)
    assert nerr(dx, xr.grad) < 1e-2
    # ImageBertForMultipleChoice: pooled.view(B, 4H) -> Linear(4H -> num_labels) -> CE
    W4 = (0.02 * torch.randn(nc, nc * H)).to(DEV)
    lab = torch.randint(0, nc, (B,)).to(DEV)
    logits4 = torch.empty(B, nc, device=DEV)
    ops.mc_scorer_ce_fwd(x, W4, None, B, nc * H, nc, nc, lab, None, logits4, loss)
    lr4 = x.float().view(B, nc * H) @ W4.t()
    torch.testing.assert_close(logits4, lr4, atol=2e-4, rtol=1e-4)
    torch.testing.assert_close(loss[0], torch.nn.functional.cross_entropy(lr4, lab), atol=1e-5, rtol=1e-5)


def test_scorer_bce_relu_rowln():
    """loss_type 'bce' (instance_bce_with_logits, modeling/modeling_bert.py:416-421) of the fused scorer, and the pieces of
    the 'mlp' classifier (:511-528): ReLU in the GEMM epilogue, relu_bwd, row LayerNorm over hidden*cls_hidden_scale."""
    F = torch.nn.functional
    torch.manual_seed(3)
    B, K, G = 6, 1536, 4
    x = rnd(B, K, seed=1)
    W = (0.03 * torch.randn(G, K)).to(DEV)
    b = (0.05 * torch.randn(G)).to(DEV)
    tgt = torch.rand(B, G).to(DEV)
    logits, loss = torch.empty(B, G, device=DEV), torch.empty(1, device=DEV)
    ops.mc_scorer_ce_fwd(x, W, b, B, K, G, G, None, tgt, logits, loss, loss_kind=1)
    xr = x.float().requires_grad_(True)
    Wr, br = W.clone().requires_grad_(True), b.clone().requires_grad_(True)
    lr = xr @ Wr.t() + br
    loss_r = F.binary_cross_entropy_with_logits(lr, tgt) * G
    torch.testing.assert_close(logits, lr.detach(), atol=2e-4, rtol=1e-4)
    torch.testing.assert_close(loss[0], loss_r.detach(), atol=1e-5, rtol=1e-5)
    loss_r.backward()
    dx = torch.empty(B, K, dtype=torch.bfloat16, device=DEV)
    dW, db = torch.zeros(G, K, device=DEV), torch.zeros(G, device=DEV)
    ops.mc_scorer_ce_bwd(x, W, b, B, K, G, G, None, tgt, logits, 1.0, dx, dW, db, loss_kind=1)
    assert nerr(dx, xr.grad) < 1e-2 and nerr(dW, Wr.grad) < 1e-4 and nerr(db, br.grad) < 1e-4
    with pytest.raises(RuntimeError, match="loss_kind"):
        ops.mc_scorer_ce_fwd(x, W, b, B, K, G, G, torch.zeros(B, dtype=torch.int64, device=DEV), None, logits, loss,
                             loss_kind=1)
    # ReLU epilogue (few rows: M = 6 questions against a 128-row tile) + relu_bwd
    Hin, Hm = 3072, 1536
    xin, W1 = rnd(B, Hin, seed=4), rnd(Hm, Hin, scale=0.02, seed=5)
    b1 = rnd(Hm, seed=6, dtype=torch.float32)
    h = torch.empty(B, Hm, dtype=torch.bfloat16, device=DEV)
    ops.gemm(xin, W1, h, B, Hm, Hin, bias=b1, act=3)
    href = torch.relu(xin.float() @ W1.float().t() + b1)
    assert nerr(h, href) < 1e-2 and float(h.float().min()) >= 0.0
    dy = rnd(B, Hm, seed=7)
    dpre = torch.empty_like(dy)
    ops.relu_bwd(dy, h, dpre)
    assert torch.equal(dpre, torch.where(h.float() > 0, dy.float(), torch.zeros_like(dy.float())).to(torch.bfloat16))
    # row LayerNorm forward / backward
    gamma = (1.0 + 0.1 * torch.randn(Hm)).to(DEV)
    beta = (0.1 * torch.randn(Hm)).to(DEV)
    z = torch.empty_like(h)
    mean, rstd = torch.empty(B, device=DEV), torch.empty(B, device=DEV)
    ops.rowln_fwd(h, gamma, beta, z, mean, rstd, B, Hm, 1e-12)
    hr = h.float().requires_grad_(True)
    gr, ber = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    zr = F.layer_norm(hr, (Hm,), gr, ber, 1e-12)
    assert nerr(z, zr) < 1e-2
    torch.testing.assert_close(mean, hr.detach().mean(-1), atol=1e-5, rtol=1e-4)
    (zr * dy.float()).sum().backward()
    dh = torch.empty_like(h)
    dg, dbt = torch.zeros(Hm, device=DEV), torch.zeros(Hm, device=DEV)
    ops.rowln_bwd(dy, h, mean, rstd, gamma, dh, dg, dbt, B, Hm)
    assert nerr(dh, hr.grad) < 1e-2 and nerr(dg, gr.grad) < 1e-3 and nerr(dbt, ber.grad) < 1e-3


def test_copy_rows_colsum_add_tanh():
    N, S, T, H = 3, 9, 6, 256
    a, b = rnd(N, S, H, seed=1), rnd(N, S, H, seed=2)
    X = torch.zeros(N, 3 * (T - 1), H, dtype=torch.bfloat16, device=DEV)
    ops.copy_rows(a, S, 1, X, 3 * (T - 1), 0, N, T - 1, H)
    ops.copy_rows(b, S, 1, X, 3 * (T - 1), T - 1, N, T - 1, H)
    assert torch.equal(X[:, :T - 1], a[:, 1:T]) and torch.equal(X[:, T - 1:2 * (T - 1)], b[:, 1:T])
    acc = a.clone()
    ops.copy_rows(X, 3 * (T - 1), T - 1, acc, S, 1, N, T - 1, H, accumulate=True)
    assert nerr(acc[:, 1:T], a[:, 1:T].float() + b[:, 1:T].float()) < 1e-2
    cs = torch.zeros(H, device=DEV)
    ops.colsum(a, H, N * S, H, cs)
    assert nerr(cs, a.float().view(-1, H).sum(0)) < 1e-4
    y = torch.empty_like(a)
    ops.add_bf16(a, b, y)
    assert torch.equal(y, (a.float() + b.float()).to(torch.bfloat16))
    t = torch.tanh(a.float()).to(torch.bfloat16)
    dx = torch.empty_like(a)
    ops.tanh_bwd(b, t, dx)
    assert nerr(dx, b.float() * (1 - t.float() ** 2)) < 1e-2


def test_errors_are_reported_not_swallowed():
    x = rnd(8, 100, seed=1)
    with pytest.raises(RuntimeError, match="multiple of 256"):
        ops.layernorm_fwd(x, x, x, x, None, None, 8, 100, 1e-12)
    with pytest.raises(RuntimeError):
        ops.prefix_attn_fwd(1, 600, 0, 2, 0, 0, x, None, x, None, 0, x, x)


# ------------------------------------------------------------------------------- dropout (training mode)
def host_keep(seed, site, idx, p):
    """Host restatement of the counter-based dropout mask (csrc/ptx.cuh drop_hash / drop_keep), exact."""
    import numpy as np
    M = np.uint64(0xFFFFFFFF)
    x = ((idx.astype(np.uint64) ^ np.uint64(seed & 0xFFFFFFFF)) * np.uint64(0x9E3779B1)) & M
    x ^= np.uint64((site * 0x85EBCA77) & 0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & M
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & M
    x ^= x >> np.uint64(16)
    thr = max(1, min(0xFFFFFFFF, int(float(np.float32(p)) * 4294967296.0)))
    return torch.from_numpy((x >= np.uint64(thr)))


def host_keep_attn(seed, site, rows, kid, p):
    """Two-level mask of the attention kernels (csrc/ptx.cuh drop_keep_row): row hash, then one round on the key id."""
    import numpy as np
    M = np.uint64(0xFFFFFFFF)
    x = ((rows.astype(np.uint64) ^ np.uint64(seed & 0xFFFFFFFF)) * np.uint64(0x9E3779B1)) & M
    x ^= np.uint64((site * 0x85EBCA77) & 0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & M
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & M
    x ^= x >> np.uint64(16)
    y = x ^ ((kid.astype(np.uint64) * np.uint64(0x9E3779B1)) & M)
    y ^= y >> np.uint64(15)
    y = (y * np.uint64(0x2C1B3C6D)) & M
    y ^= y >> np.uint64(13)
    thr = max(1, min(0xFFFFFFFF, int(float(np.float32(p)) * 4294967296.0)))
    return torch.from_numpy((y >= np.uint64(thr)))


def _seed_tensor(v):
    return torch.tensor([v], dtype=torch.int32, device=DEV)


def test_dropout_kernel_matches_host_mask():
    import numpy as np
    n, p, site, seedv = 64 * 768, 0.3, 12345, 987654321
    x16 = rnd(n, seed=31)
    x32 = torch.randn(n, generator=torch.Generator().manual_seed(32)).to(DEV)
    res = torch.randn(n, generator=torch.Generator().manual_seed(33)).to(DEV)
    keep = host_keep(seedv, site, np.arange(n), p).to(DEV)
    assert abs(float(keep.float().mean()) - (1 - p)) < 0.01
    out16 = torch.empty_like(x16)
    io32 = x32.clone()
    ops.dropout(x16, out16, io32, res, n, _seed_tensor(seedv), site, p)
    scale = 1.0 / (1.0 - p)
    assert torch.equal(out16 == 0, ~keep | (x16 == 0))
    assert nerr(out16, x16.float() * keep * scale) < 1e-2
    torch.testing.assert_close(io32, x32 * keep * scale + res, atol=1e-5, rtol=1e-5)
    other = torch.empty_like(x16)
    ops.dropout(x16, other, None, None, n, _seed_tensor(seedv), site + 1, p)      # another site: another mask
    assert float(((other == 0) != (out16 == 0)).float().mean()) > 0.2
    again = torch.empty_like(x16)
    ops.dropout(x16, again, None, None, n, _seed_tensor(seedv), site, p)          # same (seed, site): same mask
    assert torch.equal(again, out16)


@pytest.mark.parametrize("mode,P,S,T", [(0, 10, 178, 128), (2, 0, 178, 128), (0, 0, 300, 200), (0, 3, 130, 0)])
def test_prefix_attention_dropout_fwd_bwd(mode, P, S, T):
    """softmax -> dropout -> P V with the host-restated mask (persistent and tiled kernels use different key-row
    orders in shared memory but the same logical mask), forward and backward."""
    import numpy as np
    N, heads, p, site, seedv = 2, 4, 0.1, 777, 13579
    H = heads * 64
    qkv = rnd(N, S, 3 * H, seed=41 + mode)
    pkv = rnd(N, P, 2 * H, seed=42) if P > 0 else None
    key_mask = torch.ones(N, P + S)
    lens = [S] * N
    if T > 0:
        lens = [T, max(T // 2, 3)]
        for n in range(N):
            key_mask[n, P + lens[n]:P + T] = 0.0
    key_mask = key_mask.to(DEV)
    cm = _chunk_mask(N, T, lens, 5).to(DEV) if mode != 0 else None
    cwords = (T + 31) // 32 if mode != 0 else 0
    kbias = torch.empty(N, P + S, device=DEV)
    ops.key_bias(key_mask, kbias)
    cbits = None
    if mode != 0:
        cbits = torch.zeros(N, T, cwords, dtype=torch.int32, device=DEV)
        ops.chunk_bits(cm, cbits, N, T, cwords)
    want_align = mode == 2
    R = S - T
    seed = _seed_tensor(seedv)
    ctx = torch.empty(N, S, H, dtype=torch.bfloat16, device=DEV)
    lse = torch.empty(N, heads, S, device=DEV)
    probs_img = torch.zeros(N, heads, T, R, device=DEV) if want_align else None
    Tk = T if (mode != 0 or want_align) else 0
    ops.prefix_attn_fwd(N, S, P, heads, Tk, mode, qkv, pkv, kbias, cbits, cwords, ctx, lse, probs_img, drop=(seed, site, p))
    # host mask in the reference's key order [prefix ; sequence]: logical id of prefix key j is S + j, of sequence key j is j
    K = P + S
    kid = torch.cat([S + torch.arange(P), torch.arange(S)])
    rows = torch.arange(N * heads * S).view(N, heads, S, 1).expand(N, heads, S, K).numpy()
    kids = kid.view(1, 1, 1, K).expand(N, heads, S, K).numpy()
    keep = host_keep_attn(seedv, site, rows, kids, p).to(DEV).float() / (1.0 - p)
    qkv_r = qkv.float().requires_grad_(True)
    pkv_r = pkv.float().requires_grad_(True) if P > 0 else None
    am = _dense_mask(mode, N, S, T, P, key_mask, cm)
    q, k, v = qkv_r.split(H, -1)
    if P > 0:
        pk, pv = pkv_r.split(H, -1)
        k, v = torch.cat([pk, k], 1), torch.cat([pv, v], 1)
    hsplit = lambda t: t.view(N, -1, heads, 64).permute(0, 2, 1, 3)  # noqa: E731
    scores = hsplit(q) @ hsplit(k).transpose(-1, -2) * 0.125 + am
    probs = torch.softmax(scores, -1) * keep
    ctx_r = (probs @ hsplit(v)).permute(0, 2, 1, 3).reshape(N, S, H)
    assert nerr(ctx, ctx_r) < 1.5e-2
    if want_align:
        torch.testing.assert_close(probs_img, probs[:, :, :T, T:].detach(), atol=3e-5, rtol=2e-3)
    dctx = rnd(N, S, H, seed=43)
    g = torch.Generator().manual_seed(5)
    dalign = (torch.randn(N, T, R, generator=g) * 0.5).to(DEV) if want_align else None
    loss = (ctx_r * dctx.float()).sum()
    if dalign is not None:
        loss = loss + (probs[:, :, :T, T:].sum(1) * dalign).sum()
    loss.backward()
    dqkv = torch.zeros(N, S, 3 * H, dtype=torch.bfloat16, device=DEV)
    dpkv = torch.zeros(N, P, 2 * H, dtype=torch.bfloat16, device=DEV) if P > 0 else None
    ws = torch.empty(ops.attn_bwd_workspace_bytes(N, S, heads) // 4, device=DEV)
    ops.prefix_attn_bwd(N, S, P, heads, Tk, mode, qkv, pkv, kbias, cbits, cwords, ctx, lse, probs_img, dctx, dalign, dqkv,
                        dpkv, ws, drop=(seed, site, p))
    gq, gk, gv = qkv_r.grad.split(H, -1)
    dq, dk, dv = dqkv.split(H, -1)
    assert nerr(dv, gv) < 2e-2, "dV"
    assert nerr(dk, gk) < 2e-2, "dK"
    assert nerr(dq, gq) < 2e-2, "dQ"
    if P > 0:
        assert nerr(dpkv, pkv_r.grad) < 2e-2, "dprefix"


def test_head_attention_dropout_fwd_bwd():
    """ClsLayer2 pooling (weights dropout, v10:828) and cross_attention_lyx (probability dropout, v10:785) with the
    host-restated masks, forward and backward."""
    import numpy as np
    N, L, H, heads, p, site, seedv = 3, 69, 768, 8, 0.3, 4242, 2468
    seed = _seed_tensor(seedv)
    # ---- cls_pool
    q = torch.randn(N, H, generator=torch.Generator().manual_seed(1)).mul(0.05).to(DEV).requires_grad_(True)
    k = torch.randn(N, L, H, generator=torch.Generator().manual_seed(2)).to(DEV).requires_grad_(True)
    mask = torch.zeros(N, L, device=DEV)
    mask[1, L - 9:] = NEG
    w = torch.empty(N, L, device=DEV)
    pooled = torch.empty(N, H, dtype=torch.bfloat16, device=DEV)
    ops.cls_pool_attn_fwd(q.detach(), k.detach(), mask, w, pooled, N, L, H, drop=(seed, site, p))
    keep = host_keep(seedv, site, np.arange(N * L), p).view(N, L).to(DEV).float() / (1 - p)
    w_r = torch.softmax((q.unsqueeze(1) @ k.transpose(1, 2)).squeeze(1) + mask, -1)
    pooled_r = ((w_r * keep).unsqueeze(1) @ k).squeeze(1)
    torch.testing.assert_close(w, w_r.detach(), atol=1e-5, rtol=1e-4)          # returned weights are undropped
    assert nerr(pooled, pooled_r) < 1e-2
    dpool = rnd(N, H, seed=3)
    (pooled_r * dpool.float()).sum().backward()
    dq = torch.empty(N, H, dtype=torch.bfloat16, device=DEV)
    dk = torch.empty(N, L, H, dtype=torch.bfloat16, device=DEV)
    ops.cls_pool_attn_bwd(dpool, q.detach(), k.detach(), w, dq, dk, N, L, H, drop=(seed, site, p))
    assert nerr(dq, q.grad) < 2e-2 and nerr(dk, k.grad) < 2e-2
    # ---- cross attention
    q2, k2, v2 = rnd(N, H, seed=21), rnd(N, L, H, seed=22), rnd(N, L, H, seed=23)
    out = torch.empty(N, H, dtype=torch.bfloat16, device=DEV)
    lse = torch.empty(N, heads, device=DEV)
    ops.cross_attn_q1_fwd(q2, k2, v2, out, N, L, H, heads, lse, drop=(seed, site + 1, 0.1))
    hd = H // heads
    keep2 = host_keep(seedv, site + 1, np.arange(N * heads * L), 0.1).view(N, heads, 1, L).to(DEV).float() / 0.9
    qf = q2.float().requires_grad_(True)
    kf, vf = k2.float().requires_grad_(True), v2.float().requires_grad_(True)
    sc = (qf * hd ** -0.5).view(N, heads, 1, hd) @ kf.view(N, L, heads, hd).permute(0, 2, 3, 1)
    pr = torch.softmax(sc, -1) * keep2
    ref = (pr @ vf.view(N, L, heads, hd).permute(0, 2, 1, 3)).permute(0, 2, 1, 3).reshape(N, H)
    assert nerr(out, ref) < 1e-2
    dout = rnd(N, H, seed=24)
    (ref * dout.float()).sum().backward()
    dq2 = torch.empty(N, H, dtype=torch.bfloat16, device=DEV)
    dk2, dv2 = torch.empty(N, L, H, dtype=torch.bfloat16, device=DEV), torch.empty(N, L, H, dtype=torch.bfloat16, device=DEV)
    ops.cross_attn_q1_bwd(q2, k2, v2, out, lse, dout, dq2, dk2, dv2, N, L, H, heads, drop=(seed, site + 1, 0.1))
    assert nerr(dq2, qf.grad) < 2e-2 and nerr(dk2, kf.grad) < 2e-2 and nerr(dv2, vf.grad) < 2e-2


def test_fused_adamw_with_clipping_matches_reference_update():
    """Two launches (gradient norm, AdamW) against the per-tensor arithmetic of the reference driver:
    clip_grad_norm_ + transformers.AdamW (eps inside the bias-corrected denominator), two LR groups, a tensor without
    a gradient left untouched, odd-sized tensors, three steps."""
    from modcr_b200.training import FusedAdamW
    g = torch.Generator().manual_seed(0)
    shapes = [(768, 768), (3072,), (13, 7), (2, 768), (5,)]
    ps = [torch.randn(*sh, generator=g).to(DEV).requires_grad_(True) for sh in shapes]
    ref = [p.detach().clone().double() for p in ps]
    lr, b1, b2, eps, max_norm = 2e-3, 0.9, 0.999, 1e-5, 1.0
    groups = [{"params": ps[:3], "lr": lr}, {"params": ps[3:], "lr": lr * 0.1}]
    lrs = [lr, lr, lr, lr * 0.1, lr * 0.1]
    opt = FusedAdamW(groups, lr=lr, betas=(b1, b2), eps=eps, max_grad_norm=max_norm)
    m = [torch.zeros_like(r) for r in ref]
    v = [torch.zeros_like(r) for r in ref]
    for t in range(1, 4):
        grads = [torch.randn(*sh, generator=g).to(DEV) * (3.0 if t == 2 else 0.01) for sh in shapes]
        for i, p in enumerate(ps):
            p.grad = None if (i == 4 and t == 2) else grads[i].clone()
        versions = [p._version for p in ps]
        norm = opt.step()
        present = [grads[i].double() for i in range(5) if ps[i].grad is not None]
        tot = torch.sqrt(sum((x * x).sum() for x in present))
        assert abs(float(norm) - float(tot)) < 1e-3 * float(tot)
        clip = min(1.0, max_norm / (float(tot) + 1e-6))
        for i in range(5):
            if ps[i].grad is None:
                continue
            gi = grads[i].double() * clip
            m[i] = b1 * m[i] + (1 - b1) * gi
            v[i] = b2 * v[i] + (1 - b2) * gi * gi
            step = lrs[i] * (1 - b2 ** t) ** 0.5 / (1 - b1 ** t)
            ref[i] = ref[i] - step * m[i] / (v[i].sqrt() + eps)
        for i in range(5):
            assert float((ps[i].detach().double() - ref[i]).abs().max()) < 2e-6, (t, i)
            assert torch.equal(ps[i].grad, grads[i]) if ps[i].grad is not None else True    # .grad itself is not modified
        assert all(p._version == vv + 1 for p, vv in zip(ps, versions))     # operand caches see the update


def test_gemm_epilogue_dropout_and_layernorm_masked_gradient():
    """Dense-output dropout fused into the GEMM epilogue (C = dropout(acc + bias) + residual) and its backward half
    fused into the LayerNorm backward (second output dx * mask, bias gradient from the masked values), against the
    host-restated mask."""
    import numpy as np
    M, H, K, p, site, seedv = 1424, 768, 768, 0.3, 31, 424242
    seed = _seed_tensor(seedv)
    A, W = rnd(M, K, seed=51), rnd(H, K, scale=0.05, seed=52)
    bias = torch.randn(H, generator=torch.Generator().manual_seed(53)).to(DEV)
    res = torch.randn(M, H, generator=torch.Generator().manual_seed(54)).to(DEV)
    pre = torch.empty(M, H, device=DEV)
    ops.gemm(A, W, pre, M, H, K, bias=bias, epi_in=res, in_op=1, drop=(seed, site, p))
    keep = host_keep(seedv, site, np.arange(M * H), p).view(M, H).to(DEV).float() / (1 - p)
    ref = (A.float() @ W.float().t() + bias) * keep + res
    assert nerr(pre, ref) < 1e-4
    # LayerNorm backward with the masked second output
    gamma = (1 + 0.1 * torch.randn(H, generator=torch.Generator().manual_seed(55))).to(DEV)
    x = pre.clone().requires_grad_(True)
    y = torch.nn.functional.layer_norm(x, (H,), gamma, torch.zeros(H, device=DEV), 1e-12)
    dy = rnd(M, H, seed=56)
    (y * dy.float()).sum().backward()
    mean = pre.mean(-1)
    rstd = (pre.var(-1, unbiased=False) + 1e-12).rsqrt()
    dx = torch.empty(M, H, dtype=torch.bfloat16, device=DEV)
    dxm = torch.empty(M, H, dtype=torch.bfloat16, device=DEV)
    dgam, dbet, dbias = torch.zeros(H, device=DEV), torch.zeros(H, device=DEV), torch.zeros(H, device=DEV)
    ops.layernorm_bwd(dy, pre, mean, rstd, gamma, dx, dgam, dbet, dbias, M, H, dx_masked=dxm, drop=(seed, site, p))
    assert nerr(dx, x.grad) < 1e-2
    assert nerr(dxm, x.grad * keep) < 1e-2
    assert nerr(dbias, (x.grad * keep).sum(0)) < 1e-2


def test_multi_copy_gathers_and_scatters_gradient_buckets():
    """modcr_multi_copy (csrc/optim.cu): fp32 <-> bf16, scale, odd lengths and unaligned slots, > 64 tensors."""
    g = torch.Generator().manual_seed(9)
    sizes = [768 * 768, 3072, 7, 1, 30567 * 8 + 3, 2, 8191, 8193] + [5 + i for i in range(70)]
    srcs = [torch.randn(n, generator=g).to(DEV) for n in sizes]
    flat = torch.full((sum(sizes) + 16,), float("nan"), dtype=torch.bfloat16, device=DEV)
    views, off = [], 1                                   # deliberately unaligned first slot
    for n in sizes:
        views.append(flat[off:off + n])
        off += n
    ops.multi_copy(srcs, views, 0.5)
    for s_, v in zip(srcs, views):
        assert torch.equal(v, (s_ * 0.5).to(torch.bfloat16))
    back = [torch.empty_like(s_) for s_ in srcs]
    ops.multi_copy(views, back, 2.0)
    for s_, b_ in zip(srcs, back):
        assert torch.equal(b_, (s_ * 0.5).to(torch.bfloat16).float() * 2.0)
    f32 = [torch.empty_like(s_) for s_ in srcs[:5]]
    ops.multi_copy(srcs[:5], f32)
    assert all(torch.equal(a, b) for a, b in zip(srcs[:5], f32))
    assert bool(torch.isnan(flat[0].float())) and bool(torch.isnan(flat[off:].float()).all())   # nothing outside the slots
