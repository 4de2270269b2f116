"""Developer check: run each hot forward op many times on fixed inputs and compare outputs bitwise (cfg-2 shapes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import modcr_b200
from modcr_b200 import ops

dev = "cuda"
N, T, R, P, H, heads, I = 128, 128, 50, 10, 768, 12, 3072
S = T + R
M = N * S
REPS = int(os.environ.get("REPS", 300))
g = torch.Generator().manual_seed(0)
rnd = lambda *s, sc=1.0, dt=torch.bfloat16: (torch.randn(*s, generator=g) * sc).to(dt).to(dev)  # noqa: E731


def check(name, fn, outs):
    fn()
    torch.cuda.synchronize()
    ref = [o.clone() for o in outs]
    bad, first = 0, None
    for i in range(REPS):
        for o in outs:
            o.fill_(7.0) if o.dtype != torch.int32 else None
        fn()
        for k, (o, r) in enumerate(zip(outs, ref)):
            if not torch.equal(o, r):
                bad += 1
                if first is None:
                    d = (o.float() - r.float()).abs()
                    idx = torch.nonzero(d > 0)
                    first = (i, k, int(idx.shape[0]), idx[:4].tolist(), idx[-2:].tolist(), float(d.max()))
                break
    print("%-34s mismatching runs %d / %d  first: %s" % (name, bad, REPS, first), flush=True)


x = rnd(M, H)
x32 = x.float()
Wqkv, bqkv = rnd(3 * H, H, sc=0.03), rnd(3 * H, sc=0.1, dt=torch.float32)
qkv = torch.empty(M, 3 * H, dtype=torch.bfloat16, device=dev)
check("gemm qkv (bf16 out, bias)", lambda: ops.gemm(x, Wqkv, qkv, M, 3 * H, H, bias=bqkv), [qkv])
hist = rnd(N * P, H, sc=0.5)
pkv = torch.empty(N * P, 2 * H, dtype=torch.bfloat16, device=dev)
check("gemm prefix kv (M=1280)", lambda: ops.gemm(hist, Wqkv[H:], pkv, N * P, 2 * H, H, bias=bqkv[H:]), [pkv])
lens = torch.randint(T // 2, T + 1, (N,), generator=g)
mask = torch.cat([(torch.arange(T)[None] < lens[:, None]).float(), torch.ones(N, R)], 1)
kb = ((1.0 - torch.cat([torch.ones(N, P), mask], 1)) * -10000.0).to(dev).contiguous()
ctx = torch.empty(M, H, dtype=torch.bfloat16, device=dev)
lse = torch.empty(N, heads, S, dtype=torch.float32, device=dev)
check("prefix_attn_fwd mode 0 P=10", lambda: ops.prefix_attn_fwd(N, S, P, heads, 0, 0, qkv, pkv, kb, None, 0, ctx, lse),
      [ctx, lse])
kb0 = ((1.0 - mask) * -10000.0).to(dev).contiguous()
check("prefix_attn_fwd mode 0 P=0", lambda: ops.prefix_attn_fwd(N, S, 0, heads, 0, 0, qkv, None, kb0, None, 0, ctx, lse),
      [ctx, lse])
Wo, bo = rnd(H, H, sc=0.03), rnd(H, sc=0.1, dt=torch.float32)
pre = torch.empty(M, H, dtype=torch.float32, device=dev)
check("gemm attn-out (fp32 out + fp32 res)", lambda: ops.gemm(ctx, Wo, pre, M, H, H, bias=bo, epi_in=x32, in_op=1), [pre])
gam, bet = rnd(H, dt=torch.float32), rnd(H, dt=torch.float32)
a, a32 = torch.empty(M, H, dtype=torch.bfloat16, device=dev), torch.empty(M, H, dtype=torch.float32, device=dev)
mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
check("layernorm_fwd (fp32 in)", lambda: ops.layernorm_fwd(pre, gam, bet, a, mean, rstd, M, H, 1e-12, y32=a32),
      [a, a32, mean, rstd])
W1, b1 = rnd(I, H, sc=0.03), rnd(I, sc=0.1, dt=torch.float32)
h = torch.empty(M, I, dtype=torch.bfloat16, device=dev)
aux = torch.empty(M, I, dtype=torch.bfloat16, device=dev)
check("gemm ffn1 (+GELU, aux)", lambda: ops.gemm(a, W1, h, M, I, H, bias=b1, act=1, aux=aux), [h, aux])
W2, b2 = rnd(H, I, sc=0.02), rnd(H, sc=0.1, dt=torch.float32)
check("gemm ffn2 (fp32 out + fp32 res)", lambda: ops.gemm(h, W2, pre, M, H, I, bias=b2, epi_in=a32, in_op=1), [pre])

# ---- backward ops whose outputs are written once (no atomics): must be bit-identical as well
ops.prefix_attn_fwd(N, S, P, heads, 0, 0, qkv, pkv, kb, None, 0, ctx, lse)
dctx = rnd(M, H, sc=0.1)
dqkv = torch.empty(M, 3 * H, dtype=torch.bfloat16, device=dev)
dpkv = torch.empty(N * P, 2 * H, dtype=torch.bfloat16, device=dev)
ws = torch.empty(max(ops.attn_bwd_workspace_bytes(N, S, heads) // 4, 1), dtype=torch.float32, device=dev)
check("prefix_attn_bwd mode 0 P=10", lambda: ops.prefix_attn_bwd(N, S, P, heads, 0, 0, qkv, pkv, kb, None, 0, ctx, lse, None,
                                                                 dctx, None, dqkv, dpkv, ws), [dqkv, dpkv])
dh = torch.empty(M, I, dtype=torch.bfloat16, device=dev)
dpre = rnd(M, H, sc=0.1)
check("gemm dgrad ffn2 (x GELU')", lambda: ops.gemm(dpre, W2, dh, M, I, H, b_mn=1, epi_in=aux, in_op=2), [dh])
dx = torch.empty(M, H, dtype=torch.bfloat16, device=dev)
dres = rnd(M, H, sc=0.1)
check("gemm dgrad ffn1 (+bf16 residual)", lambda: ops.gemm(dh, W1, dx, M, H, I, b_mn=1, epi_in=dres, in_op=1), [dx])
dW = torch.empty(I, H, dtype=torch.float32, device=dev)
check("gemm wgrad ffn1 (no split-K)", lambda: ops.gemm(dh, a, dW, I, H, M, a_mn=1, b_mn=1), [dW])
