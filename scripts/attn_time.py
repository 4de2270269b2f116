"""Developer timing of the prefix-attention kernels alone at the cfg-2 shape (CUDA events, L2 flushed), modes 0/1/2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import modcr_b200
from modcr_b200 import ops

dev = "cuda"
N, T, R, heads = 128, int(os.environ.get("T", 128)), int(os.environ.get("R", 50)), 12
S, H = T + R, heads * 64
BF16 = torch.bfloat16
flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream()


def timed(fn, reps=8):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); fn(); b.record(st); b.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps * 1e3


g = torch.Generator().manual_seed(0)
for mode, P, variant in ((0, 10, ""), (0, 0, ""), (1, 0, ""), (2, 0, "dense"), (2, 0, "sparse"), (2, 0, "noprobs")):
    qkv = (torch.randn(N, S, 3 * H, generator=g) * 0.5).to(dev, BF16)
    pkv = (torch.randn(N, P, 2 * H, generator=g) * 0.5).to(dev, BF16) if P else None
    kbias = torch.zeros(N, P + S, device=dev)
    cw = (T + 31) // 32
    cbits = torch.randint(0, 2 ** 31 - 1, (N, T, cw), dtype=torch.int32, device=dev) if mode else None
    ctx, lse = torch.empty(N, S, H, dtype=BF16, device=dev), torch.empty(N, heads, S, device=dev)
    dctx = (torch.randn(N, S, H, generator=g) * 0.5).to(dev, BF16)
    dqkv = torch.empty(N, S, 3 * H, dtype=BF16, device=dev)
    dpkv = torch.empty(N, P, 2 * H, dtype=BF16, device=dev) if P else None
    ws = torch.empty(max(ops.attn_bwd_workspace_bytes(N, S, heads) // 4, 1), device=dev)
    probs = torch.empty(N, heads, T, R, device=dev) if (mode == 2 and variant != "noprobs") else None
    sel = None
    if variant == "sparse":      # ~4 aligned rows per sequence, like the synthetic batches
        sel = (torch.rand(N, T, generator=g) < 4.0 / T).to(torch.int64).to(dev)
    dalign = torch.randn(N, T, R, device=dev) * 0.1 if probs is not None else None
    if sel is not None:
        dalign = dalign * sel.unsqueeze(-1).float()
    Tm = T if mode else 0
    tf = timed(lambda: ops.prefix_attn_fwd(N, S, P, heads, Tm, mode, qkv, pkv, kbias, cbits, cw if mode else 0, ctx, lse, probs,
                                           row_sel=sel))
    tb = timed(lambda: ops.prefix_attn_bwd(N, S, P, heads, Tm, mode, qkv, pkv, kbias, cbits, cw if mode else 0, ctx, lse, probs,
                                           dctx, dalign, dqkv, dpkv, ws, row_sel=sel))
    print("mode %d P=%2d S=%d %-7s: fwd %.1f us  bwd %.1f us" % (mode, P, S, variant, tf, tb), flush=True)
