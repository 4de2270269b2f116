"""Developer experiment: clock64() timeline of attn_bwd_kernel at the cfg-2 shape (needs `make trace`).
    MODCR_B200_LIB=multimodal-context-reasoning_b200/libmodcr_b200_trace.so python scripts/attn_bwd_trace.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("MODCR_B200_LIB", os.path.join(ROOT, "multimodal-context-reasoning_b200", "libmodcr_b200_trace.so"))
sys.path.insert(0, ROOT)
import torch
import modcr_b200
from modcr_b200 import _lib, ops

dev = "cuda"
N, S, P, heads = int(os.environ.get("N", 128)), int(os.environ.get("S", 178)), int(os.environ.get("P", 10)), 12
H = heads * 64
BF16 = torch.bfloat16
g = torch.Generator().manual_seed(0)
qkv = (torch.randn(N, S, 3 * H, generator=g) * 0.5).to(dev, BF16)
pkv = (torch.randn(N, P, 2 * H, generator=g) * 0.5).to(dev, BF16) if P else None
kbias = torch.zeros(N, P + S, device=dev)
ctx, lse = torch.empty(N, S, H, dtype=BF16, device=dev), torch.empty(N, heads, S, device=dev)
dctx = (torch.randn(N, S, H, generator=g) * 0.5).to(dev, BF16)
dqkv = torch.empty(N, S, 3 * H, dtype=BF16, device=dev)
dpkv = torch.empty(N, P, 2 * H, dtype=BF16, device=dev) if P else None
ws = torch.empty(max(ops.attn_bwd_workspace_bytes(N, S, heads) // 4, 1), device=dev)
ops.prefix_attn_fwd(N, S, P, heads, 0, 0, qkv, pkv, kbias, None, 0, ctx, lse, None)
for _ in range(3):
    ops.prefix_attn_bwd(N, S, P, heads, 0, 0, qkv, pkv, kbias, None, 0, ctx, lse, None, dctx, None, dqkv, dpkv, ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.prefix_attn_bwd(N, S, P, heads, 0, 0, qkv, pkv, kbias, None, 0, ctx, lse, None, dctx, None, dqkv, dpkv, ws)
e1.record()
torch.cuda.synchronize()
print("attn_bwd N=%d S=%d P=%d: %.1f us (traced build)" % (N, S, P, e0.elapsed_time(e1) * 1e3))
lib = _lib.load()
CT, SL = 32, 128
buf = (ctypes.c_longlong * (CT * SL))()
lib.modcr_debug_read_attn_trace.argtypes = [ctypes.c_void_p, ctypes.c_int32]
n = lib.modcr_debug_read_attn_trace(buf, CT * SL)
assert n == CT * SL, n
names = {0: "R entry", 1: "R delta done", 2: "R after sync", 60: "R final bar_free", 61: "R final read_out", 62: "R after end sync",
         64: "M entry", 65: "M after sync", 66: "M K/V landed", 67: "M S/dP(0) issued"}
for st in range(4):
    b = 3 + st * 12
    names.update({b: "R s%d c0 S/dP ready" % st, b + 1: "R s%d c0 loaded" % st, b + 2: "R s%d c0 math done" % st,
                  b + 3: "R s%d bar_free(prev) ok" % st, b + 4: "R s%d arrive pds" % st, b + 5: "R s%d read_out done" % st,
                  b + 6: "R s%d c1 S/dP ready" % st, b + 7: "R s%d c1 loaded" % st, b + 8: "R s%d c1 math done" % st})
    m = 68 + st * 6
    names.update({m: "M s%d cons0 + S/dP(next,0) issued" % st, m + 1: "M s%d cons1 + S/dP(next,1) issued" % st,
                  m + 2: "M s%d P/dS ready" % st, m + 3: "M s%d dV/dK/dQ issued" % st})
for w in range(10):
    names[100 + w] = "warp %d reaches the prologue barrier" % w
for kb in range(2):
    names[110 + kb * 3] = "R read_out dK/dV(kb%d) start" % kb
    names[111 + kb * 3] = "R read_out dK/dV(kb%d) tmem read, arrived dkv_free" % kb
for st in range(4):
    names[116 + st * 3] = "OBS S/dP(s%d,0) complete" % st
    names[117 + st * 3] = "OBS S/dP(s%d,1) complete" % st
    names[118 + st * 3] = "OBS dV/dK/dQ(s%d) complete (bar_free)" % st
for cta in (16, 21):
    t = [buf[cta * SL + i] for i in range(SL)]
    t0 = t[0]
    ev = sorted((t[i] - t0, names.get(i, "slot %d" % i)) for i in range(SL) if t[i] != 0)
    print("--- CTA %d (cycles since row-thread entry)" % cta)
    prev = 0
    for c, nm in ev:
        print("  %8d  (+%6d)  %s" % (c, c - prev, nm))
        prev = c
