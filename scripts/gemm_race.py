"""Developer check: hammer the fp32-output + fp32-residual GEMM (BertOutput.dense shape) and report every mismatch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import modcr_b200
from modcr_b200 import ops

dev = "cuda"
M, H, I = 128 * 178, 768, int(os.environ.get("K", 3072))
REPS = int(os.environ.get("REPS", 2000))
RES = int(os.environ.get("RES", 1))
F32OUT = int(os.environ.get("F32OUT", 1))
g = torch.Generator().manual_seed(0)
rnd = lambda *s, sc=1.0, dt=torch.bfloat16: (torch.randn(*s, generator=g) * sc).to(dt).to(dev)  # noqa: E731
h, W2, b2 = rnd(M, I), rnd(H, I, sc=0.02), rnd(H, sc=0.1, dt=torch.float32)
res = rnd(M, H, dt=torch.float32 if F32OUT else torch.bfloat16)
out = torch.empty(M, H, dtype=torch.float32 if F32OUT else torch.bfloat16, device=dev)


def fn():
    if RES:
        ops.gemm(h, W2, out, M, H, I, bias=b2, epi_in=res, in_op=1)
    else:
        ops.gemm(h, W2, out, M, H, I, bias=b2)


fn()
torch.cuda.synchronize()
ref = out.clone()
nbad = 0
for i in range(REPS):
    out.fill_(777.0)
    fn()
    if not torch.equal(out, ref):
        nbad += 1
        d = (out.float() - ref.float()).abs()
        idx = torch.nonzero(d > 0)
        if nbad <= 12:
            r0, c0 = idx[0].tolist()
            dd = (out[r0, c0:c0 + 4] - ref[r0, c0:c0 + 4]).tolist()
            hyp = {}
            for sh in (-96, -64, -32, -16, 16, 32, 64, 96):
                if 0 <= c0 + sh and c0 + sh + 4 <= H:
                    hyp["in%+d" % sh] = (res[r0, c0 + sh:c0 + sh + 4].float() - res[r0, c0:c0 + 4].float()).tolist()
                    hyp["out%+d" % sh] = (ref[r0, c0 + sh:c0 + sh + 4].float() - ref[r0, c0:c0 + 4].float()).tolist()
            best = sorted(hyp.items(), key=lambda kv: sum((a - b) ** 2 for a, b in zip(kv[1], dd)))[:2]
            print("   delta", [round(x, 4) for x in dd], "closest hypotheses:",
                  [(k, [round(x, 4) for x in v]) for k, v in best], flush=True)
            print("rep %d: %d elements; rows %s cols %s; got %s want %s; res %s" % (
                i, idx.shape[0], sorted(set(idx[:, 0].tolist()))[:6], sorted(set(idx[:, 1].tolist()))[:8],
                out[r0, c0:c0 + 4].tolist(), ref[r0, c0:c0 + 4].tolist(), res[r0, c0:c0 + 4].tolist()), flush=True)
print("K=%d RES=%d F32OUT=%d NCTA=%s: %d mismatching runs of %d" % (I, RES, F32OUT, os.environ.get("MODCR_GEMM_NCTA", "2"),
                                                                  nbad, REPS))
