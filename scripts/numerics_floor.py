"""Developer experiment (CPU, test infrastructure only): where does the bf16-operand error of the logits come from?
Runs the fp32 oracle with GEMM operands (and q/k/v outputs) rounded to bf16 in selected parts of the model and reports
max |dlogits| against the reference golden.  Usage: python scripts/numerics_floor.py [cfg5_large|fwd256:<batch>] ..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, torch.nn.functional as F
import params, restate, statedict
import modcr_b200
from modcr_b200 import synth

torch.set_num_threads(os.cpu_count())
GOLD = os.path.join(ROOT, "tests", "golden")
r = lambda t: t.to(torch.bfloat16).float()
orig = restate.linear


def make_linear(pred, round_out_all=False):
    def linear(sd_, name, x):
        if not pred(name):
            return orig(sd_, name, x)
        y = F.linear(r(x), r(sd_[name + ".weight"]), sd_.get(name + ".bias"))
        return r(y) if any(s in name for s in ("query", "key", "value")) else y
    return linear


def run(case, variants_filter=None):
    if case.startswith("fwd256"):
        g = torch.load(os.path.join(GOLD, "cfg2_fwd256.pt")); m = g["meta"]; i = int(case.split(":")[1])
        cfg = restate.default_cfg()
        b = synth.make_batch(B=m["B"], T=m["T"], R=m["R"], seed=m["seed0"] + i)
        prefix = synth.make_prefix(12, m["P"], 768, seed=m["seed0"] - 1)
        ref = g["logits"][i]
        wseed = m["wseed"]
    else:
        g = torch.load(os.path.join(GOLD, case + ".pt")); m = g["meta"]
        cfg = restate.default_cfg(**m.get("cfg_over", {}))
        b = synth.make_batch(B=m["B"], T=m["T"], R=m["R"], seed=m["seed"], region_pad=m["region_pad"])
        prefix = synth.make_prefix(cfg["num_hidden_layers"], m["P"], cfg["hidden_size"], seed=m["seed"] + 1)
        ref = g["logits"]; wseed = m["wseed"]
    sd = params.fill_params(statedict.materialise(statedict.chunkalign_keys(cfg)), wseed)
    enc = lambda n: n.startswith("global_enc.") or n.startswith("seq_enc.")
    variants = {
        "all": lambda n: True, "encoders only": enc, "head only": lambda n: not enc(n),
        "head: cls_q/align_k only": lambda n: ("cls_q_proj" in n or "align_k_proj" in n),
        "head: all but cls_q/align_k": lambda n: (not enc(n)) and not ("cls_q_proj" in n or "align_k_proj" in n),
    }
    out = {}
    for name, pred in variants.items():
        if variants_filter is not None and name not in variants_filter:
            continue
        restate.linear = make_linear(pred)
        try:
            with torch.no_grad():
                _, ex = restate.chunkalign_forward(sd, cfg, b, prefix)
        finally:
            restate.linear = orig
        d = (ex["logits"] - ref).abs()
        out[name] = dict(max=float(d.max()), frac_within_2e2=float((d.max(dim=-1).values < 2e-2).float().mean()))
        print("%-14s %-30s max|dlogits| %.4f  (|logits| max %.2f, |CLS| max %.2f)" % (
            case, name, float(d.max()), float(ref.abs().max()), float(ex["CLS"].abs().max())), flush=True)
    return out


if len(sys.argv) > 1 and sys.argv[1] == "--fixture":
    # tests/golden/bf16_floor.json: max |dlogits| of the fp32 ORACLE when nothing but the tensor-core operands (GEMM inputs,
    # weights, stored Q/K/V) are rounded to bf16 -- the error any bf16-compute implementation of the case carries
    import json
    res = {}
    for c in ["ragged_small", "cfg1_prefix10", "cfg1_noprefix", "cfg3_vcr", "cfg5_large"] + ["fwd256:%d" % i for i in range(32)]:
        res[c] = run(c, ("all",))["all"]
        json.dump(res, open(os.path.join(GOLD, "bf16_floor.json"), "w"), indent=1, sort_keys=True)
else:
    for c in (sys.argv[1:] or ["cfg5_large", "fwd256:5", "fwd256:12"]):
        run(c)
