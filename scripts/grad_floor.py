"""Developer experiment (CPU, test infrastructure only): relative L2 error of the GRADIENTS of the fp32 oracle when only
the tensor-core operands are rounded to bf16 in the forward pass (autograd through the rounding), against the reference
golden gradients -- the gradient error floor of a bf16-operand implementation.  Usage: python scripts/grad_floor.py cfg5_large"""
import os, sys, json
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


def linear(sd_, name, x):
    y = F.linear(r(x), r(sd_[name + ".weight"]), sd_.get(name + ".bias"))
    return r(y) if any(s in name for s in ("query", "key", "value")) else y


def l2err(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


out_path = os.path.join(GOLD, "bf16_grad_floor.json")
res = json.load(open(out_path)) if os.path.exists(out_path) else {}
for case in sys.argv[1:]:
    g = torch.load(os.path.join(GOLD, case + ".pt")); m = g["meta"]
    cfg = restate.default_cfg(**m.get("cfg_over", {}))
    b = synth.make_batch(B=m["B"], T=m["T"], R=m["R"], seed=m["seed"], region_pad=m["region_pad"])
    prefix = synth.make_prefix(cfg["num_hidden_layers"], m["P"], cfg["hidden_size"], seed=m["seed"] + 1).requires_grad_(True)
    sd = params.fill_params(statedict.materialise(statedict.chunkalign_keys(cfg)), m["wseed"])
    names = sorted({k.split("[")[0] for k in g["grads"] if not k.startswith("prefix")})
    for k in names:
        sd[k].requires_grad_(True)
    restate.linear = linear
    try:
        (loss_cls, _, align_loss, _, _), ex = restate.chunkalign_forward(sd, cfg, b, prefix)
        (loss_cls + align_loss).backward()
    finally:
        restate.linear = orig
    errs = {}
    for k, ref in g["grads"].items():
        if k.startswith("prefix"):
            got = prefix.grad[:, :, ::4]
        elif "[" in k:
            base, sl = k.split("[", 1)
            gr = sd[base].grad
            if sl.startswith("ids0"):
                got = gr[b["input_ids"][0, :8], :64]
            else:
                rr, cc = sl.rstrip("]").split(",")
                got = gr[:int(rr.strip(":")), :int(cc.strip(":"))]
        else:
            got = sd[k].grad
        if k.endswith("attention.self.key.bias"):
            continue
        errs[k] = l2err(got, ref)
    worst = sorted(errs.items(), key=lambda kv: -kv[1])
    print(case, "dlogits %.4f" % float((ex["logits"] - g["logits"]).abs().max()), "grad rel-L2 floor, worst:", [(k, round(v, 4)) for k, v in worst[:8]], flush=True)
    res[case] = dict(max=worst[0][1], per_key=errs)
    json.dump(res, open(out_path, "w"), indent=1, sort_keys=True)
