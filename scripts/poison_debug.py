"""Developer check: poison the caching allocator with NaN bytes, rerun the forward and report which outputs change
(finds reads of uninitialised device memory)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import modcr_b200
from modcr_b200 import synth
from modcr_b200.modeling_bert import BertImgModel
from modcr_b200.modeling_utils import BertConfig
from modcr_b200.modeling_vcr_chunkalign_v10 import ChunkAlign_CLS_enc4_align, SeqBertImgModel

dev = "cuda"
B, T, R, P = int(os.environ.get("B", 32)), 128, 50, 10
N = 4 * B
torch.manual_seed(0)
cfg = BertConfig(vocab_size=30567, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
m = ChunkAlign_CLS_enc4_align(BertImgModel(cfg), SeqBertImgModel(cfg), 4).to(dev).train()
b = synth.make_batch(B=B, T=T, R=R, seed=88)
bd = synth.to_device(b, dev)
prefix = synth.make_prefix(12, P, 768, seed=89).to(dev).requires_grad_(True)


def poison(val):
    free, total = torch.cuda.mem_get_info()
    n = int(min(free * 0.6, 60e9)) // 4
    x = torch.empty(n, dtype=torch.float32, device=dev)
    x.fill_(val)
    # the small pool (< 1 MB requests) has its own segments: poison those too
    small = [torch.full((sz // 4,), val, dtype=torch.float32, device=dev) for sz in (512, 4096, 65536, 524288, 1 << 20)
             for _ in range(256)]
    del x, small   # stays in the caching allocator: later torch.empty() calls hand these bytes out


def run(backward):
    hist = [prefix[i].unsqueeze(0).expand(N, -1, -1) for i in range(12)]
    out = m(bd["input_ids"], bd["img_feat"], input_mask=bd["input_mask"], label=bd["label"],
            token_type_ids=bd["token_type_ids"], encoder_history_states=hist, offsets=b["offsets"],
            chunk_attention_mask=bd["chunk_attention_mask"], gather_index=b["gather_index"],
            align_pos=bd["align_pos"], total_label=bd["total_label"])
    res = dict(logits=m.last_logits.clone(), loss=out[0].detach().clone(), align=out[2].detach().clone(),
               block=m.last_align_block.clone(), **{k: v.clone() for k, v in m.last_intermediates.items()})
    if backward:
        (out[0] + out[2]).backward()
        res["grads"] = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        res["grads"]["prefix"] = prefix.grad.clone()
        for p in m.parameters():
            p.grad = None
        prefix.grad = None
    return res


live = bd["input_mask"][:, :, None]
for val, name in ((0.0, "zeros"), (float("nan"), "nan"), (3.0e4, "3e4"), (float("nan"), "nan-again")):
    torch.cuda.empty_cache()
    poison(val)
    r = run(True)
    torch.cuda.synchronize()
    if name == "zeros":
        ref = r
        continue
    print("== poison", name)
    for k in r:
        if k == "grads":
            bad = [(g, float((r["grads"][g] - ref["grads"][g]).abs().max())) for g in r["grads"]
                   if not torch.equal(r["grads"][g], ref["grads"][g])]
            nan = [g for g in r["grads"] if not bool(torch.isfinite(r["grads"][g]).all())]
            print("  grads differing: %d of %d; non-finite: %d; worst: %s" % (
                len(bad), len(r["grads"]), len(nan), sorted(bad, key=lambda kv: -kv[1] if kv[1] == kv[1] else -1e30)[:4]))
            continue
        a, c = r[k].float(), ref[k].float()
        if a.dim() == 3 and a.shape[1] == T + R:
            a, c = a * live, c * live
            a, c = torch.where(live.bool().expand_as(a), a, torch.zeros_like(a)), torch.where(live.bool().expand_as(c), c, torch.zeros_like(c))
        same = torch.equal(a, c)
        print("  %-22s equal=%s finite=%s maxdiff=%g" % (k, same, bool(torch.isfinite(a).all()), float((a - c).abs().nan_to_num(1e30).max())))
