"""Developer check: repeat the forward many times and compare every output bitwise against the first run."""
import gc, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import modcr_b200
from modcr_b200 import synth
from modcr_b200.modeling_bert import BertImgModel
from modcr_b200.modeling_utils import BertConfig
from modcr_b200.modeling_vcr_chunkalign_v10 import ChunkAlign_CLS_enc4_align, SeqBertImgModel

dev = "cuda"
B, T, R, P = int(os.environ.get("B", 32)), 128, 50, 10
REPS = int(os.environ.get("REPS", 200))
BWD = int(os.environ.get("BWD", 0))
N = 4 * B
torch.manual_seed(0)
cfg = BertConfig(vocab_size=30567, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
m = ChunkAlign_CLS_enc4_align(BertImgModel(cfg), SeqBertImgModel(cfg), 4).to(dev).train()
b = synth.make_batch(B=B, T=T, R=R, seed=88)
bd = synth.to_device(b, dev)
prefix = synth.make_prefix(12, P, 768, seed=89).to(dev).requires_grad_(True)


def run(i):
    hist = [prefix[i].unsqueeze(0).expand(N, -1, -1) for i in range(12)]
    out = m(bd["input_ids"], bd["img_feat"], input_mask=bd["input_mask"], label=bd["label"],
            token_type_ids=bd["token_type_ids"], encoder_history_states=hist, offsets=b["offsets"],
            chunk_attention_mask=bd["chunk_attention_mask"], gather_index=b["gather_index"],
            align_pos=bd["align_pos"], total_label=bd["total_label"])
    res = dict(logits=m.last_logits.clone(), loss=out[0].detach().clone(), align=out[2].detach().clone(),
               block=m.last_align_block.clone(), **{k: v.clone() for k, v in m.last_intermediates.items()})
    if BWD:
        (out[0] + out[2]).backward()
        for p in m.parameters():
            p.grad = None
        prefix.grad = None
    del out
    gc.collect()     # a forward without backward leaves its saved activations in reference cycles
    return res


ref = run(0)
torch.cuda.synchronize()
counts = {k: 0 for k in ref}
first = {}
for i in range(1, REPS):
    r = run(i)
    for k in r:
        if not torch.equal(r[k], ref[k]):
            counts[k] += 1
            if k not in first:
                d = (r[k].float() - ref[k].float()).abs()
                idx = torch.nonzero(d > 0)
                first[k] = (i, int(idx.shape[0]), idx[:6].tolist(), float(d.max()))
print("reps", REPS, "mismatch counts:", counts)
for k, v in first.items():
    print("  first mismatch of %s: run %d, %d elements, e.g. %s, max |d| %g" % (k, v[0], v[1], v[2], v[3]))
