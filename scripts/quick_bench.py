"""Developer timing loop (not the contract bench): fwd+bwd of ChunkAlign_CLS_enc4_align at cfg-2."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import modcr_b200
from modcr_b200 import synth
from modcr_b200.modeling_bert import BertImgModel
from modcr_b200.modeling_utils import BertConfig
from modcr_b200.modeling_vcr_chunkalign_v10 import ChunkAlign_CLS_enc4_align, SeqBertImgModel

B = int(os.environ.get("B", 32)); T = int(os.environ.get("T", 128)); R = int(os.environ.get("R", 50)); P = int(os.environ.get("P", 10))
steps = int(os.environ.get("STEPS", 5))
dev = "cuda"
torch.manual_seed(0)
cfg = BertConfig(vocab_size=30567, hidden_dropout_prob=float(os.environ.get("P_HIDDEN", 0.0)),
                 attention_probs_dropout_prob=float(os.environ.get("P_ATTN", 0.0)))
m = ChunkAlign_CLS_enc4_align(BertImgModel(cfg), SeqBertImgModel(cfg), 4).to(dev)
m.train()
b = synth.make_batch(B=B, T=T, R=R, seed=88)
bd = synth.to_device(b, dev)
N = B * 4
prefix = (torch.randn(12, P, 768, device=dev) * 0.02).requires_grad_(True)
def step():
    hist = [prefix[i].unsqueeze(0).expand(N, -1, -1) for i in range(12)] if P else None
    out = m(bd["input_ids"], bd["img_feat"], input_mask=bd["input_mask"], label=bd["label"], token_type_ids=bd["token_type_ids"],
            encoder_history_states=hist, offsets=b["offsets"], chunk_attention_mask=bd["chunk_attention_mask"],
            gather_index=b["gather_index"], align_pos=bd["align_pos"], total_label=bd["total_label"])
    loss = out[0] + out[2]
    loss.backward()
    for p in m.parameters(): p.grad = None
    prefix.grad = None
    return loss
for _ in range(int(os.environ.get("WARM", 3))): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(steps): l = step()
e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
ms = e0.elapsed_time(e1) / steps
print("B=%d T=%d R=%d P=%d  %.2f ms/step (wall %.2f)  %.1f samples/s  loss %.4f  mem %.1f GB" % (
    B, T, R, P, ms, (t1 - t0) * 1e3 / steps, B / ms * 1e3, float(l), torch.cuda.max_memory_allocated() / 2**30))
if os.environ.get("PROFILE"):
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step(); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=60))
