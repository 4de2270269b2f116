import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import modcr_b200
from modcr_b200 import synth, ops, engine
from modcr_b200.modeling_bert import BertImgModel
from modcr_b200.modeling_utils import BertConfig
from modcr_b200.modeling_vcr_chunkalign_v10 import ChunkAlign_CLS_enc4_align, SeqBertImgModel
dev = "cuda"
def try_capture(name, fn, warm=2):
    try:
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warm): fn()
        torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = fn()
        g.replay(); torch.cuda.synchronize()
        print("CAPTURE OK  ", name, flush=True)
        return g, out
    except Exception as e:
        print("CAPTURE FAIL", name, str(e)[:300].replace("\n", " "), flush=True)
        torch.cuda.synchronize()
        return None, None
A = torch.randn(512, 768, device=dev).bfloat16(); Bm = torch.randn(768, 768, device=dev).bfloat16(); C = torch.empty(512, 768, device=dev, dtype=torch.bfloat16)
try_capture("gemm", lambda: ops.gemm(A, Bm, C, 512, 768, 768))
x = torch.randn(512, 768, device=dev); y = torch.empty(512, 768, device=dev, dtype=torch.bfloat16); g_ = torch.ones(768, device=dev); b_ = torch.zeros(768, device=dev)
mean = torch.empty(512, device=dev); rstd = torch.empty(512, device=dev)
try_capture("layernorm", lambda: ops.layernorm_fwd(x, g_, b_, y, mean, rstd, 512, 768, 1e-12))
B, T, R, P = 4, 128, 50, 10
cfg = BertConfig(vocab_size=30567, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
m = ChunkAlign_CLS_enc4_align(BertImgModel(cfg), SeqBertImgModel(cfg), 4).to(dev); m.train()
b = synth.make_batch(B=B, T=T, R=R, seed=88)
cid, _ = engine.pack_chunk_ids(b["gather_index"], T, "cpu"); b["chunk_id"] = cid
bd = {k: v.to(dev) for k, v in b.items() if torch.is_tensor(v)}
N = B * 4
prefix = (torch.randn(12, P, 768, device=dev) * 0.02).requires_grad_(True)
def genc_nograd():
    with torch.no_grad():
        return m.global_enc(bd["input_ids"], token_type_ids=bd["token_type_ids"], attention_mask=bd["input_mask"], img_feats=bd["img_feat"])[0]
try_capture("global_enc no_grad", genc_nograd)
def senc_nograd():
    with torch.no_grad():
        return m.seq_enc(bd["input_ids"], token_type_ids=bd["token_type_ids"], attention_mask=bd["chunk_attention_mask"], input_mask=bd["input_mask"], img_feats=bd["img_feat"], gather_index=bd["chunk_id"])[0][0]
try_capture("seq_enc no_grad", senc_nograd)
def fwd():
    hist = [prefix[i].unsqueeze(0).expand(N, -1, -1) for i in range(12)]
    out = m(bd["input_ids"], bd["img_feat"], input_mask=bd["input_mask"], label=bd["label"], token_type_ids=bd["token_type_ids"],
            encoder_history_states=hist, chunk_attention_mask=bd["chunk_attention_mask"], gather_index=bd["chunk_id"],
            align_pos=bd["align_pos"], total_label=bd["total_label"])
    return out[0] + out[2]
def fwd_nograd():
    with torch.no_grad():
        return fwd()
try_capture("full fwd no_grad", fwd_nograd)
try_capture("full fwd (grad graph built, no backward)", fwd)
def genc_bwd():
    o = m.global_enc(bd["input_ids"], token_type_ids=bd["token_type_ids"], attention_mask=bd["input_mask"], img_feats=bd["img_feat"])
    o[0].float().sum().backward()
    return o[0]
try_capture("global_enc fwd+bwd", genc_bwd)
def fb():
    l = fwd(); l.backward(); return l
try_capture("full fwd+bwd", fb)
