"""TEST INFRASTRUCTURE ONLY -- stand-in for the language-model side of ``Abstract_Specific``.

The reference's RoBERTa-large with ``prompt_embeddings`` lives in ``local_transformers.zip``, which the reference
repository does not ship (parity unpinned).  The tests therefore drive ``Abstract_Specific`` -- the real
reference class in tests/test_oracle.py, the CUDA build in tests/test_model_gpu.py, the restatement in
oracle/restate.py -- with this small deterministic module that has the same call signature: the 10 prompt embeddings
are prepended to the token embeddings, masked-mean pooled and passed through a tanh pooler; element 1 of the result is
the pooled [N,1024] vector (modeling_ensemble.py:502-504)."""
import torch
from torch import nn


class StandInLM(nn.Module):
    def __init__(self, vocab=128, hidden=1024, seed=17):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.emb = nn.Embedding(vocab, hidden)
        self.dense = nn.Linear(hidden, hidden)
        with torch.no_grad():
            self.emb.weight.copy_(torch.randn(vocab, hidden, generator=g) * 0.5)
            self.dense.weight.copy_(torch.randn(hidden, hidden, generator=g) * 0.03)
            self.dense.bias.copy_(torch.randn(hidden, generator=g) * 0.1)

    def forward(self, input_ids=None, token_type_ids=None, attention_mask=None, prompt_embeddings=None,
                input_mask=None):
        x = torch.cat([prompt_embeddings.to(self.emb.weight.dtype), self.emb(input_ids)], dim=1)
        m = torch.cat([input_mask, attention_mask], dim=1).to(x.dtype).unsqueeze(-1)
        mean = (x * m).sum(1) / m.sum(1).clamp_min(1.0)
        return x, torch.tanh(self.dense(mean))


def make_inputs(N, Tr=12, vocab=128, seed=0, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, vocab, (N, Tr), generator=g)
    lens = torch.randint(Tr // 2, Tr + 1, (N,), generator=g)
    mask = (torch.arange(Tr).unsqueeze(0) < lens.unsqueeze(1)).to(torch.float32)
    return dict(r_input_ids=ids.to(device), r_token_type_ids=torch.zeros_like(ids).to(device),
                r_attention_mask=mask.to(device))
