"""Synthetic PMR / VCR-shaped batches.

Follows the batch contract of the reference collate function
(Data/VCRChunkAlign.py:690-741) and the chunk mask / offsets semantics of the
offline chunker (utils/GetChunk_v4_vcr.py:98-159):

* ``input_ids``  int64 [N,T]  ``[CLS] premise [SEP] answer [SEP]``, pad 0, N = 4*B
* ``token_type_ids`` int64 [N,T]
* ``input_mask`` float32 [N,T+R]  (text mask || region mask)
* ``img_feat``   float32 [N,R,2054]  (same image for the 4 choices of a question)
* ``label``      int64 [N] (1 = correct choice) -- ``label_onehot`` float32 [N]
* ``chunk_attention_mask`` float32 [N,T,T] 0/1
* ``gather_index`` list of N int64 tensors (chunk id of tokens 1..L-2)
* ``offsets``    list of N lists of lists of token indices
* ``total_label`` / ``align_pos`` int64 [N,T]
"""
import torch

CLS_ID, SEP_ID, DET_BASE = 101, 102, 30522


def make_batch(B=2, T=128, R=50, feat_dim=2054, num_choices=4, seed=88, rank=0,
               min_len=None, n_align=4, region_pad=0, device="cpu"):
    g = torch.Generator().manual_seed(seed + rank)
    N = B * num_choices
    min_len = max(6, T // 2) if min_len is None else min_len
    lens = torch.randint(min_len, T + 1, (N,), generator=g)
    lens[int(torch.randint(0, N, (1,), generator=g))] = T  # one full-length row
    input_ids = torch.zeros(N, T, dtype=torch.int64)
    token_type = torch.zeros(N, T, dtype=torch.int64)
    text_mask = torch.zeros(N, T, dtype=torch.float32)
    chunk_mask = torch.zeros(N, T, T, dtype=torch.float32)
    total_label = torch.zeros(N, T, dtype=torch.int64)
    gather_index, offsets = [], []
    for n in range(N):
        L = int(lens[n])
        ids = torch.randint(1000, 30000, (L,), generator=g)
        ids[0] = CLS_ID
        ids[L - 1] = SEP_ID
        split = int(torch.randint(2, L - 2, (1,), generator=g))
        ids[split] = SEP_ID
        token_type[n, split + 1:L] = 1
        text_mask[n, :L] = 1.0
        # chunks: consecutive runs of 1..4 tokens covering 1..L-2
        m = torch.eye(L)
        m[0, :L] = 1.0
        m[L - 1, :L] = 1.0
        offs, gi, t = [], [], 1
        while t <= L - 2:
            run = int(torch.randint(1, 5, (1,), generator=g))
            run = min(run, L - 1 - t)
            idx = list(range(t, t + run))
            for a in idx:
                for b in idx:
                    m[a, b] = 1.0
            gi.extend([len(offs)] * run)
            offs.append(idx)
            t += run
        chunk_mask[n, :L, :L] = m
        offsets.append(offs)
        gather_index.append(torch.tensor(gi, dtype=torch.int64))
        # aligned ("<|detK|>") tokens
        k = min(n_align, L - 2)
        pos = torch.randperm(L - 2, generator=g)[:k] + 1
        reg = torch.randint(1, min(R - region_pad - 1, 44) + 1, (k,), generator=g)  # class id < R
        for p_, r_ in zip(pos.tolist(), reg.tolist()):
            if p_ == split:
                continue
            total_label[n, p_] = r_
            ids[p_] = DET_BASE + r_
        input_ids[n, :L] = ids
    align_pos = (total_label != 0).to(torch.int64)
    feats = torch.randn(B, R, feat_dim, generator=g)
    img_feat = feats.repeat_interleave(num_choices, dim=0).contiguous()
    img_mask = torch.ones(N, R, dtype=torch.float32)
    if region_pad > 0:  # some questions have fewer boxes than the batch max
        for b in range(0, B, 2):
            img_mask[b * num_choices:(b + 1) * num_choices, R - region_pad:] = 0.0
            img_feat[b * num_choices:(b + 1) * num_choices, R - region_pad:] = 0.0
    input_mask = torch.cat([text_mask, img_mask], dim=1)
    correct = torch.randint(0, num_choices, (B,), generator=g)
    label = torch.zeros(N, dtype=torch.int64)
    label[torch.arange(B) * num_choices + correct] = 1
    batch = dict(
        input_ids=input_ids, token_type_ids=token_type, input_mask=input_mask,
        img_feat=img_feat, label=label, label_onehot=label.to(torch.float32),
        chunk_attention_mask=chunk_mask, gather_index=gather_index, offsets=offsets,
        total_label=total_label, align_pos=align_pos)
    if device != "cpu":
        batch = to_device(batch, device)
    return batch


def to_device(batch, device, non_blocking=False):
    out = {}
    for k, v in batch.items():
        if torch.is_tensor(v):
            out[k] = v.to(device, non_blocking=non_blocking)
        elif k == "gather_index":
            out[k] = [t.to(device, non_blocking=non_blocking) for t in v]
        else:
            out[k] = v
    return out


def make_prefix(L=12, P=10, H=768, seed=7):
    """Learned per-layer prefix hidden states ~ N(0, 0.02): [L,P,H]."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(L, P, H, generator=g) * 0.02
