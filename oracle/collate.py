"""ORACLE (test infrastructure, CPU): restatement of the reference's batch assembly for the ModCR path.

* ``chunk_mask_from_offsets``  utils/GetChunk_v4_vcr.py:104-140 -- the dense [L,L] chunk mask the offline chunker stores
  next to ``offsets`` (eye + full first / last row + one block per chunk).
* ``gather_index_from_offsets``  Data/VCRChunkAlign.py:667-670.
* ``snli_gpt_gen_collate``  Data/VCRChunkAlign.py:690-741 (``SNLIGPT_gen_collate``), the fields the hot path consumes
  (including the RoBERTa-side ``r_*`` fields when the samples carry them).

The reference's data module cannot be imported (toolz / lmdb / clip ... are absent), so this is a
line-by-line restatement; tests/test_oracle.py checks it against the synthetic batch generator, whose padded tensors
were used to produce the reference golden vectors.
"""
import torch
from torch.nn.utils.rnn import pad_sequence


def chunk_mask_from_offsets(L, offsets):
    m = torch.eye(L)                                   # :106
    m[0, :L] = 1                                       # :108
    for chunk in offsets:                              # :121-135 (symmetric links inside a chunk)
        chunk = [chunk] if isinstance(chunk, int) else chunk
        for a in chunk:
            for b in chunk:
                m[a][b] = 1
    m[L - 1, :L] = 1                                   # :139
    return m


def gather_index_from_offsets(offsets):
    gi = []
    for idx, s in enumerate(offsets):                  # :667-669
        gi.extend([idx] * len(s))
    return torch.tensor(gi, dtype=torch.int64)


def snli_gpt_gen_collate(samples):
    """``samples``: list over (question, choice) of dicts with the names of the reference's per-choice tuple
    (:683-686): input_ids [L], segment_ids [L], input_mask [L], img_feat [Rmax,D], img_mask [Rmax], target (0-d),
    chunk_mask [L,L], gather_index, offsets, total_label [L], align_pos [L]."""
    g = lambda k: [s[k] for s in samples]  # noqa: E731
    input_mask = pad_sequence(g("input_mask"), batch_first=True, padding_value=0)      # :698
    input_ids = pad_sequence(g("input_ids"), batch_first=True, padding_value=0)        # :699
    total_label = pad_sequence(g("total_label"), batch_first=True, padding_value=0)    # :700
    align_pos = pad_sequence(g("align_pos"), batch_first=True, padding_value=0)        # :701
    segment_ids = pad_sequence(g("segment_ids"), batch_first=True, padding_value=0)    # :703
    target = torch.stack(g("target"), dim=0).type(torch.FloatTensor)                   # :704-705
    img_mask = torch.stack(g("img_mask"), dim=0)                                       # :707
    img_feat = torch.stack(g("img_feat"), dim=0)                                       # :708
    max_img = int(torch.max(torch.sum(img_mask, dim=-1)).item())                       # :709
    img_mask, img_feat = img_mask[:, :max_img], img_feat[:, :max_img]                  # :710-711
    input_mask = torch.cat((input_mask, img_mask), -1)                                 # :712
    max_hypo = input_ids.size(1)
    padded = []
    for item in g("chunk_mask"):                                                       # :715-721
        item = torch.cat((item, torch.zeros((max_hypo - item.size(0), item.size(1)))), 0)
        item = torch.cat((item, torch.zeros((max_hypo, max_hypo - item.size(1)))), 1)
        padded.append(item)
    batch = {"input_ids": input_ids, "token_type_ids": segment_ids, "input_mask": input_mask, "img_feat": img_feat,
             "label": target, "chunk_attention_mask": torch.stack(padded, 0), "gather_index": g("gather_index"),
             "offsets": g("offsets"), "total_label": total_label, "align_pos": align_pos}
    if samples and samples[0].get("r_input_ids") is not None:                           # :695-697, :727-729
        batch["r_input_ids"] = pad_sequence(g("r_input_ids"), batch_first=True, padding_value=0)
        batch["r_attention_mask"] = pad_sequence(g("r_input_mask"), batch_first=True, padding_value=0)
        batch["r_token_type_ids"] = pad_sequence(g("r_segment_ids"), batch_first=True, padding_value=0)
    return batch


def decode_descriptors(buf, lay, D):
    """Numpy restatement of what ``modcr_pack_batch`` (csrc/pack.cu) must produce from the packed descriptor buffer of
    ``packer.BatchPacker.host_pack`` (format: include/modcr_b200.h, ``modcr_pack_desc``).  Returns the batch dict with
    the dense chunk mask; used on the CPU to check the host half of the packer against ``snli_gpt_gen_collate``."""
    import numpy as np
    N, T, R, U = lay["N"], lay["T"], lay["R"], lay["U"]
    nb = buf.numpy()
    header = nb[:lay["hdr_bytes"]].view(np.int32).reshape(N, 4)
    tokens = nb[lay["hdr_bytes"]:lay["hdr_bytes"] + lay["tok_bytes"]].view(np.int32).reshape(N, T, 2)
    feats = nb[lay["feat_off"]:lay["total"]].view(np.float32).reshape(U, R, D)
    ids = torch.zeros(N, T, dtype=torch.int64)
    seg, lab, apos = torch.zeros_like(ids), torch.zeros_like(ids), torch.zeros_like(ids)
    cid = torch.full((N, T), -1, dtype=torch.int32)
    mask = torch.zeros(N, T + R)
    cm = torch.zeros(N, T, T)
    img = torch.zeros(N, R, D)
    for n in range(N):
        L, nbox, _, blk = (int(v) for v in header[n])
        w1 = tokens[n, :L, 1].astype(np.int64) & 0xFFFFFFFF
        ids[n, :L] = torch.from_numpy(tokens[n, :L, 0].astype(np.int64))
        seg[n, :L] = torch.from_numpy((w1 >> 10) & 1)
        apos[n, :L] = torch.from_numpy((w1 >> 11) & 1)
        lab[n, :L] = torch.from_numpy(w1 >> 12)
        c = torch.from_numpy((w1 & 1023) - 1).to(torch.int32)
        cid[n, :L] = c
        mask[n, :L] = 1.0
        mask[n, T:T + nbox] = 1.0
        m = torch.eye(L)
        m[0, :L] = 1
        m[L - 1, :L] = 1
        same = (c[:, None] == c[None, :]) & (c[:, None] >= 0)
        m[same] = 1
        cm[n, :L, :L] = m
        img[n, :nbox] = torch.from_numpy(feats[blk, :nbox].copy())
    return {"input_ids": ids, "token_type_ids": seg, "total_label": lab, "align_pos": apos, "input_mask": mask,
            "chunk_attention_mask": cm, "img_feat": img, "chunk_id": cid,
            "label": torch.from_numpy(header[:, 2].astype(np.float32))}


def edge_case_samples(feat_dim=2054):
    """Degenerate but legal samples for the collate / packer tests: a sequence of only [CLS][SEP] (no chunk at all), a
    single one-token chunk, mixed chunk lengths, one chunk covering every inner token; the four share one feature block
    with fewer boxes than rows."""
    g = torch.Generator().manual_seed(12)
    feat = torch.randn(3, feat_dim, generator=g)
    feat[2:] = 0
    fmask = torch.tensor([1.0, 1.0, 0.0])

    def sample(L, offs):
        ids = torch.arange(100, 100 + L)
        ids[0], ids[-1] = 101, 102
        tl = torch.zeros(L, dtype=torch.int64)
        if L > 2:
            tl[1] = 2
        return dict(input_ids=ids, segment_ids=torch.zeros(L, dtype=torch.int64), input_mask=torch.ones(L), img_feat=feat,
                    img_mask=fmask, target=torch.tensor(int(L == 7)), chunk_mask=chunk_mask_from_offsets(L, offs),
                    gather_index=gather_index_from_offsets(offs), offsets=offs, total_label=tl,
                    align_pos=(tl != 0).long())
    return [sample(2, []), sample(3, [[1]]), sample(7, [[1, 2], [3], [4, 5]]), sample(5, [[1, 2, 3]])]
