"""Device-side batch packer (csrc/pack.cu + modcr_b200/packer.py) against the CPU restatement of the
reference collate (oracle/collate.py: Data/VCRChunkAlign.py:690-741, utils/GetChunk_v4_vcr.py:104-140).
Integer / mask / index work: bit-exact."""
import pytest
import torch

import collate
import restate
import modcr_b200  # noqa: F401
from modcr_b200 import engine as E
from modcr_b200 import synth
from modcr_b200.packer import BatchPacker, unbatch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("B,T,R,region_pad,seed", [(2, 24, 7, 0, 3), (3, 70, 11, 4, 5), (8, 128, 50, 0, 88), (2, 200, 100, 10, 9)])
def test_packer_matches_restated_collate_bit_exact(B, T, R, region_pad, seed):
    b = synth.make_batch(B=B, T=T, R=R, seed=seed, region_pad=region_pad, min_len=6 if T < 30 else None)
    samples = unbatch(b)
    ref = collate.snli_gpt_gen_collate(samples)
    pk = BatchPacker(DEV, dense_chunk_mask=True)
    out = pk.pack(samples)
    for k in ("input_ids", "token_type_ids", "total_label", "align_pos", "input_mask", "img_feat", "label"):
        assert out[k].dtype == ref[k].dtype and tuple(out[k].shape) == tuple(ref[k].shape), k
        assert torch.equal(out[k].cpu(), ref[k]), k
    assert torch.equal(out["label_long"].cpu(), ref["label"].long())
    assert torch.equal(out["dense_chunk_mask"].cpu(), ref["chunk_attention_mask"])
    # packed forms == what the engine derives from the reference tensors
    bits, _ = E.make_chunk_bits(ref["chunk_attention_mask"].to(DEV))
    assert torch.equal(out["chunk_attention_mask"], bits)
    cid, mx = E.pack_chunk_ids(ref["gather_index"], ref["input_ids"].shape[1], DEV)
    assert torch.equal(out["gather_index"], cid) and out["max_chunks"] == mx
    # the four choices of a question travel as one feature block
    N, Tm, Rm = ref["input_ids"].shape[0], ref["input_ids"].shape[1], ref["img_feat"].shape[1]
    assert pk.h2d_bytes <= N * 16 + N * Tm * 8 + 16 + (N // 4) * Rm * 2054 * 4


def test_packer_degenerate_samples():
    """[CLS][SEP]-only sequences (no chunk), one-token chunks, a chunk spanning all inner tokens, fewer boxes than rows."""
    ss = collate.edge_case_samples()
    ref = collate.snli_gpt_gen_collate(ss)
    out = BatchPacker(DEV, dense_chunk_mask=True).pack(ss)
    for k in ("input_ids", "token_type_ids", "total_label", "align_pos", "input_mask", "img_feat", "label"):
        assert torch.equal(out[k].cpu(), ref[k]), k
    assert torch.equal(out["dense_chunk_mask"].cpu(), ref["chunk_attention_mask"])
    bits, _ = E.make_chunk_bits(ref["chunk_attention_mask"].to(DEV))
    assert torch.equal(out["chunk_attention_mask"], bits)
    assert out["gather_index"][2].tolist() == [-1, 0, 0, 1, 2, 2, -1] and bool((out["gather_index"][0] == -1).all())


def test_packer_fixed_shapes_and_model_equivalence():
    """pad_T / pad_R (static shapes for graph replay) and: the model fed by the packer == the model fed by the
    reference-shaped batch, bit for bit (same kernels, same inputs)."""
    from modcr_b200.modeling_bert import BertImgModel
    from modcr_b200.modeling_utils import BertConfig
    from modcr_b200.modeling_vcr_chunkalign_v10 import ChunkAlign_CLS_enc4_align, SeqBertImgModel
    CFG = restate.default_cfg()
    B, T, R = 2, 40, 9
    b = synth.make_batch(B=B, T=T, R=R, seed=4)
    samples = unbatch(b)
    out = BatchPacker(DEV, pad_T=T, pad_R=R).pack(samples)
    assert tuple(out["input_ids"].shape) == (4 * B, T) and tuple(out["img_feat"].shape) == (4 * B, R, 2054)
    torch.manual_seed(0)
    cfg = BertConfig(vocab_size=CFG["vocab_size"], hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                     num_hidden_layers=12)
    m = ChunkAlign_CLS_enc4_align(BertImgModel(cfg), SeqBertImgModel(cfg), 4).to(DEV).eval()
    bd = synth.to_device(b, DEV)
    keys = ("input_mask", "token_type_ids", "offsets", "chunk_attention_mask", "gather_index", "align_pos", "total_label")
    with torch.no_grad():
        r1 = m(bd["input_ids"], bd["img_feat"], label=bd["label"], **{k: bd[k] for k in keys})
        r2 = m(out["input_ids"], out["img_feat"], label=out["label"], **{k: out[k] for k in keys})
    assert float(r1[0]) == float(r2[0]) and float(r1[2]) == float(r2[2])
    assert torch.equal(r1[1], r2[1]) and int(r1[3]) == int(r2[3]) and int(r1[4]) == int(r2[4])
    with pytest.raises(ValueError, match="pad_T"):
        BatchPacker(DEV, pad_T=T - 8, pad_R=R).pack(samples)


def test_collate_takes_the_reference_dataloader_tuples():
    """``BatchPacker.collate`` as ``collate_fn``: per-question tuples of the reference's per-choice 19-tuples
    (Data/VCRChunkAlign.py:683-686, 692-693) in, the reference's batch keys out (incl. the padded language-model side)."""
    from modcr_b200.packer import REF_FIELDS
    b = synth.make_batch(B=2, T=24, R=7, seed=3, min_len=6)
    samples = unbatch(b)
    g = torch.Generator().manual_seed(1)
    for i, s_ in enumerate(samples):
        n_r = int(torch.randint(5, 20, (1,), generator=g))
        s_.update(img_id="q%d" % (i // 4), image="", text="", ques="?", ans="a%d" % i,
                  r_input_ids=torch.randint(3, 99, (n_r,), generator=g), r_segment_ids=torch.zeros(n_r, dtype=torch.int64),
                  r_input_mask=torch.ones(n_r))
    inputs = [tuple(tuple(s_[f] for f in REF_FIELDS) for s_ in samples[q:q + 4]) for q in range(0, len(samples), 4)]
    out = BatchPacker(DEV).collate(inputs)
    ref = collate.snli_gpt_gen_collate(samples)
    for k in ("input_ids", "token_type_ids", "total_label", "align_pos", "input_mask", "img_feat", "label",
              "r_input_ids", "r_attention_mask", "r_token_type_ids"):
        assert out[k].is_cuda and torch.equal(out[k].cpu(), ref[k]), k
    assert out["img_id"] == [s_["img_id"] for s_ in samples] and out["ans_str"] == [s_["ans"] for s_ in samples]
