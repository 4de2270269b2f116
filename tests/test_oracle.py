"""CPU tests of the oracle restatement: golden vectors produced by the REAL
reference (oracle/make_golden.py), the real reference itself when it is present
(build container), the state_dict surface, and the known-answer invariants of
SURVEY.md section 4."""
import math
import os

import pytest
import torch

import params
import ref_shim
import restate
import statedict
import modcr_b200  # noqa: F401
from modcr_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CFG = restate.default_cfg()


def _weights(seed, cfg=CFG):
    return params.fill_params(statedict.materialise(statedict.chunkalign_keys(cfg)), seed)


def _case(name):
    g = torch.load(os.path.join(GOLD, name + ".pt"))
    m = g["meta"]
    cfg = restate.default_cfg(**m.get("cfg_over", {}))          # cfg-5: the Oscar-large overrides
    b = synth.make_batch(B=m["B"], T=m["T"], R=m["R"], seed=m["seed"], region_pad=m["region_pad"])
    prefix = (synth.make_prefix(cfg["num_hidden_layers"], m["P"], cfg["hidden_size"], seed=m["seed"] + 1)
              if m["P"] > 0 else None)
    return g, b, prefix, _weights(m["wseed"], cfg), cfg


@pytest.mark.parametrize("name", ["ragged_small", "cfg1_prefix10", "cfg1_noprefix", "cfg3_vcr", "cfg5_large"])
def test_restatement_matches_reference_golden(name):
    g, b, prefix, sd, cfg = _case(name)
    grad_names = [k for k in g["grads"] if "[" not in k]
    for k in grad_names:
        sd[k].requires_grad_(True)
    if prefix is not None:
        prefix.requires_grad_(True)
    with torch.set_grad_enabled(name == "ragged_small"):
        (loss_cls, matched, align_loss, correct, total), ex = restate.chunkalign_forward(sd, cfg, b, prefix)
    # same torch ops on the same machine class: tight tolerances
    assert abs(float(loss_cls) - float(g["loss_cls"])) < 1e-5
    assert abs(float(align_loss) - float(g["align_loss"])) < 1e-5
    assert bool((matched == g["matched"]).all())
    assert correct == g["correct"] and total == g["total_sum"]
    torch.testing.assert_close(ex["logits"], g["logits"], atol=1e-5, rtol=1e-4)
    torch.testing.assert_close(ex["align_block"], g["align_block"], atol=1e-5, rtol=1e-4)
    torch.testing.assert_close(ex["global_output"][:, ::7, ::5], g["global_output_sub"], atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(ex["seq_output"][:, ::7, ::5], g["seq_output_sub"], atol=2e-5, rtol=1e-4)
    torch.testing.assert_close(ex["chunk_hidden_states"][:, ::7, ::5], g["chunk_hidden_sub"], atol=2e-5, rtol=1e-4)
    if name == "ragged_small":  # backward too (small case keeps the CPU suite fast)
        (loss_cls + align_loss).backward()
        for k in grad_names:
            torch.testing.assert_close(sd[k].grad, g["grads"][k], atol=2e-6, rtol=2e-3)
        torch.testing.assert_close(prefix.grad[:, :, ::4], g["grads"]["prefix[:,:,::4]"], atol=2e-6, rtol=2e-3)


def test_restatement_matches_reference_golden_forward_many():
    """Two of the 32 batches of the 256-question forward-only golden (tests/golden/cfg2_fwd256.pt, REAL reference):
    the restatement reproduces logits / argmax / alignment statistics of each."""
    g = torch.load(os.path.join(GOLD, "cfg2_fwd256.pt"))
    m = g["meta"]
    sd = _weights(m["wseed"])
    prefix = synth.make_prefix(12, m["P"], 768, seed=m["seed0"] - 1)
    for i in (0, m["n_batches"] - 1):
        b = synth.make_batch(B=m["B"], T=m["T"], R=m["R"], seed=m["seed0"] + i)
        with torch.no_grad():
            (_, matched, align_loss, correct, total), ex = restate.chunkalign_forward(sd, CFG, b, prefix)
        torch.testing.assert_close(ex["logits"], g["logits"][i], atol=1e-5, rtol=1e-4)
        assert bool((ex["pre"] == g["pre"][i]).all()) and bool((matched == g["matched"][i]).all())
        assert correct == g["correct"][i] and total == g["total_sum"][i]
        assert abs(float(align_loss) - g["align_loss"][i]) < 1e-5


@pytest.mark.reference
@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")
def test_restatement_matches_live_reference_and_state_dict_surface():
    ns = ref_shim.load()
    cfg = ref_shim.make_config(ns)
    m = ns.v10.ChunkAlign_CLS_enc4_align(ns.tfm.BertImgModel(cfg), ns.v10.SeqBertImgModel(cfg), 4)
    ref_keys = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert ref_keys == dict(statedict.chunkalign_keys(CFG))
    me = ns.v10.ChunkAlign_CLS_enc4_align_ensemble(ns.tfm.BertImgModel(cfg), ns.v10.SeqBertImgModel(cfg), 4)
    assert {k: tuple(v.shape) for k, v in me.state_dict().items()} == dict(statedict.chunkalign_ensemble_keys(CFG))
    sd = params.fill_params(m.state_dict(), 3)
    m.load_state_dict(sd)
    m.eval()
    b = synth.make_batch(B=1, T=24, R=7, seed=3)
    with torch.no_grad():
        ref = m(b["input_ids"], b["img_feat"], input_mask=b["input_mask"], label=b["label"],
                token_type_ids=b["token_type_ids"], offsets=b["offsets"],
                chunk_attention_mask=b["chunk_attention_mask"], gather_index=b["gather_index"],
                align_pos=b["align_pos"], total_label=b["total_label"])
        (o, _) = restate.chunkalign_forward({k: v.clone() for k, v in sd.items()}, CFG, b, None)
    assert float((ref[0] - o[0]).abs()) == 0.0 and float((ref[2] - o[2]).abs()) == 0.0
    assert bool((ref[1] == o[1]).all()) and ref[3] == o[3] and ref[4] == o[4]
    # ensemble head
    sde = params.fill_params(me.state_dict(), 4)
    me.load_state_dict(sde)
    me.eval()
    with torch.no_grad():
        r2 = me(b["input_ids"], b["img_feat"], input_mask=b["input_mask"], label=b["label_onehot"],
                token_type_ids=b["token_type_ids"], offsets=b["offsets"],
                chunk_attention_mask=b["chunk_attention_mask"], gather_index=b["gather_index"],
                align_pos=b["align_pos"], total_label=b["total_label"])
        c2, al2 = restate.chunkalign_ensemble_forward({k: v.clone() for k, v in sde.items()}, CFG, b, None)
    assert float((r2[0] - c2).abs().max()) == 0.0 and float((r2[1] - al2).abs()) == 0.0


@pytest.mark.reference
@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")
@pytest.mark.parametrize("which,classifier,use_ln,loss_type", [
    ("image", "mlp", 1, "bce"), ("image", "mlp", None, "sfmx"), ("image", "linear", 1, "bce"),
    ("oscar", "mlp", 1, "bce"), ("oscar", "linear", 1, "bxe")])
def test_restated_mc_heads_match_live_reference(which, classifier, use_ln, loss_type):
    """the restated classifier / loss variants of ImageBertForMultipleChoice and OscarForMultipleChoice
    (modeling/modeling_bert.py:492-568, 574-712) against the live reference classes (2-layer encoder to stay fast)."""
    ns = ref_shim.load()
    B, C, T, R = 2, 4, 12, 5
    num_labels = 4 if which == "image" else 2
    cfg = ref_shim.make_config(ns, num_hidden_layers=2, classifier=classifier, use_img_layernorm=use_ln,
                               loss_type=loss_type, num_labels=num_labels, num_choice=C, output_attentions=False)
    torch.manual_seed(5)
    cls = ns.mbert.ImageBertForMultipleChoice if which == "image" else ns.mbert.OscarForMultipleChoice
    m = cls(cfg).eval()
    b = synth.make_batch(B=B, T=T, R=R, seed=21)
    shp = lambda t: t.reshape(B, C, *t.shape[1:])  # noqa: E731
    g = torch.Generator().manual_seed(3)
    if loss_type == "bce":
        labels = torch.rand(B, num_labels, generator=g) if which == "image" else torch.rand(B * C, num_labels, generator=g)
    elif which == "image":
        labels = torch.randint(0, num_labels, (B,), generator=g)
    else:
        labels = b["label"].reshape(B, C)
    with torch.no_grad():
        ref_loss, ref_logits = m(shp(b["input_ids"]), token_type_ids=shp(b["token_type_ids"]),
                                 attention_mask=shp(b["input_mask"]), labels=labels, img_feats=shp(b["img_feat"]))[:2]
        pooled = m.bert(b["input_ids"], token_type_ids=b["token_type_ids"], attention_mask=b["input_mask"],
                        img_feats=b["img_feat"])[1]
        sd = m.state_dict()
        if which == "image":
            loss, logits = restate.image_bert_mc_head(sd, pooled, labels, C, num_labels, loss_type, cfg.layer_norm_eps)
        else:
            loss, logits = restate.oscar_mc_head(sd, pooled, labels, num_labels, loss_type, cfg.layer_norm_eps)
    assert float((logits - ref_logits).abs().max()) == 0.0 and float((loss - ref_loss).abs()) == 0.0
    # state_dict surface of our drop-in class
    from modcr_b200.modeling_bert import ImageBertForMultipleChoice, OscarForMultipleChoice, BertConfig as OurCfg
    ocfg = OurCfg(vocab_size=cfg.vocab_size, num_hidden_layers=2)
    for k in ("classifier", "use_img_layernorm", "loss_type", "num_labels", "num_choice"):
        setattr(ocfg, k, getattr(cfg, k))
    ours = (ImageBertForMultipleChoice if which == "image" else OscarForMultipleChoice)(ocfg)
    ours_keys = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    ref_keys = {k: tuple(v.shape) for k, v in sd.items()}
    assert ours_keys == ref_keys


@pytest.mark.reference
@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")
def test_restated_optional_residuals_match_live_reference():
    """--add_residual / --add_local_residual of the run scripts (run_PMR_ModCR.py:574-575, 747-748; v10:212-223)."""
    ns = ref_shim.load()
    cfg = ref_shim.make_config(ns, add_residual=True, add_local_residual=True)
    m = ns.v10.SeqBertImgModel(cfg).eval()
    sd = params.fill_params(m.state_dict(), 6)
    m.load_state_dict(sd)
    b = synth.make_batch(B=1, T=20, R=6, seed=9)
    T = 20
    with torch.no_grad():
        (r_out, r_pool, _), r_ch = m(b["input_ids"], img_feats=b["img_feat"], img_mask=b["input_mask"][:, T:],
                                     input_mask=b["input_mask"], attention_mask=b["chunk_attention_mask"],
                                     token_type_ids=b["token_type_ids"], head_mask=None, offsets=b["offsets"],
                                     gather_index=b["gather_index"])
        cfg_r = dict(CFG, add_residual=True, add_local_residual=True)
        (o, pool, _), ch = restate.seq_enc_forward({"seq_enc." + k: v for k, v in sd.items()}, cfg_r, b["input_ids"],
                                                   b["token_type_ids"], b["chunk_attention_mask"], b["input_mask"],
                                                   b["img_feat"], b["offsets"], b["gather_index"])
        (o0, _, _), _ = restate.seq_enc_forward({"seq_enc." + k: v for k, v in sd.items()}, CFG, b["input_ids"],
                                                b["token_type_ids"], b["chunk_attention_mask"], b["input_mask"],
                                                b["img_feat"], b["offsets"], b["gather_index"])
    assert float((r_out - o).abs().max()) == 0.0 and float((r_pool - pool).abs().max()) == 0.0
    assert float((r_ch - ch).abs().max()) == 0.0
    assert float((o - o0).abs().max()) > 0.1     # the flags change the result


def test_restated_collate_reproduces_the_synthetic_batches():
    """oracle/collate.py (Data/VCRChunkAlign.py:690-741, utils/GetChunk_v4_vcr.py:104-140) on the ragged samples of a
    synthetic batch gives back exactly the padded tensors the golden vectors were produced from."""
    import collate
    from modcr_b200.packer import unbatch
    for kw in (dict(B=2, T=24, R=7, seed=3, min_len=6), dict(B=3, T=70, R=11, seed=5, region_pad=4)):
        b = synth.make_batch(**kw)
        samples = unbatch(b)
        for n, s in enumerate(samples):
            L = s["input_ids"].numel()
            assert torch.equal(collate.chunk_mask_from_offsets(L, s["offsets"]), s["chunk_mask"])
            assert torch.equal(collate.gather_index_from_offsets(s["offsets"]), b["gather_index"][n])
        out = collate.snli_gpt_gen_collate(samples)
        full = any(int(s["input_ids"].numel()) == kw["T"] for s in samples)
        assert full   # make_batch always has one full-length row, so the padded width is T
        for k in ("input_ids", "token_type_ids", "input_mask", "img_feat", "chunk_attention_mask", "total_label",
                  "align_pos"):
            assert torch.equal(out[k], b[k]), k
        assert torch.equal(out["label"], b["label_onehot"])
        # host half of the device packer (packer.BatchPacker.host_pack, no CUDA needed): its descriptor buffer decodes
        # to exactly the collated batch, and holds one feature block per question
        from modcr_b200.packer import BatchPacker
        buf, lay = BatchPacker("cpu").host_pack(samples)
        dec = collate.decode_descriptors(buf, lay, 2054)
        for k in ("input_ids", "token_type_ids", "input_mask", "img_feat", "chunk_attention_mask", "total_label",
                  "align_pos", "label"):
            assert torch.equal(dec[k], out[k]), k
        assert lay["U"] == kw["B"] and lay["total"] <= lay["feat_off"] + kw["B"] * lay["R"] * 2054 * 4
        with pytest.raises(RuntimeError, match="no CPU path"):
            BatchPacker("cpu").pack(samples)
        # the DataLoader-shaped input of the reference collate (per-question tuples of per-choice 19-tuples) maps back
        # to the same samples, and the language-model side is padded like :695-697
        from modcr_b200.packer import REF_FIELDS, pad_language_side, samples_from_reference
        g = torch.Generator().manual_seed(1)
        for s_ in samples:
            n_r = int(torch.randint(5, 20, (1,), generator=g))
            s_.update(img_id="q", image="", text="", ques="?", ans="!", r_input_ids=torch.randint(3, 99, (n_r,), generator=g),
                      r_segment_ids=torch.zeros(n_r, dtype=torch.int64), r_input_mask=torch.ones(n_r))
        inputs = [tuple(tuple(s_[f] for f in REF_FIELDS) for s_ in samples[q:q + 4]) for q in range(0, len(samples), 4)]
        back = samples_from_reference(inputs)
        assert len(back) == len(samples) and all(a[f] is b_[f] for a, b_ in zip(back, samples) for f in REF_FIELDS)
        lang, ref_lang = pad_language_side(back), collate.snli_gpt_gen_collate(samples)
        for k in ("r_input_ids", "r_attention_mask", "r_token_type_ids"):
            assert torch.equal(lang[k], ref_lang[k]), k
        with pytest.raises(ValueError, match="19-field|19"):
            samples_from_reference([((1, 2, 3),)])


def test_packer_host_half_on_degenerate_samples():
    """[CLS][SEP]-only sequences (no chunk), one-token chunks, a chunk spanning all inner tokens, fewer boxes than rows."""
    import collate
    from modcr_b200.packer import BatchPacker
    ss = collate.edge_case_samples()
    ref = collate.snli_gpt_gen_collate(ss)
    buf, lay = BatchPacker("cpu").host_pack(ss)
    dec = collate.decode_descriptors(buf, lay, 2054)
    assert (lay["T"], lay["R"], lay["U"]) == (7, 2, 1)
    for k in ("input_ids", "token_type_ids", "input_mask", "img_feat", "chunk_attention_mask", "total_label", "align_pos",
              "label"):
        assert torch.equal(dec[k], ref[k]), k
    assert bool((dec["chunk_id"][0] == -1).all()) and dec["chunk_id"][2].tolist() == [-1, 0, 0, 1, 2, 2, -1]


# ---- known-answer invariants, small shapes ----------------
def _small():
    sd = _weights(11)
    b = synth.make_batch(B=1, T=20, R=6, seed=9)
    return sd, b


def test_masked_prefix_equals_no_prefix_and_live_prefix_differs():
    sd, b = _small()
    N, P = 4, 3
    hist = [torch.randn(N, P, 768) * 0.5 for _ in range(12)]
    with torch.no_grad():
        base, _, _ = restate.global_enc_forward(sd, CFG, b["input_ids"], b["token_type_ids"], b["input_mask"],
                                                b["img_feat"], None)
        masked, _, _ = restate.global_enc_forward(sd, CFG, b["input_ids"], b["token_type_ids"],
                                                  torch.cat([torch.zeros(N, P), b["input_mask"]], -1),
                                                  b["img_feat"], hist)
        live, _, _ = restate.global_enc_forward(sd, CFG, b["input_ids"], b["token_type_ids"],
                                                torch.cat([torch.ones(N, P), b["input_mask"]], -1),
                                                b["img_feat"], hist)
    assert float((base - masked).abs().max()) == 0.0
    assert float((base - live).abs().max()) > 1e-3


def test_align_loss_near_log_R_and_block_bounds_and_phase_masks():
    sd, b = _small()
    with torch.no_grad():
        (out, ex) = restate.chunkalign_forward(sd, CFG, b, None)
        (s_out, _, att), _ = restate.seq_enc_forward(sd, CFG, b["input_ids"], b["token_type_ids"],
                                                     b["chunk_attention_mask"], b["input_mask"], b["img_feat"],
                                                     b["offsets"], b["gather_index"])
    T, R = 20, 6
    assert abs(float(out[2]) - math.log(R)) < 0.05
    blk = ex["align_block"]
    assert float(blk.sum(-1).max()) <= 36.0 + 1e-4 and float(blk.sum(-1).min()) > 0
    for l in (9, 10, 11):   # image rows: identity only
        p = att[l]
        assert bool((torch.diagonal(p[:, :, T:, T:], dim1=-2, dim2=-1) == 1.0).all())
        assert float(p[:, :, T:, :T].abs().max()) == 0.0
    for l in (0, 1, 2):     # phase A: image->text exactly 0, text->text outside chunk exactly 0
        p = att[l]
        assert float(p[:, :, T:, :T].abs().max()) == 0.0
        outside = (b["chunk_attention_mask"] == 0).unsqueeze(1).expand(-1, 12, -1, -1)
        assert float(p[:, :, :T, :T][outside].abs().max()) == 0.0


def test_singleton_chunks_make_chunk_mean_identity():
    sd, b = _small()
    N, T = b["input_ids"].shape
    lens = b["input_mask"][:, :T].sum(-1).long()
    b2 = dict(b)
    b2["offsets"] = [[[t] for t in range(1, int(L) - 1)] for L in lens]
    b2["gather_index"] = [torch.arange(int(L) - 2) for L in lens]
    q = torch.randn(N, T + 6, 768)
    assert float((restate.chunk_mean_queries(q, b2["offsets"], b2["gather_index"]) - q).abs().max()) == 0.0
    assert float((restate.chunk_mean_queries(q, b["offsets"], b["gather_index"]) - q).abs().max()) > 0


def test_choice_permutation_permutes_logits():
    sd, b = _small()
    perm = torch.tensor([2, 0, 3, 1])
    b2 = dict(b)
    for k in ("input_ids", "token_type_ids", "input_mask", "img_feat", "label", "chunk_attention_mask",
              "total_label", "align_pos"):
        b2[k] = b[k][perm]
    b2["gather_index"] = [b["gather_index"][i] for i in perm.tolist()]
    b2["offsets"] = [b["offsets"][i] for i in perm.tolist()]
    with torch.no_grad():
        _, e1 = restate.chunkalign_forward(sd, CFG, b, None)
        _, e2 = restate.chunkalign_forward(sd, CFG, b2, None)
    torch.testing.assert_close(e1["logit_expl"][0][perm], e2["logit_expl"][0], atol=1e-6, rtol=1e-5)


def test_mc_scorers():
    torch.manual_seed(0)
    pooled = torch.randn(8, 16)
    w, bias = torch.randn(1, 16), torch.randn(1)
    onehot = torch.tensor([0, 1, 0, 0, 0, 0, 0, 1.0])
    loss, logits = restate.abstract_scorer(pooled, w, bias, onehot)
    ref = -(torch.log_softmax(logits, -1) * onehot.view(2, 4)).sum(-1).mean()
    assert abs(float(loss - ref)) < 1e-6
    w4, b4 = torch.randn(3, 64), torch.randn(3)
    loss4, lg4 = restate.choice_flattened_scorer(pooled, w4, b4, torch.tensor([2, 0]))
    assert lg4.shape == (2, 3) and float(loss4) > 0


@pytest.mark.reference
@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")
def test_abstract_specific_restatement_matches_live_reference():
    """the REAL ``Abstract_Specific`` (modeling/modeling_ensemble.py:424-540) with the test stand-in for
    its language model against ``restate.abstract_specific_forward`` -- forward values and the gradients that reach the
    two mapping networks and the ClsLayer_lyx head -- plus the state_dict surface of this package's class."""
    import standin_lm
    from modcr_b200.modeling_bert import BertImgModel
    from modcr_b200.modeling_ensemble import Abstract_Specific
    from modcr_b200.modeling_utils import BertConfig
    from modcr_b200.modeling_vcr_chunkalign_v10 import ChunkAlign_CLS_enc4_align_ensemble, SeqBertImgModel
    ns = ref_shim.load_ensemble()
    cfg = ref_shim.make_config(ns)
    calec = ns.v10.ChunkAlign_CLS_enc4_align_ensemble(ns.tfm.BertImgModel(cfg), ns.v10.SeqBertImgModel(cfg), 4)
    lm = standin_lm.StandInLM()
    m = ns.ens.Abstract_Specific(roberta_model=lm, calec_model=calec, clip_model=None, num_labels=4)
    lm_sd = {k: v.clone() for k, v in lm.state_dict().items()}
    sd = params.fill_params(m.state_dict(), 6)   # fills in place (the state_dict aliases the parameters)
    for k, v in lm_sd.items():
        sd["roberta." + k] = v                   # keep the stand-in's own deterministic weights
    m.load_state_dict(sd)
    m.eval()
    mine_cfg = BertConfig(vocab_size=CFG["vocab_size"], hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    mine = Abstract_Specific(standin_lm.StandInLM(),
                             ChunkAlign_CLS_enc4_align_ensemble(BertImgModel(mine_cfg), SeqBertImgModel(mine_cfg), 4))
    assert {k: tuple(v.shape) for k, v in mine.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}

    b = synth.make_batch(B=1, T=24, R=7, seed=3)
    rb = standin_lm.make_inputs(4, seed=1)
    out = m(None, None, rb["r_input_ids"], rb["r_token_type_ids"], rb["r_attention_mask"], b["input_ids"], b["img_feat"],
            input_mask=b["input_mask"], token_type_ids=b["token_type_ids"], offsets=b["offsets"],
            chunk_attention_mask=b["chunk_attention_mask"], gather_index=b["gather_index"], label=b["label_onehot"],
            align_pos=b["align_pos"], total_label=b["total_label"])
    out[0].backward()
    sdr = {k: v.clone() for k, v in sd.items()}
    watch = [k for k in sdr if k.startswith("mapping_network_") or k.startswith("abst_confidence_scorer.")
             or k.startswith("calec.cls_ensemble_1.") or k.startswith("calec.cls_layer_lyx.0.cross_attention.v_proj.")]
    for k in watch:
        sdr[k].requires_grad_(True)
    lm2 = standin_lm.StandInLM()
    loss, logits, prefix_emb, cls = restate.abstract_specific_forward(sdr, CFG, b, lm2, rb)
    loss.backward()
    assert out[1][0] is None and out[1][1] is None and out[1][3] is None
    assert float((out[0] - loss).detach().abs()) < 1e-6 and float((out[1][2] - loss).detach().abs()) < 1e-6
    torch.testing.assert_close(out[2], logits, atol=1e-6, rtol=1e-5)
    named = dict(m.named_parameters())
    for k in watch:
        torch.testing.assert_close(named[k].grad, sdr[k].grad, atol=1e-7, rtol=1e-4)


def test_standin_roberta_oracle_matches_huggingface_roberta():
    """SURVEY 8 f4 (parity UNPINNED vs the reference's missing local_transformers blob): the stand-in oracle of the
    language-model side (oracle/standin_roberta.py) against HuggingFace RobertaModel of the installed transformers --
    embeddings incl. the padding-aware position ids, prompt embeddings concatenated after the embedding layer, the
    post-LN encoder, the pooler."""
    import standin_roberta
    from transformers import RobertaConfig, RobertaModel
    hc = RobertaConfig(vocab_size=120, hidden_size=128, num_hidden_layers=3, num_attention_heads=2, intermediate_size=256,
                       max_position_embeddings=40, type_vocab_size=1, pad_token_id=1, hidden_dropout_prob=0.0,
                       attention_probs_dropout_prob=0.0, layer_norm_eps=1e-5)
    hc._attn_implementation = "eager"
    torch.manual_seed(3)
    hf = RobertaModel(hc, add_pooling_layer=True).eval()
    sd = {k: v.clone() for k, v in hf.state_dict().items()}
    g = torch.Generator().manual_seed(4)
    ids = torch.randint(3, 120, (3, 12), generator=g)
    ids[1, 8:] = 1
    ids[2, 5:] = 1
    am = (ids != 1).float()
    prompt = torch.randn(3, 4, 128, generator=g)
    pmask = torch.ones(3, 4)
    cfg = dict(num_hidden_layers=3, num_attention_heads=2, layer_norm_eps=1e-5, pad_token_id=1)
    with torch.no_grad():
        # no prompts: the whole HF model, untouched
        ref = hf(input_ids=ids, attention_mask=am)
        seq, pooled = standin_roberta.roberta_prompt_forward(sd, cfg, ids, None, am)
        torch.testing.assert_close(seq * am.unsqueeze(-1), ref.last_hidden_state * am.unsqueeze(-1), atol=2e-5, rtol=1e-4)
        torch.testing.assert_close(pooled, ref.pooler_output, atol=2e-5, rtol=1e-4)
        # prompts: HF's own embedding layer / encoder / pooler objects around the concatenation
        e = hf.embeddings(input_ids=ids, token_type_ids=torch.zeros_like(ids))
        x = torch.cat([prompt, e], 1)
        ext = (1.0 - torch.cat([pmask, am], 1))[:, None, None, :] * -10000.0
        out = hf.encoder(x, attention_mask=ext)
        hseq = out[0] if isinstance(out, tuple) else out.last_hidden_state
        hpool = hf.pooler(hseq)
        seq, pooled = standin_roberta.roberta_prompt_forward(sd, cfg, ids, None, am, prompt, pmask)
        keep = torch.cat([pmask, am], 1).unsqueeze(-1)
        torch.testing.assert_close(seq * keep, hseq * keep, atol=2e-5, rtol=1e-4)
        torch.testing.assert_close(pooled, hpool, atol=2e-5, rtol=1e-4)


@pytest.mark.reference
@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")
def test_restated_gttoken_region_embedding_matches_live_reference():
    """BertImgModel_gtToken (modeling/modeling_bert.py:293-412): box projection Linear(7->H) + LayerNorm, region type
    embedding and feature projection summed, image LayerNorm, encoder, pooler -- restatement vs the live class
    (``FusedLayerNorm``, whose import the reference has commented out, bound to torch.nn.LayerNorm)."""
    ns = ref_shim.load()
    if not hasattr(ns.mbert, "FusedLayerNorm"):
        ns.mbert.FusedLayerNorm = torch.nn.LayerNorm
    cfg = ref_shim.make_config(ns, num_hidden_layers=2, output_attentions=False)
    m = ns.mbert.BertImgModel_gtToken(cfg)
    sd = params.fill_params(m.state_dict(), 5)
    m.load_state_dict(sd)
    m.eval()
    b = synth.make_batch(B=1, T=16, R=6, seed=9)
    g = torch.Generator().manual_seed(3)
    N, R = b["img_feat"].shape[:2]
    pos = torch.rand(N, R, 7, generator=g)
    tt = torch.randint(0, 2, (N, R), generator=g)
    with torch.no_grad():
        ref = m(b["input_ids"], token_type_ids=b["token_type_ids"], attention_mask=b["input_mask"], img_feats=b["img_feat"],
                img_token_type_ids=tt, img_pos_feat=pos)
        ocfg = restate.default_cfg(num_hidden_layers=2)
        seq, pooled, _ = restate.global_enc_forward({k: v for k, v in sd.items()}, ocfg, b["input_ids"], b["token_type_ids"],
                                                    b["input_mask"], b["img_feat"], None, pfx="", img_pos_feat=pos,
                                                    img_token_type_ids=tt)
    assert float((ref[0] - seq).abs().max()) == 0.0 and float((ref[1] - pooled).abs().max()) == 0.0
    assert {"img_token_type_embeddings.weight", "img_pos_layer_norm.weight", "img_pos_layer_norm.bias", "img_pos_linear.weight",
            "img_pos_linear.bias"} <= set(sd.keys())


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present")
def test_restated_3d_attention_mask_matches_live_reference():
    """``BertImgModel.forward`` with a 3-D ``[N, S, S]`` mask (modeling/modeling_transfomres.py:630-631): the restatement
    the GPU test ``test_global_enc_3d_attention_mask_matches_oracle`` compares against, pinned to the live class."""
    ns = ref_shim.load()
    cfg = ref_shim.make_config(ns, num_hidden_layers=2, output_attentions=False)
    m = ns.tfm.BertImgModel(cfg)
    sd = params.fill_params(m.state_dict(), 11)
    m.load_state_dict(sd)
    m.eval()
    b = synth.make_batch(B=1, T=20, R=6, seed=13)
    N, S = b["input_ids"].shape[0], 26
    g = torch.Generator().manual_seed(3)
    mask = (torch.rand(N, S, S, generator=g) < 0.6).float()
    mask[:, torch.arange(S), torch.arange(S)] = 1.0
    mask[1, 3] = 0.0                                          # a row with no visible key
    with torch.no_grad():
        ref = m(b["input_ids"], token_type_ids=b["token_type_ids"], attention_mask=mask, img_feats=b["img_feat"])
        seq, pooled, _ = restate.global_enc_forward({k: v for k, v in sd.items()}, restate.default_cfg(num_hidden_layers=2),
                                                    b["input_ids"], b["token_type_ids"], mask, b["img_feat"], None, pfx="")
    assert float((ref[0] - seq).abs().max()) == 0.0 and float((ref[1] - pooled).abs().max()) == 0.0
