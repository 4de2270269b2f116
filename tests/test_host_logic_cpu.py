"""CPU tests of the host-side logic: module surface / state_dict keys, checkpoint helpers, chunk-id packing,
FLOP accounting and the bench JSON helpers."""
import importlib.util
import os

import pytest
import torch

import restate
import statedict
import modcr_b200  # noqa: F401
from modcr_b200 import engine, synth
from modcr_b200.modeling_bert import BertImgModel, ImageBertForMultipleChoice
from modcr_b200.modeling_utils import BertConfig
from modcr_b200.modeling_vcr_chunkalign_v10 import ChunkAlign_CLS_enc4_align, SeqBertImgModel, binary_to_mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = restate.default_cfg()


def small_cfg(**kw):
    d = dict(vocab_size=64, hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=512,
             max_position_embeddings=32, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, img_feature_dim=24)
    d.update(kw)
    return BertConfig(**d)


def test_state_dict_surface_matches_reference_key_list():
    cfg = BertConfig(vocab_size=CFG["vocab_size"])
    m = ChunkAlign_CLS_enc4_align(BertImgModel(cfg), SeqBertImgModel(cfg), 4)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == dict(statedict.chunkalign_keys(CFG))
    n_params = sum(p.numel() for p in m.parameters())
    assert abs(n_params - 250.0e6) < 1.0e6          # 250.0 M parameters


def test_init_weights_follow_bert_init():
    torch.manual_seed(0)
    g = BertImgModel(small_cfg())
    assert float(g.embeddings.word_embeddings.weight[0].abs().max()) == 0.0          # padding row zeroed
    assert float(g.encoder.layer[0].attention.self.query.bias.abs().max()) == 0.0
    assert abs(float(g.encoder.layer[0].intermediate.dense.weight.std()) - 0.02) < 2e-3
    assert float((g.LayerNorm.weight - 1).abs().max()) == 0.0


def test_save_and_from_pretrained_roundtrip_and_gamma_beta_rename(tmp_path):
    torch.manual_seed(1)
    g = BertImgModel(small_cfg())
    g.save_pretrained(str(tmp_path))
    h = BertImgModel.from_pretrained(str(tmp_path), config=small_cfg())
    for (k1, v1), (k2, v2) in zip(g.state_dict().items(), h.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    sd = {("bert." + k).replace("LayerNorm.weight", "LayerNorm.gamma").replace("LayerNorm.bias", "LayerNorm.beta"): v
          for k, v in g.state_dict().items()}
    h2 = BertImgModel.from_pretrained(None, config=small_cfg(), state_dict=sd)
    assert not h2.load_info["missing_keys"]
    assert torch.equal(h2.embeddings.LayerNorm.weight, g.embeddings.LayerNorm.weight)


def test_resize_token_embeddings_keeps_old_rows():
    g = BertImgModel(small_cfg())
    old = g.embeddings.word_embeddings.weight.detach().clone()
    g.resize_token_embeddings(64 + 45)                # the 45 "<|detN|>" tokens of run_PMR_ModCR.py:715
    assert g.embeddings.word_embeddings.weight.shape == (109, 256)
    assert torch.equal(g.embeddings.word_embeddings.weight[:64], old)
    assert g.config.vocab_size == 109


def test_from_pretrained_raises_on_size_mismatch_and_maps_the_bert_prefix_both_ways():
    from modcr_b200.modeling_bert import OscarForMultipleChoice
    torch.manual_seed(2)
    g = BertImgModel(small_cfg())
    sd = g.state_dict()
    # a vocabulary-resized checkpoint must not leave word_embeddings silently at its random init (HF raises too)
    bad = dict(sd)
    bad["embeddings.word_embeddings.weight"] = torch.zeros(64 + 45, 256)
    with pytest.raises(RuntimeError, match="size mismatch"):
        BertImgModel.from_pretrained(None, config=small_cfg(), state_dict=bad)
    # a prefix-less encoder checkpoint loads into the wrapper's ``bert.`` submodule
    w = OscarForMultipleChoice.from_pretrained(None, config=small_cfg(), state_dict=sd)
    assert torch.equal(w.bert.encoder.layer[1].output.dense.weight, g.encoder.layer[1].output.dense.weight)
    assert sorted(w.load_info["missing_keys"]) == ["classifier.bias", "classifier.weight"]
    assert not w.load_info["unexpected_keys"]
    # the wrappers resize / expose the embeddings of their encoder
    assert w.get_input_embeddings() is w.bert.embeddings.word_embeddings
    w.resize_token_embeddings(70)
    assert w.bert.embeddings.word_embeddings.weight.shape == (70, 256) and w.bert.config.vocab_size == 70


def test_cpu_tensors_are_rejected_loudly():
    g = BertImgModel(small_cfg())
    ids = torch.randint(1, 60, (2, 6))
    with pytest.raises(RuntimeError):
        g(ids, img_feats=torch.randn(2, 3, 24), attention_mask=torch.ones(2, 9))


def test_dropout_state_host_logic():
    """Training-mode dropout is driven by ``engine.Drop``: absent in eval / p = 0 (so the deterministic kernels are
    untouched), otherwise a per-pass device seed that advances every pass and distinct site ids per layer and role."""
    engine.manual_seed(7)
    assert engine.make_drop(False, 0.3, 0.1, "cpu", 1 << 20) is None
    assert engine.make_drop(True, 0.0, 0.0, "cpu", 1 << 20) is None
    d1 = engine.make_drop(True, 0.3, 0.1, "cpu", 1 << 20)
    d2 = engine.make_drop(True, 0.3, 0.1, "cpu", 2 << 20)
    assert d1.p_hidden == pytest.approx(0.3) and d1.p_attn == pytest.approx(0.1)
    assert d1.seed.dtype == torch.int32 and int(d1.seed) != int(d2.seed)          # the state advanced
    sites = {d.site(layer, k) for d in (d1, d2) for layer in range(13) for k in range(3)}
    assert len(sites) == 2 * 13 * 3                                              # no two call sites share a mask stream
    engine.manual_seed(7)
    assert int(engine.make_drop(True, 0.3, 0.1, "cpu", 1 << 20).seed) == int(d1.seed)   # reproducible
    # the product path still has no CPU fallback, dropout or not
    g = BertImgModel(small_cfg(hidden_dropout_prob=0.3))
    g.train()
    with pytest.raises(RuntimeError):
        g(torch.randint(1, 60, (2, 6)), img_feats=torch.randn(2, 3, 24), attention_mask=torch.ones(2, 9))


def test_pack_chunk_ids_is_bit_exact():
    b = synth.make_batch(B=2, T=40, R=5, seed=4)
    ids, mx = engine.pack_chunk_ids(b["gather_index"], 40, "cpu")
    assert ids.dtype == torch.int32 and ids.shape == (8, 40)
    for n, gi in enumerate(b["gather_index"]):
        L = gi.numel()
        assert torch.equal(ids[n, 1:1 + L].long(), gi)
        assert int(ids[n, 0]) == -1 and bool((ids[n, 1 + L:] == -1).all())
        # chunk ids agree with the offsets lists of the collate contract (Data/VCRChunkAlign.py:664-670)
        for c, toks in enumerate(b["offsets"][n]):
            assert all(int(ids[n, t]) == c for t in toks)
    assert mx == max(int(g.max()) + 1 for g in b["gather_index"])


def test_binary_to_mp_matches_oracle():
    x = torch.randn(8, 2)
    assert torch.equal(binary_to_mp(x, 4), restate.binary_to_mp(x, 4))


def test_bench_flop_accounting_matches_survey():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    gf = bench.flops_per_sample(bench.WORKLOAD) / 1e9
    assert abs(gf - 778.4) < 1.0                       # 778.4 GF fwd+bwd per sample at cfg-1/2
    peak, src, _ = bench.peaks()
    assert peak > 100 and isinstance(src, str)


def test_multiple_choice_head_surface():
    m = ImageBertForMultipleChoice(small_cfg(num_labels=4, num_choice=4))
    assert m.classifier.weight.shape == (4, 4 * 256)


def test_multiple_choice_classifier_variants_surface():
    """state_dict layout of the classifier variants (modeling/modeling_bert.py:511-530, 656-669; checked key-for-key
    against the live reference classes in tests/test_oracle.py when the reference tree is present)."""
    from modcr_b200.modeling_bert import OscarForMultipleChoice
    cfg = small_cfg(num_labels=4, num_choice=4)
    cfg.classifier, cfg.use_img_layernorm, cfg.loss_type = "mlp", 1, "bce"
    m = ImageBertForMultipleChoice(cfg)
    keys = {k: tuple(v.shape) for k, v in m.state_dict().items() if k.startswith("classifier.")}
    assert keys == {"classifier.0.weight": (512, 1024), "classifier.0.bias": (512,), "classifier.2.weight": (512,),
                    "classifier.2.bias": (512,), "classifier.3.weight": (4, 512), "classifier.3.bias": (4,)}
    cfg2 = small_cfg(num_labels=2, num_choice=4)
    cfg2.classifier, cfg2.loss_type = "mlp", "bce"
    o = OscarForMultipleChoice(cfg2)
    keys = {k: tuple(v.shape) for k, v in o.state_dict().items() if k.startswith("classifier.")}
    assert keys == {"classifier.0.weight": (512, 256), "classifier.0.bias": (512,), "classifier.2.weight": (2, 512),
                    "classifier.2.bias": (2,)}
    cfg3 = small_cfg(num_labels=2)
    cfg3.classifier = "conv"
    with pytest.raises(ValueError, match="linear.*mlp"):
        ImageBertForMultipleChoice(cfg3)
    # CPU tensors never reach a kernel: the head raises instead of falling back
    x = torch.zeros(2, 4, 6, dtype=torch.int64)
    with pytest.raises(RuntimeError):
        m(x, img_feats=torch.zeros(2, 4, 3, 2054))


def test_lr_schedules_match_transformers():
    """``--scheduler linear | constant`` of the run scripts (run_PMR_ModCR.py:137-144): the multipliers handed to
    ``FusedAdamW.set_lr`` equal what transformers' LambdaLR schedulers produce step by step."""
    import transformers
    from modcr_b200.training import constant_schedule_with_warmup, linear_schedule_with_warmup
    for warm, total in ((0, 10), (3, 17), (5, 5)):
        for kind in ("linear", "constant"):
            opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1.0)
            sch = (transformers.get_linear_schedule_with_warmup(opt, warm, total) if kind == "linear"
                   else transformers.get_constant_schedule_with_warmup(opt, warm))
            for step in range(total + 3):
                ours = (linear_schedule_with_warmup(step, warm, total) if kind == "linear"
                        else constant_schedule_with_warmup(step, warm))
                assert abs(ours - sch.get_last_lr()[0]) < 1e-12, (kind, warm, total, step)
                opt.step()
                sch.step()


def test_bench_reference_arm_prints_the_contract_line_and_product_arm_refuses_the_cpu():
    """`bench.py --impl reference` (the oracle port on the host cores, the one place besides tests / smoke that may run
    `oracle/`): one JSON line with the same metric / unit / config dict as the product arm plus `impl`, `cpu_baseline`
    and a zero-copy `e2e`.  The product arm itself must not fall back to the CPU."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, MODCR_REF_BUDGET_S="4", CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT
    assert d["config"] == bench.config_dict(dict(bench.WORKLOADS["cfg2"]), 1, False)
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "questions" in cb["sample"]
    assert d["e2e"] == dict(value=d["value"], unit=bench.UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    # the product arm: no device, no number
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
    assert not [l for l in out.stdout.splitlines() if l.startswith("{")]
