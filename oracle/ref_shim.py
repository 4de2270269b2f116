"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference classes.

Imports the unmodified reference model code from /root/reference (plus the
vendored a_transformers.zip) on CPU under a small compatibility shim, so that
the CPU restatement in ``oracle/restate.py`` can be validated against it and
golden vectors can be generated (``oracle/make_golden.py``).

/root/reference exists only in the build container: nothing in the product
package, ``-m gpu`` tests, ``smoke()`` or ``bench.py`` imports this module.

Shim recipe follows SURVEY.md section 8(c).
"""
import os
import sys
import types
import zipfile

REF_ROOT = os.environ.get("MODCR_REFERENCE_ROOT", "/root/reference")
_A_TF_DIR = os.environ.get("MODCR_A_TF_DIR", "/tmp/modcr_a_tf")


def available() -> bool:
    return os.path.isdir(REF_ROOT) and os.path.isfile(os.path.join(REF_ROOT, "a_transformers.zip"))


_loaded = None


def load():
    """Returns a namespace with the reference modules (v10, transfomres, bert, a_bert)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    import torch
    import transformers
    import transformers.modeling_utils as tmu
    import transformers.pytorch_utils as tpu

    if not os.path.isdir(os.path.join(_A_TF_DIR, "a_transformers")):
        zipfile.ZipFile(os.path.join(REF_ROOT, "a_transformers.zip")).extractall(_A_TF_DIR)
    for p in (_A_TF_DIR, REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)

    # symbols that moved / vanished between transformers 4.x and 5.x
    for name in ("apply_chunking_to_forward", "prune_linear_layer", "Conv1D",
                 "find_pruneable_heads_and_indices"):
        if not hasattr(tmu, name):
            if hasattr(tpu, name):
                setattr(tmu, name, getattr(tpu, name))
            else:
                setattr(tmu, name, lambda *a, **k: (_ for _ in ()).throw(NotImplementedError(name)))
    for name in ("prune_conv1d_layer", "SequenceSummary"):
        if not hasattr(tmu, name):
            setattr(tmu, name, type(name, (), {}) if name[0].isupper()
                    else (lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())))
    for name, val in (("TF_WEIGHTS_NAME", "model.ckpt"), ("WEIGHTS_NAME", "pytorch_model.bin"),
                      ("TF2_WEIGHTS_NAME", "tf_model.h5"), ("CONFIG_NAME", "config.json")):
        if not hasattr(transformers, name):
            setattr(transformers, name, val)
        if not hasattr(tmu, name):
            setattr(tmu, name, val)
    if not hasattr(transformers, "BeamScorer"):
        transformers.BeamScorer = type("BeamScorer", (), {})
    for modname, names in (
        ("transformers.generation_stopping_criteria",
         ["MaxLengthCriteria", "MaxTimeCriteria", "StoppingCriteriaList", "validate_stopping_criteria"]),
        ("transformers.generation_logits_process",
         ["EncoderNoRepeatNGramLogitsProcessor", "ForcedBOSTokenLogitsProcessor",
          "ForcedEOSTokenLogitsProcessor", "HammingDiversityLogitsProcessor",
          "InfNanRemoveLogitsProcessor", "LogitsProcessorList", "MinLengthLogitsProcessor",
          "NoBadWordsLogitsProcessor", "NoRepeatNGramLogitsProcessor",
          "PrefixConstrainedLogitsProcessor", "RepetitionPenaltyLogitsProcessor",
          "TemperatureLogitsWarper", "TopKLogitsWarper", "TopPLogitsWarper"]),
    ):
        if modname not in sys.modules:
            m = types.ModuleType(modname)
            for n in names:
                setattr(m, n, type(n, (), {}))
            sys.modules[modname] = m
    import transformers.file_utils as tfu
    if not hasattr(tfu, "cached_path"):
        tfu.cached_path = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("cached_path"))

    import importlib
    try:
        a_bert = importlib.import_module("a_transformers.modeling_bert")
    except Exception:
        # a_transformers/file_utils etc. may want more 4.x names; surface the error
        raise
    # transformers-5 init_weights needs post_init state -> plain apply
    # (OscarForMultipleChoice calls the transformers-1.x form ``self.apply(self.init_weights)``, modeling_bert.py:671)
    a_bert.BertPreTrainedModel.init_weights = (
        lambda self, module=None: self.apply(self._init_weights) if module is None else self._init_weights(module))

    # CPU only: the chunk-mean loop calls .cuda(device) (v10:71,74)
    torch.Tensor.cuda = lambda self, *a, **k: self

    # utils/cbs.py (captioning beam search, unused on this path) imports anytree
    if "anytree" not in sys.modules:
        try:
            importlib.import_module("anytree")
        except ImportError:
            at = types.ModuleType("anytree")
            at.AnyNode = type("AnyNode", (), {})
            at.Node = type("Node", (), {})
            at.RenderTree = type("RenderTree", (), {})
            sys.modules["anytree"] = at

    v10 = importlib.import_module("modeling.modeling_vcr_chunkalign_v10")
    tfm = importlib.import_module("modeling.modeling_transfomres")
    mbert = importlib.import_module("modeling.modeling_bert")
    ns = types.SimpleNamespace(v10=v10, tfm=tfm, mbert=mbert, a_bert=a_bert)
    _loaded = ns
    return ns


def load_ensemble():
    """``modeling.modeling_ensemble`` of the reference (``Abstract_Specific``, ``PrefixMLP``).  Its RoBERTa import comes
    from ``local_transformers.zip``, which the reference does not ship: that one name is stubbed (the tests pass their
    own language-model stand-in to the constructor, the class never touches ``RobertaModel`` itself)."""
    ns = load()
    if getattr(ns, "ens", None) is not None:
        return ns
    import importlib
    import transformers
    chain = "local_transformers.adapter_transformers.models.roberta"
    parts = chain.split(".")
    for i in range(1, len(parts) + 1):
        name = ".".join(parts[:i])
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules[chain].RobertaModel = type("RobertaModel", (), {})
    for n in ("AutoTokenizer", "AutoModelForSeq2SeqLM", "LogitsProcessorList", "MinLengthLogitsProcessor"):
        try:
            getattr(transformers, n)
        except Exception:
            setattr(transformers, n, type(n, (), {}))
    ns.ens = importlib.import_module("modeling.modeling_ensemble")
    return ns


def make_config(ns, **over):
    """BertConfig carrying the attributes the run scripts monkey-patch on
    (run_PMR_ModCR.py:717-726,736-748) plus transformers-5 gaps."""
    from a_transformers.configuration_bert import BertConfig
    cfg = BertConfig()
    base = dict(
        vocab_size=30567, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
        intermediate_size=3072, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
        img_feature_dim=2054, img_feature_type="faster_r-cnn", use_img_layernorm=1,
        img_layer_norm_eps=1e-12, loss_type="sfmx", output_attentions=True,
        output_hidden_states=False, max_hypo=128, add_residual=False, add_local_residual=False,
        is_decoder=False, add_cross_attention=False, chunk_size_feed_forward=0,
        num_labels=2, num_choice=4,
    )
    base.update(over)
    for k, v in base.items():
        setattr(cfg, k, v)
    return cfg
