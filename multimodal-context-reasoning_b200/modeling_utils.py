"""Base classes, configuration and checkpoint surface of the hot-path modules.

Mirrors what the reference hot path inherits from ``modeling/modeling_utils.py``
(``CaptionPreTrainedModel`` :19-1043, ``ImgPreTrainedModel.from_pretrained`` :1134-1323 with the
``gamma/beta -> weight/bias`` rename :1256-1269) and from the vendored HF ``BertPreTrainedModel``
(a_transformers/modeling_bert.py:718-743, ``_init_weights``): same ``state_dict`` key names and
shapes, ``from_pretrained(path, config=)``, ``save_pretrained``, ``init_weights``,
``resize_token_embeddings``.  The generation helpers of the reference file are out of scope.
"""
import copy
import json
import os

import torch
from torch import nn

WEIGHTS_NAME = "pytorch_model.bin"
CONFIG_NAME = "config.json"


class BertConfig:
    """Attribute bag compatible with the HF ``config.json`` the reference loads
    (a_transformers/configuration_bert.py:119-156 defaults) plus the ModCR attributes the run
    scripts patch on (run_PMR_ModCR.py:717-726)."""

    def __init__(self, **kw):
        d = dict(vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                 intermediate_size=3072, hidden_act="gelu", hidden_dropout_prob=0.1,
                 attention_probs_dropout_prob=0.1, max_position_embeddings=512, type_vocab_size=2,
                 initializer_range=0.02, layer_norm_eps=1e-12, pad_token_id=0,
                 img_feature_dim=2054, img_feature_type="faster_r-cnn", use_img_layernorm=1,
                 img_layer_norm_eps=1e-12, output_attentions=True, output_hidden_states=False,
                 max_hypo=-1, add_residual=False, add_local_residual=False, num_labels=2, loss_type="sfmx",
                 num_choice=4, classifier="linear", cls_hidden_scale=2)
        d.update(kw)
        for k, v in d.items():
            setattr(self, k, v)

    @classmethod
    def from_json_file(cls, path):
        with open(path) as f:
            return cls(**json.load(f))

    @classmethod
    def from_pretrained(cls, path, **kw):
        p = os.path.join(path, CONFIG_NAME) if os.path.isdir(path) else path
        c = cls.from_json_file(p)
        for k, v in kw.items():
            setattr(c, k, v)
        return c

    def to_dict(self):
        return copy.deepcopy(self.__dict__)

    def save_pretrained(self, path):
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, CONFIG_NAME), "w") as f:
            json.dump(self.to_dict(), f, indent=2, sort_keys=True)


def cfg_get(config, name, default=None):
    return getattr(config, name, default)


def engine_cfg(config):
    """The few numbers the kernels need."""
    H = config.hidden_size
    heads = config.num_attention_heads
    if H % heads != 0 or H // heads != 64:
        raise NotImplementedError("modcr_b200 kernels are built for head_dim 64 (got hidden %d / heads %d)" % (H, heads))
    act = cfg_get(config, "hidden_act", "gelu")
    if act != "gelu":
        raise NotImplementedError("only the erf GELU of the reference config is implemented (got %r)" % (act,))
    return dict(H=H, heads=heads, L=config.num_hidden_layers, I=config.intermediate_size,
                eps=float(config.layer_norm_eps), img_eps=float(cfg_get(config, "img_layer_norm_eps", 1e-12)),
                use_img_layernorm=int(bool(cfg_get(config, "use_img_layernorm", 0))),
                pad_token_id=int(cfg_get(config, "pad_token_id", 0) or 0), n_cls_layers=3,
                add_residual=bool(cfg_get(config, "add_residual", False)),
                add_local_residual=bool(cfg_get(config, "add_local_residual", False)))


def check_no_dropout(module, config):
    """Kept for callers that want the deterministic (parity / benchmark) arithmetic guaranteed: raises when the module
    would apply training-mode dropout.  The kernels implement dropout (counter-based masks, ``engine.Drop``)."""
    if module.training and (float(cfg_get(config, "hidden_dropout_prob", 0.0)) > 0.0 or
                            float(cfg_get(config, "attention_probs_dropout_prob", 0.0)) > 0.0):
        raise RuntimeError("modcr_b200: module is in training mode with dropout (hidden %.2f / attention %.2f)" % (
            config.hidden_dropout_prob, config.attention_probs_dropout_prob))


class CaptionPreTrainedModel(nn.Module):
    """Base of every hot-path model: config, weight init, checkpoint I/O."""
    config_class = BertConfig
    base_model_prefix = "bert"

    def __init__(self, config, *inputs, **kwargs):
        super().__init__()
        self.config = config

    # ---- a_transformers/modeling_bert.py:729-743
    def _init_weights(self, module):
        std = cfg_get(self.config, "initializer_range", 0.02)
        if isinstance(module, nn.Linear):
            module.weight.data.normal_(mean=0.0, std=std)
            if module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.Embedding):
            module.weight.data.normal_(mean=0.0, std=std)
            if module.padding_idx is not None:
                module.weight.data[module.padding_idx].zero_()
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)

    def init_weights(self):
        self.apply(self._init_weights)

    # ---- token-embedding resize (run_PMR_ModCR.py:730 adds 45 "<|detN|>" tokens)
    def _base_model(self):
        """The encoder that owns the embeddings: ``self`` for BertImgModel / SeqBertImgModel, ``self.bert`` for the
        *ForMultipleChoice wrappers (HF ``base_model_prefix`` convention)."""
        return getattr(self, self.base_model_prefix, self)

    def get_input_embeddings(self):
        return self._base_model().embeddings.word_embeddings

    def resize_token_embeddings(self, new_num_tokens=None):
        base = self._base_model()
        old = base.embeddings.word_embeddings
        if new_num_tokens is None or new_num_tokens == old.num_embeddings:
            return old
        new = nn.Embedding(new_num_tokens, old.embedding_dim, padding_idx=old.padding_idx)
        new.to(old.weight.device, dtype=old.weight.dtype)
        self._init_weights(new)
        n = min(old.num_embeddings, new_num_tokens)
        new.weight.data[:n] = old.weight.data[:n]
        base.embeddings.word_embeddings = new
        self.config.vocab_size = new_num_tokens
        base.config.vocab_size = new_num_tokens
        base._param_store = None
        return new

    # ---- checkpoints
    def save_pretrained(self, save_directory):
        os.makedirs(save_directory, exist_ok=True)
        if hasattr(self.config, "save_pretrained"):
            self.config.save_pretrained(save_directory)
        torch.save(self.state_dict(), os.path.join(save_directory, WEIGHTS_NAME))

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, *model_args, config=None, state_dict=None, **kwargs):
        """``path`` is a directory with ``config.json`` + ``pytorch_model.bin`` or a weights file.
        Keys are matched with and without the ``bert.`` prefix; old ``gamma``/``beta`` LayerNorm
        names are renamed (modeling/modeling_utils.py:1256-1269); missing / unexpected keys are
        tolerated and reported on ``model.load_info`` exactly like the reference's loader.  A key whose SHAPE differs
        (e.g. a vocabulary-resized ``word_embeddings``) raises, like HF's loader (``size mismatch``) -- it is never
        silently left at its random initialisation.  A prefix-less encoder checkpoint loads into a wrapper model
        (``OscarForMultipleChoice.bert.*``) and a ``bert.``-prefixed one into a bare encoder."""
        path = pretrained_model_name_or_path
        if config is None:
            config = BertConfig.from_pretrained(path)
        model = cls(config, *model_args, **kwargs)
        if state_dict is None and path is not None:
            wfile = os.path.join(path, WEIGHTS_NAME) if os.path.isdir(path) else path
            if os.path.isfile(wfile):
                state_dict = torch.load(wfile, map_location="cpu")
        missing, unexpected = [], []
        if state_dict is not None:
            renamed = {}
            for k, v in state_dict.items():
                nk = k.replace("gamma", "weight") if "gamma" in k else k
                nk = nk.replace("beta", "bias") if "beta" in nk else nk
                renamed[nk] = v
            own = model.state_dict()
            load, mismatched = {}, []
            pfx = cls.base_model_prefix + "."
            for k, v in renamed.items():
                if k in own:
                    kk = k
                elif k.startswith(pfx) and k[len(pfx):] in own:      # "bert.x" checkpoint -> bare encoder
                    kk = k[len(pfx):]
                elif pfx + k in own:                                    # bare-encoder checkpoint -> wrapper model
                    kk = pfx + k
                else:
                    unexpected.append(k)
                    continue
                if tuple(own[kk].shape) != tuple(v.shape):
                    mismatched.append("%s: checkpoint %s vs model %s" % (k, tuple(v.shape), tuple(own[kk].shape)))
                    continue
                load[kk] = v
            if mismatched:
                raise RuntimeError("from_pretrained: size mismatch for " + "; ".join(mismatched))
            missing = [k for k in own if k not in load]
            model.load_state_dict(load, strict=False)
        model.load_info = dict(missing_keys=missing, unexpected_keys=unexpected)
        model.eval()
        return model


class ImgPreTrainedModel(CaptionPreTrainedModel):
    """Import-surface twin of modeling/modeling_utils.py:1126-1323 (same loader semantics)."""
