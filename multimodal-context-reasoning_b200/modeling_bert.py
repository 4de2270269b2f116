"""Oscar image-text encoder with prefix-tuned self-attention -- module surface of the reference's
``modeling/modeling_bert.py`` (:25-291) / ``modeling/modeling_transfomres.py`` (:471-727) and of the
vendored BERT building blocks (a_transformers/modeling_bert.py:166-234, 237-265, 362-451, 634-646).

Same class names, constructor and ``forward`` signatures, return tuples and ``state_dict`` keys; the
arithmetic runs in the sm_100a kernels of ``libmodcr_b200.so`` (no PyTorch math on the hot path).
The submodules are parameter containers with the reference's names; the model-level ``forward``
executes the whole encoder as one autograd node (``functions.EncoderFn``).
"""
import torch
from torch import nn

from . import engine as E
from .functions import DropoutFn, EncoderFn, EncoderRun, MlpHiddenFn, ScorerFn
from .modeling_utils import (BertConfig, CaptionPreTrainedModel, ImgPreTrainedModel, cfg_get, check_no_dropout,  # noqa: F401
                             engine_cfg)

BertLayerNorm = nn.LayerNorm


class AttentionsNotMaterialised:
    """Stands where the reference returns the tuple of dense ``[N,h,S,K]`` probability tensors
    (``output_attentions=True``, run_PMR_ModCR.py:726).  The fused kernel never writes them to HBM;
    what the path consumes -- the head- and layer-summed text->region block of the last three
    layers (v10:982) -- is available as ``.align_block``."""

    def __init__(self, align_block=None):
        self.align_block = align_block

    def __getitem__(self, i):
        raise NotImplementedError("dense attention probabilities are not materialised by the fused prefix-attention "
                                  "kernel; use .align_block (sum over the last 3 layers and heads of probs[:, :T, T:])")

    def __len__(self):
        return 0


# ------------------------------------------------------------------------------------------------
# parameter containers (names = reference state_dict keys)
class BertEmbeddings(nn.Module):
    """a_transformers/modeling_bert.py:166-234."""

    def __init__(self, config):
        super().__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size,
                                            padding_idx=cfg_get(config, "pad_token_id", 0))
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.register_buffer("position_ids", torch.arange(config.max_position_embeddings).expand((1, -1)))


class CaptionBertSelfAttention(nn.Module):
    """modeling/modeling_bert.py:25-75 -- ``forward(hidden_states, attention_mask, head_mask=None,
    history_state=None)``; executed inside the encoder node (fused QKV GEMM + prefix attention)."""

    def __init__(self, config):
        super().__init__()
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = config.hidden_size // config.num_attention_heads
        self.all_head_size = config.hidden_size
        self.query = nn.Linear(config.hidden_size, self.all_head_size)
        self.key = nn.Linear(config.hidden_size, self.all_head_size)
        self.value = nn.Linear(config.hidden_size, self.all_head_size)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)


class BertSelfOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


class CaptionBertAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.self = CaptionBertSelfAttention(config)
        self.output = BertSelfOutput(config)


class BertIntermediate(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)


class BertOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


class CaptionBertLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.attention = CaptionBertAttention(config)
        self.intermediate = BertIntermediate(config)
        self.output = BertOutput(config)


class CaptionBertEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.output_attentions = cfg_get(config, "output_attentions", False)
        self.output_hidden_states = cfg_get(config, "output_hidden_states", False)
        self.layer = nn.ModuleList([CaptionBertLayer(config) for _ in range(config.num_hidden_layers)])


class BertPooler(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.activation = nn.Tanh()


class _EncoderModelBase(CaptionPreTrainedModel):
    """Shared construction / plumbing of BertImgModel and SeqBertImgModel."""

    def _build(self, config):
        self.embeddings = BertEmbeddings(config)
        self.encoder = CaptionBertEncoder(config)
        self.pooler = BertPooler(config)
        self.img_dim = config.img_feature_dim
        self.img_feature_type = cfg_get(config, "img_feature_type", "faster_r-cnn")
        self.use_img_layernorm = cfg_get(config, "use_img_layernorm", None)
        if self.img_feature_type in ("dis_code", "dis_code_t", "dis_code_scale"):
            raise NotImplementedError("img_feature_type %r is outside the ModCR hot path" % self.img_feature_type)
        self.img_embedding = nn.Linear(self.img_dim, config.hidden_size, bias=True)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        if self.use_img_layernorm:
            self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=cfg_get(config, "img_layer_norm_eps", 1e-12))
        self._param_store = None

    def _store(self):
        named = dict(self.named_parameters())
        ps = self._param_store
        if ps is None or ps.named.keys() != named.keys() or any(ps.named[k] is not named[k] for k in named):
            ps = E.ParamStore(named)
            self._param_store = ps
        for k, p in named.items():
            if p.dtype != torch.float32 or not p.is_cuda:
                raise RuntimeError("modcr_b200: parameter %s must be a float32 CUDA tensor (got %s on %s); master "
                                   "weights stay fp32, the kernels keep their own bf16 operand copies"
                                   % (k, p.dtype, p.device))
        return ps, list(named.keys()), list(named.values())

    def _resize_token_embeddings(self, new_num_tokens):
        return self.resize_token_embeddings(new_num_tokens)


class BertImgModel(_EncoderModelBase):
    """Expand from BertModel to handle image region features as input
    (modeling/modeling_transfomres.py:564-694; twin modeling/modeling_bert.py:158-291)."""

    def __init__(self, config):
        super().__init__(config)
        self._build(config)
        self.init_weights()

    def forward(self, input_ids, token_type_ids=None, attention_mask=None, position_ids=None, head_mask=None,
                img_feats=None, encoder_history_states=None):
        """Returns ``(sequence_output [N,S,H], pooled_output [N,H])`` (+ an ``AttentionsNotMaterialised``
        placeholder when ``config.output_attentions``), bf16.

        ``encoder_history_states``: list (one per layer) of ``[N,P,H]`` prefix hidden states; layer i
        attends over K/V projected from ``[history_i ; hidden]`` (modeling/modeling_bert.py:36-44).
        Unlike the reference (assert at modeling_transfomres.py:661-662) ``img_feats`` and a prefix may be
        given together; a 2-D ``attention_mask`` must cover ``P+S`` keys, or ``S`` keys in
        which case the prefix keys are always visible."""
        if head_mask is not None:
            raise NotImplementedError("head_mask is always None on the ModCR path (modeling_transfomres.py:656-657)")
        N, T = input_ids.shape
        R = 0 if img_feats is None else img_feats.shape[1]
        S = T + R
        hist = list(encoder_history_states) if encoder_history_states else []
        P = hist[0].shape[1] if hist else 0
        if hist and len(hist) != self.config.num_hidden_layers:
            raise ValueError("encoder_history_states needs one [N,P,H] tensor per layer")
        dev = input_ids.device
        if attention_mask is None:
            attention_mask = torch.ones((N, P + S), dtype=torch.float32, device=dev)
        if attention_mask.dim() != 2:
            raise NotImplementedError("only 2-D key-padding masks are implemented for BertImgModel "
                                      "(3-D masks are built internally by SeqBertImgModel)")
        if attention_mask.shape[1] == S and P > 0:
            attention_mask = torch.cat([torch.ones((N, P), dtype=attention_mask.dtype, device=dev), attention_mask], -1)
        if attention_mask.shape[1] != P + S:
            raise ValueError("attention_mask has %d keys, expected P+S = %d" % (attention_mask.shape[1], P + S))
        kbias = E.make_key_bias(attention_mask)
        ps, names, params = self._store()
        run = EncoderRun(ps, names, engine_cfg(self.config), input_ids, token_type_ids, position_ids, img_feats, kbias, P)
        run.drop = E.make_drop(self.training, cfg_get(self.config, "hidden_dropout_prob", 0.0),
                               cfg_get(self.config, "attention_probs_dropout_prob", 0.0), dev, 1 << 20)
        seq_out, pooled, _, _ = EncoderFn.apply(run, len(hist), *hist, *params)
        outputs = (seq_out, pooled)
        if cfg_get(self.config, "output_attentions", False):
            outputs = outputs + (AttentionsNotMaterialised(),)
        return outputs


# ------------------------------------------------------------------------------------------------
# multiple-choice scorer heads of this module (modeling/modeling_bert.py:416-421, 492-568, 574-712)
def instance_bce_with_logits(logits, labels, reduction="mean"):
    """modeling/modeling_bert.py:416-421 (tiny, host-side glue kept in torch)."""
    assert logits.dim() == 2
    loss = nn.functional.binary_cross_entropy_with_logits(logits, labels, reduction=reduction)
    if reduction == "mean":
        loss = loss * labels.size(1)
    return loss


def _make_classifier(config, in_features, num_labels, kind, use_img_layernorm):
    """Classifier variants of modeling/modeling_bert.py:511-530 (same nn.Sequential layout => same state_dict keys)."""
    if kind == "linear":
        return nn.Linear(in_features, num_labels)
    if kind != "mlp":
        raise ValueError("config.classifier must be 'linear' or 'mlp' (got %r)" % (kind,))
    hid = config.hidden_size * cfg_get(config, "cls_hidden_scale", 2)
    if use_img_layernorm:
        return nn.Sequential(nn.Linear(in_features, hid), nn.ReLU(),
                             BertLayerNorm(hid, eps=config.layer_norm_eps), nn.Linear(hid, num_labels))
    return nn.Sequential(nn.Linear(in_features, hid), nn.ReLU(), nn.Linear(hid, num_labels))


def _classifier_hidden(model, x):
    """Runs everything of ``model.classifier`` but its last Linear (which the fused scorer kernel applies together
    with the loss); returns ``(hidden, last_linear)``."""
    cls = model.classifier
    if isinstance(cls, nn.Linear):
        return x, cls
    lin0, last = cls[0], cls[-1]
    ln = cls[2] if len(cls) == 4 else None
    if lin0.weight.dtype != torch.float32 or not lin0.weight.is_cuda:
        raise RuntimeError("modcr_b200: classifier weights must be float32 CUDA tensors (no CPU fallback)")
    named = {"classifier.0.weight": lin0.weight}
    ps = model._cls_store
    if ps is None or ps.named["classifier.0.weight"] is not lin0.weight:
        ps = E.ParamStore(named)
        model._cls_store = ps
    z = MlpHiddenFn.apply(x, lin0.weight, lin0.bias, None if ln is None else ln.weight, None if ln is None else ln.bias,
                          ps.bf16("classifier.0.weight"), 0.0 if ln is None else float(ln.eps))
    return z, last


class ImageBertForMultipleChoice(CaptionPreTrainedModel):
    """modeling/modeling_bert.py:492-568: choice-flattened classifier
    ``pooled.view(-1, num_choice*H) -> Linear -> CE`` fused into one kernel."""

    def __init__(self, config):
        super().__init__(config)
        self.loss_type = cfg_get(config, "loss_type", "sfmx")
        self.num_choice = cfg_get(config, "num_choice", 4)
        self.num_labels = config.num_labels
        self.bert = BertImgModel(config)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.classifier = _make_classifier(config, self.num_choice * config.hidden_size, config.num_labels,
                                           cfg_get(config, "classifier", "linear"),
                                           cfg_get(config, "use_img_layernorm", None))
        self._cls_store = None
        # NB the reference forgets the call parentheses on init_weights (:531): the head keeps torch's default init

    def _pooled_dropout(self, pooled):
        """``self.dropout(pooled_output)`` (modeling/modeling_bert.py:553, 697) in training mode."""
        p = float(self.dropout.p)
        if not self.training or p <= 0.0:
            return pooled
        return DropoutFn.apply(pooled, E.next_seed(pooled.device), 5 << 20, p)

    def forward(self, input_ids, token_type_ids=None, attention_mask=None, labels=None, position_ids=None,
                head_mask=None, img_feats=None):
        nc = self.num_choice
        flat = lambda t: None if t is None else t.reshape(-1, *t.shape[2:])  # noqa: E731
        outputs = self.bert(flat(input_ids), position_ids=flat(position_ids), token_type_ids=flat(token_type_ids),
                            attention_mask=flat(attention_mask), head_mask=head_mask, img_feats=flat(img_feats))
        pooled = self._pooled_dropout(outputs[1])
        B = pooled.shape[0] // nc
        x = pooled.reshape(B, nc * pooled.shape[1])
        if labels is None:
            labels_ = torch.zeros((B,), dtype=torch.int64, device=x.device)
        else:
            labels_ = labels.view(-1)
        x, last = _classifier_hidden(self, x)
        if self.loss_type == "bce" and labels is not None:
            # instance_bce_with_logits(logits, labels.view(-1, num_labels)) (:416-421, :562)
            soft = labels.reshape(-1, self.num_labels).to(torch.float32).contiguous()
            loss, logits = ScorerFn.apply(x, last.weight, last.bias, None, soft, self.num_labels, nc, False, 1)
        else:
            loss, logits = ScorerFn.apply(x, last.weight, last.bias, labels_.long().contiguous(), None,
                                          self.num_labels, nc, False)
        out = (logits,) + outputs[2:]
        if labels is not None:
            out = (loss,) + out
        return out


class OscarForMultipleChoice(CaptionPreTrainedModel):
    """modeling/modeling_bert.py:574-712: every (question, choice) pair is scored on its own --
    ``[B,C,T]`` / ``[B,C,R,D]`` inputs flattened to ``[B*C,...]`` (:676-681), ``BertImgModel`` pooled output ->
    ``Linear(H -> 2)`` -> cross-entropy over the flattened rows (``loss_type`` 'bxe' and the default; :699-710),
    classifier + CE fused in one kernel.  Returns ``(loss), logits [B*C,2]``."""

    def __init__(self, config):
        super().__init__(config)
        self.loss_type = cfg_get(config, "loss_type", "bxe")
        self.bert = BertImgModel(config)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        kind = cfg_get(config, "classifier", None)
        if kind in (None, "linear"):
            self.classifier = nn.Linear(config.hidden_size, 2 if kind == "linear" else config.num_labels)
        else:   # 'mlp' (:646-660): Linear(H -> H*scale) -> ReLU [-> LayerNorm] -> Linear(-> 2)
            self.classifier = _make_classifier(config, config.hidden_size, 2, kind, None)
        self._cls_store = None
        self.apply(self._init_weights)

    def forward(self, input_ids, token_type_ids=None, attention_mask=None, labels=None, position_ids=None,
                head_mask=None, img_feats=None):
        flat = lambda t: None if t is None else t.reshape(-1, *t.shape[2:])  # noqa: E731
        outputs = self.bert(flat(input_ids), position_ids=flat(position_ids), token_type_ids=flat(token_type_ids),
                            attention_mask=flat(attention_mask), head_mask=head_mask, img_feats=flat(img_feats))
        pooled = ImageBertForMultipleChoice._pooled_dropout(self, outputs[1])
        pooled, last = _classifier_hidden(self, pooled)
        rows, C = pooled.shape[0], last.weight.shape[0]
        if self.loss_type == "bce" and labels is not None:
            # instance_bce_with_logits(logits, labels.view(-1, num_labels)) (:700-701)
            soft = labels.reshape(-1, C).to(torch.float32).contiguous()
            loss, logits = ScorerFn.apply(pooled, last.weight, last.bias, None, soft, C, 1, False, 1)
        else:   # 'bxe' and the default: CrossEntropyLoss over the flattened rows (:702-710)
            labels_ = (torch.zeros((rows,), dtype=torch.int64, device=pooled.device) if labels is None
                       else labels.reshape(-1).long().contiguous())
            loss, logits = ScorerFn.apply(pooled, last.weight, last.bias, labels_, None, C, 1, False)
        out = (logits,) + outputs[2:]
        if labels is not None:
            out = (loss,) + out
        return out
