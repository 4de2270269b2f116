"""Chunk-aware encoder, chunk-to-region alignment heads and the multiple-choice scorer -- module
surface of the reference's ``modeling/modeling_vcr_chunkalign_v10.py`` (:45-373, :801-1165).

Same class names, constructor / ``forward`` signatures, return tuples and ``state_dict`` keys.  All
arithmetic runs in the sm_100a kernels behind ``libmodcr_b200.so``; the Python below only assembles
kernel programs (``functions``) and the tiny host-side glue that the reference also does on the host.
"""
import torch
from torch import nn

from . import engine as E
from .functions import (AlignLossFn, ClsHeadRun, ClsLayer2HeadFn, ClsLyxHeadFn, EncoderFn, EncoderRun, ScorerFn,
                        cls_lyx_head_forward)
from .modeling_bert import (AttentionsNotMaterialised, BertIntermediate, BertOutput, CaptionBertAttention,
                            _EncoderModelBase)
from .modeling_utils import cfg_get, engine_cfg

NEG = -10000.0


def binary_to_mp(logit, num_labels=4):
    """v10:363-373 -- softmax over the 2 classes, keep class 1, group by question."""
    sm = torch.softmax(logit.float(), dim=1)
    return sm[:, 1].reshape(-1, num_labels)


class SeqBertImgModel(_EncoderModelBase):
    """Chunk-aware Oscar encoder (v10:235-360) with the 3-phase mask schedule and chunk-mean queries of
    its ``CaptionBertEncoder`` / ``CaptionBertSelfAttention`` (v10:55-107, 171-232)."""

    def __init__(self, config):
        super().__init__(config)
        self._build(config)
        self.init_weights()
        self.max_hypo = cfg_get(config, "max_hypo", -1)
        self.edge_dense = nn.Embedding(1, config.hidden_size)   # unused by the reference too (v10:259)

    def forward(self, input_ids, token_type_ids=None, attention_mask=None, input_mask=None, position_ids=None,
                head_mask=None, img_feats=None, img_mask=None, encoder_history_states=None, offsets=None,
                gather_index=None, _want_align_block=True, _align_rows=None):
        """``attention_mask`` is the 0/1 chunk mask ``[N,T,T]`` (or its int32 bit-packed form ``[N,T,ceil(T/32)]``
        from ``packer.BatchPacker``); ``input_mask`` ``[N,T+R]`` the key padding
        mask; ``gather_index`` the ragged per-sequence chunk ids (or an already packed int32 ``[N,T]``
        tensor with -1 outside the chunked span).  Returns ``((sequence_output, pooled_output, attentions),
        chunk_hidden_states)`` like v10:360; ``attentions.align_block`` is the fp32 ``[N,T,R]`` sum over the
        last three layers and all heads of ``probs[:, :T, T:]``.  ``_align_rows`` (int64 ``[N,T]``, the batch's
        ``align_pos``): only those rows of the block are computed (the alignment loss reads no other, v10:1075-1076);
        the rest of the block is zero."""
        if head_mask is not None:
            raise NotImplementedError("head_mask is always None on the ModCR path")
        if encoder_history_states:
            raise NotImplementedError("seq_enc never receives history states on the ModCR path (v10:1042-1046)")
        if img_feats is None or attention_mask is None or input_mask is None or gather_index is None:
            raise ValueError("SeqBertImgModel needs img_feats, the chunk attention_mask, input_mask and gather_index")
        N, T = input_ids.shape
        dev = input_ids.device
        kbias = E.make_key_bias(input_mask)
        if attention_mask.dtype == torch.int32:     # already packed by the device-side collate (packer.BatchPacker)
            cwords = (T + 31) // 32
            if tuple(attention_mask.shape) != (N, T, cwords):
                raise ValueError("packed chunk mask must be int32 [N,T,ceil(T/32)]")
            cbits = attention_mask.contiguous()
        else:
            cbits, cwords = E.make_chunk_bits(attention_mask)
        if torch.is_tensor(gather_index):
            chunk_id = gather_index.to(device=dev, dtype=torch.int32).contiguous()
            max_chunks = T
        else:
            chunk_id, max_chunks = E.pack_chunk_ids(gather_index, T, dev)
        ps, names, params = self._store()
        run = EncoderRun(ps, names, engine_cfg(self.config), input_ids, token_type_ids, position_ids, img_feats, kbias, 0,
                         is_seq=True, cbits=cbits, cwords=cwords, chunk_id=chunk_id, max_chunks=max_chunks,
                         want_block=_want_align_block,
                         align_rows=None if _align_rows is None else _align_rows.to(torch.int64).contiguous())
        run.drop = E.make_drop(self.training, cfg_get(self.config, "hidden_dropout_prob", 0.0),
                               cfg_get(self.config, "attention_probs_dropout_prob", 0.0), dev, 2 << 20)
        run.keep_attn = bool(getattr(self, "materialise_attentions", False))
        seq_out, pooled, chunk_hidden, block = EncoderFn.apply(run, 0, *params)
        att = AttentionsNotMaterialised(block if block.numel() else None, layers=run.attn_debug)
        return (seq_out, pooled, att), (chunk_hidden if chunk_hidden.numel() else None)


class ClsLayer2(nn.Module):
    """v10:801-837 (a BertLayer subclass: the inherited ``attention`` block exists in the state_dict but
    is never called).  Parameter container; executed by ``functions.ClsLayer2HeadFn``."""

    def __init__(self, config):
        super().__init__()
        self.attention = CaptionBertAttention(config)
        self.intermediate = BertIntermediate(config)
        self.output = BertOutput(config)
        self.cls_q_proj = nn.Linear(config.hidden_size, config.hidden_size)
        self.align_k_proj = nn.Linear(config.hidden_size, config.hidden_size)
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


class ChunkAlign_CLS_enc4_align(nn.Module):
    """Fully trainable alignment model (v10:1016-1165): global encoder (+ per-layer prefix), chunk-aware
    encoder, ``cls_ensemble``, 3 x ``ClsLayer2``, 2-way classifier + CE, ``binary_to_mp`` and the
    chunk-to-region alignment loss."""

    def __init__(self, global_enc, seq_enc, num_labels):
        super().__init__()
        self.global_enc = global_enc
        self.seq_enc = seq_enc
        H = self.global_enc.config.hidden_size
        self.cls_ensemble = nn.Linear(H + self.seq_enc.config.hidden_size, H)
        self.num_labels = num_labels
        self.cls_layer_num = 3
        self.cls_layer = nn.ModuleList([ClsLayer2(self.global_enc.config) for _ in range(self.cls_layer_num)])
        self.classifier = nn.Linear(H, 2)
        self.cls_loss_fct = nn.CrossEntropyLoss()
        self._head_store = None
        # True: the text->region block is computed on the aligned rows only (all the loss reads); False materialises the
        # whole [N,T,R] block (``last_align_block`` then equals the reference's `attention_weights` everywhere)
        self.sparse_align_block = True

    # ------------------------------------------------------------------ plumbing
    def _head_params(self):
        named = {k: p for k, p in self.named_parameters()
                 if not (k.startswith("global_enc.") or k.startswith("seq_enc.") or k.startswith("classifier."))}
        ps = self._head_store
        if ps is None or ps.named.keys() != named.keys() or any(ps.named[k] is not named[k] for k in named):
            ps = E.ParamStore(named)
            self._head_store = ps
        return ps, list(named.keys()), list(named.values())

    def _encode(self, input_ids, img_feat, input_mask, token_type_ids, position_ids, encoder_history_states, offsets,
                chunk_attention_mask, gather_index, want_block, align_rows=None):
        N, T = input_ids.shape
        g = self.global_enc(input_ids, img_feats=img_feat, attention_mask=input_mask, position_ids=position_ids,
                            token_type_ids=token_type_ids, head_mask=None,
                            encoder_history_states=encoder_history_states)
        (s_out, s_pool, att), chunk_hidden = self.seq_enc(
            input_ids, img_feats=img_feat, img_mask=input_mask[:, T:], input_mask=input_mask,
            attention_mask=chunk_attention_mask, position_ids=position_ids, token_type_ids=token_type_ids,
            head_mask=None, offsets=offsets, gather_index=gather_index, _want_align_block=want_block,
            _align_rows=align_rows)
        ps, names, params = self._head_params()
        cfg = engine_cfg(self.global_enc.config)
        cfg["n_cls_layers"] = self.cls_layer_num
        S = g[0].shape[1]
        # word mask (v10:1055-1057): (1 - input_mask[:, 1:T]) * -10000, tiled for the three sources
        wm = (1.0 - input_mask[:, 1:T].to(torch.float32)) * NEG
        wm = torch.cat((wm, wm, wm), -1).contiguous()
        run = ClsHeadRun(ps, names, cfg, N, T, S)
        run.drop = E.make_drop(self.training, cfg_get(self.global_enc.config, "hidden_dropout_prob", 0.0), 0.0,
                               input_ids.device, 3 << 20)
        cls = ClsLayer2HeadFn.apply(run, wm, g[0], s_out, chunk_hidden, g[1], s_pool, *params)
        # detached debug views (never keep the autograd graph of a finished step alive)
        self.last_intermediates = dict(global_output=g[0].detach(), global_CLS=g[1].detach(), seq_output=s_out.detach(),
                                       chunk_CLS=s_pool.detach(), chunk_hidden_states=chunk_hidden.detach(),
                                       CLS=cls.detach())
        return cls, att

    def forward(self, input_ids, img_feat, input_mask=None, label=None, token_type_ids=None, position_ids=None,
                head_mask=None, encoder_history_states=None, offsets=None, chunk_attention_mask=None,
                gather_index=None, align_pos=None, total_label=None):
        """Returns ``(loss_cls_0, matched_0, align_loss, correct, total_sum)`` (v10:1084).  ``correct`` and
        ``total_sum`` are 0-dim device tensors (the reference calls ``.item()`` inside forward, v10:1079-1082;
        here the host sync is left to the caller)."""
        want_block = total_label is not None
        cls, att = self._encode(input_ids, img_feat, input_mask, token_type_ids, position_ids, encoder_history_states,
                                offsets, chunk_attention_mask, gather_index, want_block,
                                align_pos if (want_block and self.sparse_align_block) else None)
        loss_cls_0, logits, mp, pred, matched_0 = ScorerFn.apply(
            cls, self.classifier.weight, self.classifier.bias, label.long().contiguous(), None, 2, self.num_labels, True)
        self.last_logits, self.last_logit_expl, self.last_pred = logits, mp, pred
        self.last_align_block = None if att.align_block is None else att.align_block.detach()
        align_loss, correct, total_sum = None, 0, 0
        if want_block:
            align_loss, stats = AlignLossFn.apply(att.align_block, align_pos, total_label)
            total_sum, correct = stats[0].long(), stats[1].long()
        return loss_cls_0, matched_0, align_loss, correct, total_sum

    @torch.no_grad()
    def evaluate(self, input_ids, img_feat, input_mask=None, label=None, token_type_ids=None, position_ids=None,
                 head_mask=None, encoder_history_states=None, offsets=None, chunk_attention_mask=None,
                 gather_index=None, align_pos=None, total_label=None):
        """v10:1086-1125: returns ``(matched_0, pre, logit_expl)``."""
        cls, _ = self._encode(input_ids, img_feat, input_mask, token_type_ids, position_ids, encoder_history_states,
                              offsets, chunk_attention_mask, gather_index, False)
        lab = label.long().contiguous() if label is not None else torch.zeros(
            (input_ids.shape[0],), dtype=torch.int64, device=input_ids.device)
        _, logits, mp, pred, matched_0 = ScorerFn.apply(cls, self.classifier.weight, self.classifier.bias, lab, None, 2,
                                                        self.num_labels, True)
        return matched_0, pred, mp


class cross_attention_lyx(nn.Module):
    """Parameter container of the reference's 8-head cross-attention (v10:692-797): ``k_proj``, ``v_proj``,
    ``q_proj``, ``out_proj``.  The one configuration the path uses (a single CLS query, no mask, ``tau=1``,
    ``neg_type=False``, no ``prior_score``) runs in ``cross_attn_q1_fwd`` (csrc/heads.cu)."""

    def __init__(self, embed_dim, num_heads, dropout=0.0, is_decoder=False, bias=True):
        super().__init__()
        self.embed_dim, self.num_heads, self.dropout, self.is_decoder = embed_dim, num_heads, dropout, is_decoder
        self.head_dim = embed_dim // num_heads
        if self.head_dim * num_heads != embed_dim:
            raise ValueError("embed_dim must be divisible by num_heads (got %d and %d)" % (embed_dim, num_heads))
        self.scaling = self.head_dim ** -0.5
        self.k_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.v_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.q_proj = nn.Linear(embed_dim, embed_dim, bias=bias)
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=bias)


class ClsLayer_lyx(nn.Module):
    """v10:838-870 (a BertLayer subclass: the inherited ``attention`` block, ``ensemble`` and ``dense`` exist in the
    state_dict but are never called).  Parameter container; executed by ``functions.cls_lyx_head_forward``."""

    def __init__(self, config):
        super().__init__()
        self.attention = CaptionBertAttention(config)
        self.intermediate = BertIntermediate(config)
        self.output = BertOutput(config)
        self.ensemble = nn.Linear(config.hidden_size * 2, 1)
        self.cross_attention = cross_attention_lyx(config.hidden_size, 8, dropout=0.1, is_decoder=True)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)


class ChunkAlign_CLS_enc4_align_ensemble(nn.Module):
    """Frozen-encoder alignment model used inside ``Abstract_Specific`` (v10:873-997): both encoders run without
    gradients, ``cls_ensemble_1`` + 2 x ``ClsLayer_lyx`` fuse [global tokens ; chunk tokens ; chunk hidden states]
    into ``CLS_ensem``; returns ``(CLS_ensem, align_loss, ([], None))``.  ``cls_layer``, ``classifier``,
    ``fusion_align`` and ``prior`` are part of the reference's state_dict but never called.

    The encoders are frozen (``no_grad``, like the reference); ``CLS_ensem`` carries gradients to ``cls_ensemble_1``
    and the ``ClsLayer_lyx`` stack (``functions.ClsLyxHeadFn``), which is what the mapping networks of
    ``Abstract_Specific`` train through (modeling/modeling_ensemble.py:466-490)."""

    def __init__(self, global_enc, seq_enc, num_labels):
        super().__init__()
        self.global_enc = global_enc
        self.seq_enc = seq_enc
        H = self.global_enc.config.hidden_size
        self.cls_ensemble_1 = nn.Linear(H + self.seq_enc.config.hidden_size, H)
        self.num_labels = num_labels
        self.cls_layer_num = 2
        self.cls_layer = nn.ModuleList([ClsLayer2(self.global_enc.config) for _ in range(self.cls_layer_num)])
        self.cls_layer_lyx = nn.ModuleList([ClsLayer_lyx(self.global_enc.config) for _ in range(self.cls_layer_num)])
        self.classifier = nn.Linear(H, 2)
        self.fusion_align = nn.Linear(H * 2, 1024)
        self.prior = nn.Linear(H, 1)
        self.cls_loss_fct = nn.CrossEntropyLoss()
        self._head_store = None

    def _head_params(self):
        named = {k: p for k, p in self.named_parameters()
                 if k.startswith("cls_ensemble_1.") or k.startswith("cls_layer_lyx.")}
        ps = self._head_store
        if ps is None or ps.named.keys() != named.keys() or any(ps.named[k] is not named[k] for k in named):
            ps = E.ParamStore(named)
            self._head_store = ps
        return ps, named

    def forward(self, input_ids, img_feat, input_mask=None, label=None, token_type_ids=None, position_ids=None,
                head_mask=None, encoder_history_states=None, offsets=None, chunk_attention_mask=None,
                gather_index=None, align_pos=None, total_label=None, abstract_hidden_states=None):
        ps, named = self._head_params()
        N, T = input_ids.shape
        want_block = total_label is not None
        with torch.no_grad():
            g = self.global_enc(input_ids, img_feats=img_feat, attention_mask=input_mask, position_ids=position_ids,
                                token_type_ids=token_type_ids, head_mask=None,
                                encoder_history_states=encoder_history_states)
            (s_out, s_pool, att), chunk_hidden = self.seq_enc(
                input_ids, img_feats=img_feat, img_mask=input_mask[:, T:], input_mask=input_mask,
                attention_mask=chunk_attention_mask, position_ids=position_ids, token_type_ids=token_type_ids,
                head_mask=None, offsets=offsets, gather_index=gather_index, _want_align_block=want_block,
                _align_rows=align_pos if want_block else None)
            align_loss = None
            if want_block:
                align_loss, _ = AlignLossFn.apply(att.align_block, align_pos, total_label)
        cfg = engine_cfg(self.global_enc.config)
        if torch.is_grad_enabled() and any(p.requires_grad for p in named.values()):
            run = ClsHeadRun(ps, list(named.keys()), cfg, N, T, g[0].shape[1])
            run.heads, run.n_layers = 8, self.cls_layer_num
            # hidden dropout of ClsLayer_lyx / BertOutput + the 0.1 probability dropout of cross_attention_lyx (v10:848)
            run.drop = E.make_drop(self.training, cfg_get(self.global_enc.config, "hidden_dropout_prob", 0.0),
                                   self.cls_layer_lyx[0].cross_attention.dropout, input_ids.device, 4 << 20)
            CLS_ensem = ClsLyxHeadFn.apply(run, T - 1, g[0], s_out, chunk_hidden, g[1], s_pool, *named.values())
        else:
            with torch.no_grad():
                CLS_ensem = cls_lyx_head_forward(ps, cfg, T - 1, g[0], s_out, chunk_hidden, g[1], s_pool,
                                                 n_layers=self.cls_layer_num, heads=8)
        return CLS_ensem, align_loss, ([], None)
