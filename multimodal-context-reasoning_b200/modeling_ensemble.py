"""Frozen-encoder regime of the reference: ``Abstract_Specific`` and ``PrefixMLP`` of
modeling/modeling_ensemble.py:407-540 -- the configuration the paper trains (run_PMR_ModCR.py:802).

Per batch (modeling_ensemble.py:459-539):
  1. no_grad ``[CLS] + regions`` pass of ``calec.global_enc`` (sequence = 1 text token + R regions)        (:466-471)
  2. ``mapping_network_vision``  : Dropout .1 -> Linear 768->3840 -> Tanh -> Dropout .1 -> Linear 3840->5120 -> [N,5,1024]
  3. ``calec`` (``ChunkAlign_CLS_enc4_align_ensemble``: frozen encoders, trainable ClsLayer_lyx head)       (:484-491)
  4. ``mapping_network_alignment`` on ``CLS_ensem`` -> [N,5,1024]                                            (:494)
  5. the language model (RoBERTa-large in the reference, SURVEY 8 f4: its modified source is a missing blob, so it
     stays the caller's ``roberta_model`` -- any module with the reference's call signature) gets the 10 prompt
     embeddings and returns a pooled [N,1024] vector at index 1                                              (:502-504)
  6. ``abst_confidence_scorer`` Linear 1024->1, logits viewed [B,4], CrossEntropyLoss against float labels   (:513-537)

Everything on the CALeC side runs through libmodcr_b200.so: both mapping networks are one autograd Function each
(tcgen05 GEMMs with the bias + tanh epilogue, counter-based dropout masks regenerated in backward), the scorer is the
fused classifier + CE kernel.  State-dict keys are the reference's (``mapping_network_vision.1.weight`` ...).
"""
import torch
from torch import nn

from . import engine as E
from . import ops
from .functions import ScorerFn, _as_bf16

BF16, F32 = torch.bfloat16, torch.float32


class MappingNetFn(torch.autograd.Function):
    """y = W2 . drop2(tanh(W1 . drop1(x) + b1)) + b2  (fp32 out).  ``drop`` = (seed tensor, site1, site2, p) or None."""

    @staticmethod
    def forward(ctx, x, W1, b1, W2, b2, w1_16, w2_16, drop):
        N, K = x.shape
        Hm, O = W1.shape[0], W2.shape[0]
        dev = x.device
        xb = _as_bf16(x).contiguous()
        if drop is not None:
            xd = torch.empty_like(xb)
            ops.dropout(xb, xd, None, None, xb.numel(), drop[0], drop[1], drop[3])
        else:
            xd = xb
        t = E._empty((N, Hm), BF16, dev)
        ops.gemm(xd, w1_16, t, N, Hm, K, bias=b1.detach(), act=2)
        if drop is not None:
            h = torch.empty_like(t)
            ops.dropout(t, h, None, None, t.numel(), drop[0], drop[2], drop[3])
        else:
            h = t
        y = E._empty((N, O), F32, dev)
        ops.gemm(h, w2_16, y, N, O, Hm, bias=b2.detach())
        ctx.sv = (xd, t, h, w1_16, w2_16, drop, x.dtype, x.requires_grad)
        return y

    @staticmethod
    def backward(ctx, dy):
        xd, t, h, w1_16, w2_16, drop, xdtype, need_dx = ctx.sv
        N, K = xd.shape
        Hm, O = w1_16.shape[0], w2_16.shape[0]
        dev = xd.device
        dyb = _as_bf16(dy).contiguous()
        db2 = E._zeros((O,), F32, dev)
        dW2 = E.wgrad(dyb, h, O, Hm, N, bias_out=db2)
        dh = E._empty((N, Hm), BF16, dev)
        ops.gemm(dyb, w2_16, dh, N, Hm, O, b_mn=1)
        if drop is not None:
            ops.dropout(dh, dh, None, None, dh.numel(), drop[0], drop[2], drop[3])
        dpre = torch.empty_like(dh)
        ops.tanh_bwd(dh, t, dpre)
        db1 = E._zeros((Hm,), F32, dev)
        dW1 = E.wgrad(dpre, xd, Hm, K, N, bias_out=db1)
        dx = None
        if need_dx:
            dxb = E._empty((N, K), BF16, dev)
            ops.gemm(dpre, w1_16, dxb, N, K, Hm, b_mn=1)
            if drop is not None:
                ops.dropout(dxb, dxb, None, None, dxb.numel(), drop[0], drop[1], drop[3])
            dx = dxb.to(xdtype)
        return dx, dW1, db1, dW2, db2, None, None, None


class _StoreMixin:
    """bf16 operand copies of this module's own Linear weights (``engine.ParamStore``; FusedAdamW keeps them fresh)."""
    _param_store = None

    def _store(self, names):
        named = {k: p for k, p in self.named_parameters() if k in names}
        ps = self._param_store
        if ps is None or ps.named.keys() != named.keys() or any(ps.named[k] is not named[k] for k in named):
            for k, p in named.items():
                if p.dtype != F32 or not p.is_cuda:
                    raise RuntimeError("modcr_b200: parameter %s must be a float32 CUDA tensor (no CPU fallback)" % k)
            ps = E.ParamStore(named)
            self._param_store = ps
        return ps


class PrefixMLP(nn.Module, _StoreMixin):
    """modeling_ensemble.py:407-421: Linear(feature -> hidden*prefix/2) -> Tanh -> Linear(-> hidden*prefix)."""

    def __init__(self, feature_size, hidden_size, prefix_len, hidden_dropout_prob):
        super().__init__()
        self.dense0 = nn.Linear(feature_size, (hidden_size * prefix_len) // 2)
        self.dense1 = nn.Linear((hidden_size * prefix_len) // 2, hidden_size * prefix_len)
        self.tanh = nn.Tanh()

    def forward(self, hidden_states):
        ps = self._store({"dense0.weight", "dense1.weight"})
        shape = hidden_states.shape
        x = hidden_states.reshape(-1, shape[-1])
        y = MappingNetFn.apply(x, self.dense0.weight, self.dense0.bias, self.dense1.weight, self.dense1.bias,
                               ps.bf16("dense0.weight"), ps.bf16("dense1.weight"), None)
        return y.view(*shape[:-1], -1)


def _mapping_network():
    return nn.Sequential(nn.Dropout(p=0.1), nn.Linear(768, 768 * 5, bias=True), nn.Tanh(), nn.Dropout(p=0.1),
                         nn.Linear(768 * 5, 1024 * 5, bias=True))


class Abstract_Specific(nn.Module, _StoreMixin):
    """modeling_ensemble.py:424-540.  ``calec_model``: ``ChunkAlign_CLS_enc4_align_ensemble`` of this package;
    ``roberta_model``: the caller's language model, called exactly as the reference calls it
    (``input_ids, token_type_ids, attention_mask, prompt_embeddings [N,10,1024], input_mask [N,10]``) and returning a
    sequence whose element 1 is the pooled ``[N,1024]`` vector."""

    def __init__(self, roberta_model, calec_model, clip_model=None, num_labels=4):
        super().__init__()
        self.num_labels = num_labels
        self.calec = calec_model
        self.roberta = roberta_model
        if clip_model is not None:
            self.clip_model = clip_model
            self.classifier = nn.Linear(1024 + 768 + 512, 1)
        else:
            self.classifier = nn.Linear(768 + 768, 1)
        self.abst_confidence_scorer = nn.Linear(1024, 1)
        self.confidence_scorer = nn.Linear(768, 1)
        self.mapping_network_alignment = _mapping_network()
        self.mapping_network_vision = _mapping_network()
        self.promptfuse = torch.nn.Embedding(2, 1024)

    def _map(self, which, x, drop_base):
        net = getattr(self, which)
        ps = self._store({"mapping_network_alignment.1.weight", "mapping_network_alignment.4.weight",
                          "mapping_network_vision.1.weight", "mapping_network_vision.4.weight"})
        drop = None
        if self.training and net[0].p > 0.0:
            if net[0].p != net[3].p:
                raise RuntimeError("modcr_b200: the two dropouts of a mapping network must share p")
            d = E.make_drop(True, net[0].p, 0.0, x.device, drop_base)
            drop = (d.seed, d.site(0, 0), d.site(0, 1), net[0].p)
        return MappingNetFn.apply(x, net[1].weight, net[1].bias, net[4].weight, net[4].bias,
                                  ps.bf16(which + ".1.weight"), ps.bf16(which + ".4.weight"), drop)

    def forward(self, image, text, roberta_input_ids, roberta_token_type_ids, roberta_attention_mask, input_ids, img_feat,
                input_mask=None, token_type_ids=None, position_ids=None, head_mask=None, encoder_history_states=None,
                offsets=None, chunk_attention_mask=None, gather_index=None, label=None, align_pos=None,
                total_label=None):
        N = input_ids.size(0)
        # vision representation: [CLS] + regions through the frozen global encoder (:466-471)
        with torch.no_grad():
            img_attention_mask = torch.cat([input_mask[:, :1], input_mask[:, -img_feat.size(1):]], dim=-1)
            image_features_ = self.calec.global_enc(input_ids[:, :1].contiguous(), img_feats=img_feat,
                                                    attention_mask=img_attention_mask.contiguous(), position_ids=None,
                                                    token_type_ids=None, head_mask=None, encoder_history_states=None)
            vis_cls = image_features_[0][:, 0, :].contiguous()
        prefix_vision = self._map("mapping_network_vision", vis_cls, 5 << 20).reshape(N, 5, 1024)
        vision_mask = input_mask[:, :1].repeat(1, 5)

        CALeC_encoder_output, align_loss, specific_alignment = self.calec(
            input_ids=input_ids, img_feat=img_feat, input_mask=input_mask, token_type_ids=token_type_ids,
            position_ids=position_ids, head_mask=head_mask, encoder_history_states=encoder_history_states,
            offsets=offsets, chunk_attention_mask=chunk_attention_mask, gather_index=gather_index, align_pos=align_pos,
            total_label=total_label, abstract_hidden_states=None)
        Alignment_prompt = self._map("mapping_network_alignment", CALeC_encoder_output, 6 << 20).view(N, 5, 1024)
        align_mask = input_mask[:, :1].repeat(1, 5)

        prefix_emb = torch.cat([prefix_vision, Alignment_prompt], dim=1)
        prompt_mask = torch.cat([vision_mask, align_mask], dim=1)
        roberta_encoder_outputs = self.roberta(input_ids=roberta_input_ids, token_type_ids=roberta_token_type_ids,
                                               attention_mask=roberta_attention_mask, prompt_embeddings=prefix_emb,
                                               input_mask=prompt_mask)
        abstract_level = roberta_encoder_outputs[1]

        soft = None if label is None else label.reshape(-1, self.num_labels).to(F32).contiguous()
        w, b = self.abst_confidence_scorer.weight, self.abst_confidence_scorer.bias
        if soft is not None:
            loss, logits = ScorerFn.apply(abstract_level, w, b, None, soft, self.num_labels, self.num_labels, False)
            loss_abstract = loss
        else:   # scoring only: the fused kernel wants some target, its loss output is discarded
            dummy = torch.zeros(N // self.num_labels, self.num_labels, dtype=F32, device=input_ids.device)
            with torch.no_grad():
                _, logits = ScorerFn.apply(abstract_level, w, b, None, dummy, self.num_labels, self.num_labels, False)
            loss = loss_abstract = None
        reshaped_logits = logits.view(-1, self.num_labels)
        return loss, (None, None, loss_abstract, None), reshaped_logits
