"""Language-model side of ModCR: RoBERTa-large (24L / 1024 / 16H) that takes ``prompt_embeddings`` --
the 5 vision + 5 alignment prompt vectors of ``Abstract_Specific`` (modeling/modeling_ensemble.py:473-503) -- next to its
token ids; call sites ``run_PMR_ModCR.py:772-781`` (construction) and ``modeling/modeling_ensemble.py:501-503``:

    roberta(input_ids=, token_type_ids=, attention_mask=, prompt_embeddings=prefix_emb [N,10,1024],
            input_mask=prompt_mask [N,10])[1]   ->  pooled [N,1024]

PARITY UNPINNED against the reference: its modified RoBERTa lives in ``local_transformers.zip`` (an adapter-transformers
fork, version unpinned), which the reference repository does not ship (``.MISSING_LARGE_BLOBS``).  The agreed stand-in
oracle (``oracle/standin_roberta.py``, pinned to HuggingFace ``RobertaModel`` of the installed transformers in
tests/test_oracle.py) is: HF RoBERTa with the prompt embeddings concatenated IN FRONT of the token embeddings right
after the embedding layer (after its LayerNorm), the key mask ``[input_mask ; attention_mask]``, the unchanged post-LN
encoder, and the pooler on the first row.  State-dict keys are HF ``RobertaModel``'s, so a ``roberta-large`` checkpoint
loads as is.

The arithmetic is the same kernel program as the Oscar encoders (``functions.EncoderFn``: fused QKV GEMM, prefix-free
mode-0 attention, GEMM epilogues, LayerNorm kernels) at the 24L / 1024 / 16H size of BASELINE.json configs[4]."""
import torch
from torch import nn

from . import engine as E
from .functions import EncoderFn, EncoderRun
from .modeling_bert import BertPooler, CaptionBertEncoder
from .modeling_utils import CaptionPreTrainedModel, cfg_get, engine_cfg


class RobertaConfig:
    """Attribute bag with the HF ``roberta-large`` defaults (config.json of the checkpoint the run scripts load)."""

    def __init__(self, **kw):
        d = dict(vocab_size=50265, hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                 hidden_act="gelu", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, max_position_embeddings=514,
                 type_vocab_size=1, initializer_range=0.02, layer_norm_eps=1e-5, pad_token_id=1, bos_token_id=0,
                 eos_token_id=2, output_attentions=False, output_hidden_states=False, use_img_layernorm=0,
                 img_feature_dim=0)
        d.update(kw)
        for k, v in d.items():
            setattr(self, k, v)


class RobertaEmbeddings(nn.Module):
    """Parameter container with HF's key names (no persistent buffers, like transformers >= 4.31)."""

    def __init__(self, config):
        super().__init__()
        self.padding_idx = config.pad_token_id
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=config.pad_token_id)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size,
                                                padding_idx=config.pad_token_id)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)


def create_position_ids_from_input_ids(input_ids, padding_idx):
    """HF modeling_roberta: non-padding tokens count up from padding_idx + 1, padding tokens keep padding_idx (index glue
    on int64 ids, bit-exact)."""
    mask = input_ids.ne(padding_idx).to(torch.int64)
    return torch.cumsum(mask, dim=1) * mask + padding_idx


class RobertaPromptModel(CaptionPreTrainedModel):
    base_model_prefix = "roberta"

    def __init__(self, config, add_pooling_layer=True):
        super().__init__(config)
        self.embeddings = RobertaEmbeddings(config)
        self.encoder = CaptionBertEncoder(config)
        self.pooler = BertPooler(config) if add_pooling_layer else None
        self._param_store = None
        self.init_weights()

    def _store(self):
        named = dict(self.named_parameters())
        ps = self._param_store
        if ps is None or ps.named.keys() != named.keys() or any(ps.named[k] is not named[k] for k in named):
            ps = E.ParamStore(named)
            self._param_store = ps
        for k, p in named.items():
            if p.dtype != torch.float32 or not p.is_cuda:
                raise RuntimeError("modcr_b200: parameter %s must be a float32 CUDA tensor (got %s on %s); there is no "
                                   "CPU / PyTorch fallback" % (k, p.dtype, p.device))
        return ps, list(named.keys()), list(named.values())

    def forward(self, input_ids=None, token_type_ids=None, attention_mask=None, prompt_embeddings=None, input_mask=None,
                position_ids=None):
        """Returns ``(sequence_output [N,Pn+T,H], pooled_output [N,H])`` (bf16).  ``prompt_embeddings`` ``[N,Pn,H]``
        carries gradients back to the mapping networks; ``input_mask`` ``[N,Pn]`` is its key mask (all ones in the
        reference, modeling_ensemble.py:472, 496)."""
        if self.pooler is None:
            raise NotImplementedError("the ModCR path reads the pooled output (modeling_ensemble.py:504)")
        N, T = input_ids.shape
        dev = input_ids.device
        H = self.config.hidden_size
        Pn = 0 if prompt_embeddings is None else prompt_embeddings.shape[1]
        if Pn and tuple(prompt_embeddings.shape) != (N, Pn, H):
            raise ValueError("prompt_embeddings must be [N,Pn,%d]" % H)
        if position_ids is None:
            position_ids = create_position_ids_from_input_ids(input_ids, self.embeddings.padding_idx)
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_ids)
        if attention_mask is None:
            attention_mask = torch.ones((N, T), dtype=torch.float32, device=dev)
        mask = attention_mask.to(torch.float32)
        if Pn:
            pm = torch.ones((N, Pn), dtype=torch.float32, device=dev) if input_mask is None else input_mask.to(torch.float32)
            mask = torch.cat([pm, mask], dim=-1)
        kbias = E.make_key_bias(mask)
        ps, names, params = self._store()
        run = EncoderRun(ps, names, engine_cfg(self.config), input_ids, token_type_ids, position_ids, None, kbias, 0,
                         n_pre=Pn)
        run.drop = E.make_drop(self.training, cfg_get(self.config, "hidden_dropout_prob", 0.0),
                               cfg_get(self.config, "attention_probs_dropout_prob", 0.0), dev, 7 << 20)
        extra = (prompt_embeddings,) if Pn else ()
        seq_out, pooled, _, _ = EncoderFn.apply(run, 0, *extra, *params)
        return seq_out, pooled
