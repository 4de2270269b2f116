"""TEST INFRASTRUCTURE ONLY -- agreed stand-in oracle of the language-model side.

The reference's RoBERTa-large with ``prompt_embeddings`` (call sites modeling/modeling_ensemble.py:501-503,
run_PMR_ModCR.py:772-781) is a modified adapter-transformers fork shipped as ``local_transformers.zip`` (version unpinned),
which the repository does NOT contain (.MISSING_LARGE_BLOBS:1).  PARITY IS THEREFORE UNPINNED against the reference; this
file restates the published RoBERTa algorithm (HuggingFace ``transformers`` -- 5.5.0 in this image --
models/roberta/modeling_roberta.py: ``RobertaEmbeddings.forward`` incl. ``create_position_ids_from_input_ids``,
``RobertaLayer`` = the post-LN BERT layer, ``RobertaPooler``) with the one documented change: the prompt embeddings are
concatenated in front of the token embeddings right after the embedding layer and masked by ``input_mask``.
tests/test_oracle.py pins it against HF ``RobertaModel`` itself."""
import torch

import restate


def position_ids(input_ids, padding_idx):
    mask = input_ids.ne(padding_idx).to(torch.int64)
    return torch.cumsum(mask, dim=1) * mask + padding_idx


def roberta_prompt_forward(sd, cfg, input_ids, token_type_ids, attention_mask, prompt_embeddings=None, input_mask=None):
    """``sd``: HF RobertaModel state_dict; ``cfg``: dict(num_hidden_layers, num_attention_heads, layer_norm_eps,
    pad_token_id).  Returns (sequence_output, pooled_output)."""
    eps, nh = cfg["layer_norm_eps"], cfg["num_attention_heads"]
    if token_type_ids is None:
        token_type_ids = torch.zeros_like(input_ids)
    pos = position_ids(input_ids, cfg["pad_token_id"])
    e = (sd["embeddings.word_embeddings.weight"][input_ids] + sd["embeddings.token_type_embeddings.weight"][token_type_ids]
         + sd["embeddings.position_embeddings.weight"][pos])
    x = restate.layer_norm(sd, "embeddings.LayerNorm", e, eps)
    mask = attention_mask.to(torch.float32)
    if prompt_embeddings is not None:
        N, Pn = prompt_embeddings.shape[:2]
        pm = torch.ones(N, Pn) if input_mask is None else input_mask.to(torch.float32)
        x = torch.cat([prompt_embeddings.to(x.dtype), x], dim=1)
        mask = torch.cat([pm, mask], dim=-1)
    ext = (1.0 - mask)[:, None, None, :] * restate.NEG
    for i in range(cfg["num_hidden_layers"]):
        x, _ = restate.bert_layer(sd, "encoder.layer.%d." % i, x, ext, nh, eps)
    pooled = torch.tanh(restate.linear(sd, "pooler.dense", x[:, 0]))
    return x, pooled
