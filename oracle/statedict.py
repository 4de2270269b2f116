"""TEST INFRASTRUCTURE ONLY -- the reference's state_dict surface,
restated as (key, shape) lists so weights can be materialised without the
reference or the product, and so the product's ``state_dict()`` can be checked
key-for-key (checkpoint compatibility, run_PMR_ModCR.py:752-763, :819-832)."""
import torch


def _bert_layer(pfx, H, I):
    out = []
    for n in ("query", "key", "value"):
        out += [(pfx + "attention.self.%s.weight" % n, (H, H)), (pfx + "attention.self.%s.bias" % n, (H,))]
    out += [(pfx + "attention.output.dense.weight", (H, H)), (pfx + "attention.output.dense.bias", (H,)),
            (pfx + "attention.output.LayerNorm.weight", (H,)), (pfx + "attention.output.LayerNorm.bias", (H,)),
            (pfx + "intermediate.dense.weight", (I, H)), (pfx + "intermediate.dense.bias", (I,)),
            (pfx + "output.dense.weight", (H, I)), (pfx + "output.dense.bias", (H,)),
            (pfx + "output.LayerNorm.weight", (H,)), (pfx + "output.LayerNorm.bias", (H,))]
    return out


def encoder_keys(pfx, cfg, seq=False):
    """BertImgModel (modeling_transfomres.py:567-598) / SeqBertImgModel (v10:239-259)."""
    H, I, L = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_hidden_layers"]
    out = [(pfx + "embeddings.position_ids", (1, cfg["max_position_embeddings"])),
           (pfx + "embeddings.word_embeddings.weight", (cfg["vocab_size"], H)),
           (pfx + "embeddings.position_embeddings.weight", (cfg["max_position_embeddings"], H)),
           (pfx + "embeddings.token_type_embeddings.weight", (2, H)),
           (pfx + "embeddings.LayerNorm.weight", (H,)), (pfx + "embeddings.LayerNorm.bias", (H,))]
    for i in range(L):
        out += _bert_layer("%sencoder.layer.%d." % (pfx, i), H, I)
    out += [(pfx + "pooler.dense.weight", (H, H)), (pfx + "pooler.dense.bias", (H,)),
            (pfx + "img_embedding.weight", (H, cfg["img_feature_dim"])), (pfx + "img_embedding.bias", (H,))]
    if cfg["use_img_layernorm"]:
        out += [(pfx + "LayerNorm.weight", (H,)), (pfx + "LayerNorm.bias", (H,))]
    if seq:
        out += [(pfx + "edge_dense.weight", (1, H))]
    return out


def cls_layer2_keys(pfx, cfg):
    """ClsLayer2(BertLayer) v10:806-814 -- includes the inherited, never-called attention.*"""
    H, I = cfg["hidden_size"], cfg["intermediate_size"]
    out = [k for k in _bert_layer(pfx, H, I)]
    out += [(pfx + "cls_q_proj.weight", (H, H)), (pfx + "cls_q_proj.bias", (H,)),
            (pfx + "align_k_proj.weight", (H, H)), (pfx + "align_k_proj.bias", (H,)),
            (pfx + "dense.weight", (H, H)), (pfx + "dense.bias", (H,)),
            (pfx + "LayerNorm.weight", (H,)), (pfx + "LayerNorm.bias", (H,))]
    return out


def cls_layer_lyx_keys(pfx, cfg):
    """ClsLayer_lyx(BertLayer) v10:845-854."""
    H, I = cfg["hidden_size"], cfg["intermediate_size"]
    out = [k for k in _bert_layer(pfx, H, I)]
    out += [(pfx + "ensemble.weight", (1, 2 * H)), (pfx + "ensemble.bias", (1,))]
    for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
        out += [(pfx + "cross_attention.%s.weight" % n, (H, H)), (pfx + "cross_attention.%s.bias" % n, (H,))]
    out += [(pfx + "dense.weight", (H, H)), (pfx + "dense.bias", (H,)),
            (pfx + "LayerNorm.weight", (H,)), (pfx + "LayerNorm.bias", (H,))]
    return out


def chunkalign_keys(cfg):
    """ChunkAlign_CLS_enc4_align v10:1017-1027."""
    H = cfg["hidden_size"]
    out = encoder_keys("global_enc.", cfg) + encoder_keys("seq_enc.", cfg, seq=True)
    out += [("cls_ensemble.weight", (H, 2 * H)), ("cls_ensemble.bias", (H,))]
    for i in range(3):
        out += cls_layer2_keys("cls_layer.%d." % i, cfg)
    out += [("classifier.weight", (2, H)), ("classifier.bias", (2,))]
    return out


def chunkalign_ensemble_keys(cfg):
    """ChunkAlign_CLS_enc4_align_ensemble v10:873-889."""
    H = cfg["hidden_size"]
    out = encoder_keys("global_enc.", cfg) + encoder_keys("seq_enc.", cfg, seq=True)
    out += [("cls_ensemble_1.weight", (H, 2 * H)), ("cls_ensemble_1.bias", (H,))]
    for i in range(2):
        out += cls_layer2_keys("cls_layer.%d." % i, cfg)
    for i in range(2):
        out += cls_layer_lyx_keys("cls_layer_lyx.%d." % i, cfg)
    out += [("classifier.weight", (2, H)), ("classifier.bias", (2,)),
            ("fusion_align.weight", (1024, 2 * H)), ("fusion_align.bias", (1024,)),
            ("prior.weight", (1, H)), ("prior.bias", (1,))]
    return out


def materialise(keys):
    sd = {}
    for k, shape in keys:
        if k.endswith("position_ids"):
            sd[k] = torch.arange(shape[1]).expand(shape).clone()
        else:
            sd[k] = torch.zeros(shape, dtype=torch.float32)
    return sd
