"""TEST INFRASTRUCTURE ONLY -- deterministic, construction-order independent
parameter fill keyed by the state_dict key name, so the reference classes (when
goldens are generated), the oracle restatement and the CUDA product all run on
identical weights without a checkpoint file.

Follows the *distribution* of BertPreTrainedModel._init_weights
(a_transformers/modeling_bert.py:729-743: N(0, 0.02) Linear/Embedding) but
deliberately uses non-zero biases and perturbed LayerNorm affine parameters so
that bias / gamma / beta handling is actually exercised by the parity tests."""
import zlib

import torch


def fill_params(state_dict, seed=0):
    """In-place fill of every floating tensor in ``state_dict``; returns it."""
    with torch.no_grad():
        for key in sorted(state_dict.keys()):
            t = state_dict[key]
            if not torch.is_floating_point(t):
                continue
            g = torch.Generator().manual_seed((zlib.crc32(key.encode()) + 7919 * seed) & 0x7FFFFFFF)
            r = torch.randn(t.shape, generator=g, dtype=torch.float32)
            if "LayerNorm.weight" in key:
                v = 1.0 + 0.1 * r
            elif "LayerNorm.bias" in key:
                v = 0.05 * r
            elif key.endswith(".bias"):
                v = 0.02 * r
            else:
                v = 0.02 * r
            t.copy_(v.to(t.dtype))
    return state_dict
