"""Training-driver pieces of the hot path: the optimiser step of run_PMR_ModCR.py:127-145, 211-227
-- ``clip_grad_norm_`` over every parameter followed by ``transformers.AdamW.step()`` with two learning-rate groups
(``seq_enc`` x 0.1), eps 1e-5, bias correction, weight decay never passed -- as two kernel launches over all tensors
(csrc/optim.cu) instead of ~10 tiny launches per tensor.
"""
import ctypes as C

import numpy as np
import torch

from . import ops

F32 = torch.float32
CHUNK = 8192   # elements per block of the (tensor, chunk) table; must match OPT_CHUNK in csrc/optim.cu


class _Meta(C.Structure):   # mirrors modcr_optim_tensor
    _fields_ = [("p", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("n", C.c_int64), ("lr_mul", C.c_float),
                ("wd", C.c_float)]


def reference_param_groups(model, lr):
    """The two groups of run_PMR_ModCR.py:127-136: everything that requires grad at ``lr``, ``seq_enc`` at ``0.1 lr``."""
    seq = [p for n, p in model.named_parameters() if "seq_enc" in n]
    rest = [p for n, p in model.named_parameters() if "seq_enc" not in n and p.requires_grad]
    return [{"params": rest, "lr": lr}, {"params": seq, "lr": lr * 0.1}]


class FusedAdamW:
    """HuggingFace-AdamW semantics (the optimiser the reference imports: ``transformers.AdamW``; eps is added to
    ``sqrt(v)`` before the bias correction), fp32 master weights updated in place, optional global-norm clipping fused
    into the same pass.  ``param_groups``: list of ``{"params": [...], "lr": float, "weight_decay": float}``; group
    learning rates are kept as multipliers of ``lr`` so a scheduler only has to change one number (``set_lr``)."""

    def __init__(self, param_groups, lr, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, max_grad_norm=None):
        if isinstance(param_groups, (list, tuple)) and param_groups and not isinstance(param_groups[0], dict):
            param_groups = [{"params": list(param_groups)}]
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.max_grad_norm = None if max_grad_norm is None else float(max_grad_norm)
        self.params, lr_mul, wds = [], [], []
        for g in param_groups:
            glr, gwd = float(g.get("lr", lr)), float(g.get("weight_decay", weight_decay))
            for p in g["params"]:
                if not p.is_cuda or p.dtype != F32 or not p.is_contiguous():
                    raise RuntimeError("modcr_b200: FusedAdamW needs contiguous fp32 CUDA parameters (no CPU fallback)")
                self.params.append(p)
                lr_mul.append(glr / self.lr if self.lr != 0.0 else 1.0)
                wds.append(gwd)
        if not self.params:
            raise ValueError("FusedAdamW: no parameters")
        dev = self.params[0].device
        self.device = dev
        self.step_count = 0
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        meta = (_Meta * len(self.params))()
        blk_t, blk_c = [], []
        for i, p in enumerate(self.params):
            meta[i].p, meta[i].m, meta[i].v = p.data_ptr(), self.m[i].data_ptr(), self.v[i].data_ptr()
            meta[i].n, meta[i].lr_mul, meta[i].wd = p.numel(), lr_mul[i], wds[i]
            nchunks = (p.numel() + CHUNK - 1) // CHUNK
            blk_t += [i] * nchunks
            blk_c += list(range(nchunks))
        raw = np.frombuffer(bytes(meta), dtype=np.uint8).copy()
        self.meta = torch.from_numpy(raw).to(dev)
        self.blk_tensor = torch.tensor(blk_t, dtype=torch.int32, device=dev)
        self.blk_chunk = torch.tensor(blk_c, dtype=torch.int32, device=dev)
        self.nblocks = len(blk_t)
        self._grad_ptrs_host = torch.zeros(len(self.params), dtype=torch.int64).pin_memory()
        self.grad_ptrs = torch.zeros(len(self.params), dtype=torch.int64, device=dev)
        self._hyper_host = torch.zeros(4, dtype=F32).pin_memory()
        self.hyper = torch.zeros(4, dtype=F32, device=dev)
        self.grad_sq = torch.zeros(1, dtype=F32, device=dev)
        self._staged = None   # event: the previous step's host->device copies of the pinned staging buffers are done

    def set_lr(self, lr):
        """Scheduler hook: the base learning rate (group ratios are preserved)."""
        self.lr = float(lr)

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    @torch.no_grad()
    def step(self):
        """One optimiser step on whatever gradients are present; returns the (unclipped) global gradient norm as a
        0-dim device tensor when clipping is enabled, else ``None``."""
        self.step_count += 1
        t = self.step_count
        b1, b2 = self.betas
        if self._staged is not None:
            self._staged.synchronize()       # do not overwrite pinned staging memory a pending copy still reads
        for i, p in enumerate(self.params):
            g = p.grad
            if g is not None and (g.dtype != F32 or not g.is_contiguous() or g.device != p.device):
                raise RuntimeError("modcr_b200: FusedAdamW needs contiguous fp32 gradients on the parameter's device")
            self._grad_ptrs_host[i] = 0 if g is None else g.data_ptr()
        self.grad_ptrs.copy_(self._grad_ptrs_host, non_blocking=True)
        self._hyper_host[0], self._hyper_host[1], self._hyper_host[2] = self.lr, 1.0 - b1 ** t, 1.0 - b2 ** t
        self._hyper_host[3] = self.max_grad_norm if self.max_grad_norm is not None else 0.0
        self.hyper.copy_(self._hyper_host, non_blocking=True)
        self._staged = torch.cuda.Event()
        self._staged.record()
        norm = None
        if self.max_grad_norm is not None:
            self.grad_sq.zero_()
            ops.grad_sqnorm(self.meta, self.grad_ptrs, self.blk_tensor, self.blk_chunk, self.nblocks, self.grad_sq)
            norm = self.grad_sq.sqrt().squeeze(0)
        ops.adamw_step(self.meta, self.grad_ptrs, self.blk_tensor, self.blk_chunk, self.nblocks, self.hyper,
                       self.grad_sq if self.max_grad_norm is not None else None, b1, b2, self.eps)
        # the kernels wrote through raw pointers: bump the tensors' version counters so that the cached bf16 operand
        # copies (engine.ParamStore, keyed on data_ptr + version) are refreshed on the next forward
        torch._C._autograd._unsafe_set_version_counter(tuple(self.params), tuple(p._version + 1 for p in self.params))
        return norm
