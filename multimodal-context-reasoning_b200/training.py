"""Training-driver pieces of the hot path: the optimiser step of run_PMR_ModCR.py:127-145, 211-227
-- ``clip_grad_norm_`` over every parameter followed by ``transformers.AdamW.step()`` with two learning-rate groups
(``seq_enc`` x 0.1), eps 1e-5, bias correction, weight decay never passed -- as two kernel launches over all tensors
(csrc/optim.cu) instead of ~10 tiny launches per tensor.  The same pass rewrites the bf16 / packed operand copies the
forward kernels read (``engine.ParamStore``), ``clip_grad_norm_`` reproduces the clip-after-every-backward of gradient
accumulation (run_PMR_ModCR.py:211-215) and ``linear_schedule_with_warmup`` the scheduler of :139-140.
"""
import ctypes as C

import numpy as np
import torch

from . import ops

F32 = torch.float32
CHUNK = 8192   # elements per block of the (tensor, chunk) table; must match OPT_CHUNK in csrc/optim.cu


class _Meta(C.Structure):   # mirrors modcr_optim_tensor
    _fields_ = [("p", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("shadow", C.c_void_p), ("n", C.c_int64),
                ("lr_mul", C.c_float), ("wd", C.c_float), ("shadow_kind", C.c_int32), ("reserved", C.c_int32)]


def linear_schedule_with_warmup(step, warmup_steps, total_steps):
    """LR multiplier of ``get_linear_schedule_with_warmup`` (run_PMR_ModCR.py:139-140) after ``step`` scheduler steps."""
    if step < warmup_steps:
        return float(step) / float(max(1, warmup_steps))
    return max(0.0, float(total_steps - step) / float(max(1, total_steps - warmup_steps)))


def constant_schedule_with_warmup(step, warmup_steps):
    """LR multiplier of ``get_constant_schedule_with_warmup`` (``--scheduler constant``, run_PMR_ModCR.py:137-139)."""
    if step < warmup_steps:
        return float(step) / float(max(1.0, warmup_steps))
    return 1.0


def reference_param_groups(model, lr):
    """The two groups of run_PMR_ModCR.py:127-136: everything that requires grad at ``lr``, ``seq_enc`` at ``0.1 lr``."""
    seq = [p for n, p in model.named_parameters() if "seq_enc" in n]
    rest = [p for n, p in model.named_parameters() if "seq_enc" not in n and p.requires_grad]
    return [{"params": rest, "lr": lr}, {"params": seq, "lr": lr * 0.1}]


class FusedAdamW:
    """HuggingFace-AdamW semantics (the optimiser the reference imports: ``transformers.AdamW``; eps is added to
    ``sqrt(v)`` before the bias correction), fp32 master weights updated in place, optional global-norm clipping fused
    into the same pass.  ``param_groups``: list of ``{"params": [...], "lr": float, "weight_decay": float}``; group
    learning rates are kept as multipliers of ``lr`` so a scheduler only has to change one number (``set_lr``).

    ``model``: when given, the kernel-operand copies its ``engine.ParamStore`` objects hold (bf16 weights, the packed
    QKV weight / bias) are rewritten by the update kernel itself and re-validated, so the forward pass after a step
    launches no cast kernels.  The copies exist after the first forward; bindings are re-derived whenever they change."""

    def __init__(self, param_groups, lr, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, max_grad_norm=None,
                 model=None):
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
        self._lr_mul, self._wds = lr_mul, wds
        self._stores = []
        if model is not None:
            for mod in model.modules():
                for attr in ("_param_store", "_head_store"):
                    if hasattr(mod, attr):
                        self._stores.append((mod, attr))
        self._bound = []        # (store, cache name, parameter names, cache entry) whose copies the kernel rewrites
        self._shadow = {}       # parameter index -> (tensor view kept alive, kind)
        blk_t, blk_c = [], []
        for i, p in enumerate(self.params):
            nchunks = (p.numel() + CHUNK - 1) // CHUNK
            blk_t += [i] * nchunks
            blk_c += list(range(nchunks))
        self.meta = None
        self._write_meta()
        self.blk_tensor = torch.tensor(blk_t, dtype=torch.int32, device=dev)
        self.blk_chunk = torch.tensor(blk_c, dtype=torch.int32, device=dev)
        self.nblocks = len(blk_t)
        self._grad_ptrs_host = torch.zeros(len(self.params), dtype=torch.int64).pin_memory()
        self.grad_ptrs = torch.zeros(len(self.params), dtype=torch.int64, device=dev)
        self._hyper_host = torch.zeros(4, dtype=F32).pin_memory()
        self.hyper = torch.zeros(4, dtype=F32, device=dev)
        self.grad_sq = torch.zeros(1, dtype=F32, device=dev)
        self._staged = None   # event: the previous step's host->device copies of the pinned staging buffers are done

    def _write_meta(self):
        meta = (_Meta * len(self.params))()
        for i, p in enumerate(self.params):
            meta[i].p, meta[i].m, meta[i].v = p.data_ptr(), self.m[i].data_ptr(), self.v[i].data_ptr()
            meta[i].n, meta[i].lr_mul, meta[i].wd = p.numel(), self._lr_mul[i], self._wds[i]
            sh = self._shadow.get(i)
            if sh is not None:
                meta[i].shadow, meta[i].shadow_kind = sh[0].data_ptr(), sh[1]
        raw = torch.from_numpy(np.frombuffer(bytes(meta), dtype=np.uint8).copy())
        if self.meta is None:
            self.meta = raw.to(self.device)
        else:
            self.meta.copy_(raw)    # synchronous: `raw` is pageable and dies with this call

    def _rebind(self):
        """Derive (parameter -> operand copy) from the ParamStore caches as they are now."""
        index = {id(p): i for i, p in enumerate(self.params)}
        self._bound, self._shadow = [], {}
        for mod, attr in self._stores:
            ps = getattr(mod, attr, None)
            if ps is None:
                continue
            for cname, ent in ps._cache.items():
                if cname.endswith("#pad"):
                    continue       # row-padded copy (img_embedding.weight): left to the lazy cast
                if cname.endswith("#qkv"):
                    pfx = cname[:-4]
                    names = [pfx + n + s for n in ("query", "key", "value") for s in (".weight", ".bias")]
                    W, b = ent[1], ent[2]
                    H = W.shape[1]
                    views = {}
                    for j, n in enumerate(("query", "key", "value")):
                        views[pfx + n + ".weight"] = (W[j * H:(j + 1) * H], 1)
                        views[pfx + n + ".bias"] = (b[j * H:(j + 1) * H], 2)
                else:
                    names = [cname]
                    views = {cname: (ent[1], 1)}
                ok = True
                for n, (view, kind) in views.items():
                    par = ps.named.get(n)
                    if par is None or not view.is_contiguous() or view.numel() != par.numel():
                        ok = False
                if not ok:
                    continue
                for n, (view, kind) in views.items():
                    i = index.get(id(ps.named[n]))
                    if i is not None:
                        self._shadow[i] = (view, kind)
                self._bound.append((ps, cname, names, ent))
        self._write_meta()

    def _bindings_current(self):
        n = 0
        for mod, attr in self._stores:
            ps = getattr(mod, attr, None)
            if ps is not None:
                n += sum(1 for c in ps._cache if not c.endswith("#pad"))
        if n != len(self._bound):
            return False
        return all(ps._cache.get(cname) is ent for ps, cname, _, ent in self._bound)

    def set_lr(self, lr):
        """Scheduler hook: the base learning rate (group ratios are preserved)."""
        self.lr = float(lr)

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def _stage_grad_ptrs(self):
        if self._staged is not None:
            self._staged.synchronize()       # do not overwrite pinned staging memory a pending copy still reads
        for i, p in enumerate(self.params):
            g = p.grad
            if g is not None and (g.dtype != F32 or not g.is_contiguous() or g.device != p.device):
                raise RuntimeError("modcr_b200: FusedAdamW needs contiguous fp32 gradients on the parameter's device")
            self._grad_ptrs_host[i] = 0 if g is None else g.data_ptr()
        self.grad_ptrs.copy_(self._grad_ptrs_host, non_blocking=True)

    @torch.no_grad()
    def clip_grad_norm_(self, max_norm=None):
        """``torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)`` in place over this optimiser's tensors (two
        launches); the reference calls it after EVERY backward, also on accumulation micro-steps that do not step
        (run_PMR_ModCR.py:214-215).  Returns the total norm before clipping (0-dim device tensor)."""
        max_norm = self.max_grad_norm if max_norm is None else float(max_norm)
        if max_norm is None:
            raise ValueError("clip_grad_norm_: no max_norm given and the optimiser was built without max_grad_norm")
        self._stage_grad_ptrs()
        self._staged = torch.cuda.Event()
        self._staged.record()
        self.grad_sq.zero_()
        ops.grad_sqnorm(self.meta, self.grad_ptrs, self.blk_tensor, self.blk_chunk, self.nblocks, self.grad_sq)
        norm = self.grad_sq.sqrt().squeeze(0)
        ops.grad_clip(self.meta, self.grad_ptrs, self.blk_tensor, self.blk_chunk, self.nblocks, self.grad_sq, max_norm)
        return norm

    @torch.no_grad()
    def step(self, clip=True):
        """One optimiser step on whatever gradients are present; returns the (unclipped) global gradient norm as a
        0-dim device tensor when clipping is enabled, else ``None``.  ``clip=False``: the gradients were already
        clipped in place by ``clip_grad_norm_`` (accumulation loop)."""
        self.step_count += 1
        t = self.step_count
        b1, b2 = self.betas
        if self._stores and not self._bindings_current():
            self._rebind()
        fresh = [ent[0] == ps._key(names) for ps, _, names, ent in self._bound]
        self._stage_grad_ptrs()
        self._hyper_host[0], self._hyper_host[1], self._hyper_host[2] = self.lr, 1.0 - b1 ** t, 1.0 - b2 ** t
        do_clip = clip and self.max_grad_norm is not None
        self._hyper_host[3] = self.max_grad_norm if do_clip else 0.0
        self.hyper.copy_(self._hyper_host, non_blocking=True)
        self._staged = torch.cuda.Event()
        self._staged.record()
        norm = None
        if do_clip:
            self.grad_sq.zero_()
            ops.grad_sqnorm(self.meta, self.grad_ptrs, self.blk_tensor, self.blk_chunk, self.nblocks, self.grad_sq)
            norm = self.grad_sq.sqrt().squeeze(0)
        ops.adamw_step(self.meta, self.grad_ptrs, self.blk_tensor, self.blk_chunk, self.nblocks, self.hyper,
                       self.grad_sq if do_clip else None, b1, b2, self.eps)
        # the kernels wrote through raw pointers: bump the tensors' version counters so that the cached bf16 operand
        # copies (engine.ParamStore, keyed on data_ptr + version) are refreshed on the next forward
        torch._C._autograd._unsafe_set_version_counter(tuple(self.params), tuple(p._version + 1 for p in self.params))
        # ... except the copies the kernel has just rewritten itself: re-key them to the new versions.  A copy is only
        # re-validated if it was current before the step and every one of its optimised sources had a gradient (a
        # tensor without one is skipped by the kernel and keeps its version-bumped, unchanged value).
        for (ps, cname, names, ent), ok in zip(self._bound, fresh):
            if ok and ps._cache.get(cname) is ent:
                new = (ps._key(names),) + tuple(ent[1:])
                ps._cache[cname] = new
        if self._bound:
            self._bound = [(ps, cname, names, ps._cache.get(cname, ent)) for ps, cname, names, ent in self._bound]
        return norm
