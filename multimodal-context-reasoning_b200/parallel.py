"""Data-parallel plumbing of the hot path: one process per GPU, ``torch.distributed`` (NCCL over
NVLink 5 / NVSwitch on the B200 box, gloo in CPU tests), bucketed gradient all-reduce overlapped
with backward.

The reference never shards (NCCL is initialised and then unused: run_PMR_ModCR.py:438-448, :693);
the only bucketing logic in its tree is the dead Horovod helper utils/distributed.py:46-97 (10 MiB flat
buckets).  The path shards by *questions*: every rank draws its own batch, runs the full replica and
the only exchange step is the gradient mean -- so there is exactly one collective.

Buckets are filled in gradient-arrival order -- layer by layer from inside the encoder's backward program
(``GradReducer.handoff``), from ``register_post_accumulate_grad_hook`` for the heads -- and all-reduced on a side
stream while the remaining backward kernels run.  Parameters that never receive a gradient
never enter a bucket, so nothing waits for them.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise the default process group from RANK / WORLD_SIZE / MASTER_* (torchrun contract)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            # the collectives share the SMs with the backward kernels: cap their CTAs (GradReducer reserves as many SMs)
            if COMM_SMS > 0:
                os.environ.setdefault("NCCL_MAX_CTAS", str(COMM_SMS))
            torch.cuda.set_device(local)
            # the collective's CTAs are placed ahead of the queued CTAs of the backward kernel they run beside
            # (GradReducer.open_window)
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=COMM_HIGH_PRIORITY)
            dist.init_process_group(backend, device_id=torch.device("cuda", local), pg_options=opts)
        else:
            dist.init_process_group(backend)
    return rank, local, world


# SMs the persistent kernels leave to the NCCL CTAs while gradient buckets are in flight (0 = off, the default: measured at
# 8 GPUs, cfg-2 -- 38.3 ms / step with NCCL's own CTA count and full-width kernels, 39.2 / 40.6 / 41.6 ms with 4 / 8 / 16
# SMs reserved and NCCL capped to as many CTAs, profiles/r02_bench_cfg2_n8_comm*.log)
COMM_SMS = int(os.environ.get("MODCR_COMM_SMS", "0"))
COMM_HIGH_PRIORITY = os.environ.get("MODCR_COMM_PRIORITY", "1") != "0"
_SKIP_COLLECTIVE = os.environ.get("MODCR_DEBUG_SKIP_COLLECTIVE", "0") == "1"

_ACTIVE = None


def active_reducer():
    """The reducer the backward programs hand their gradients to (``None`` on a single GPU)."""
    return _ACTIVE


class GradReducer:
    """Bucketed gradient mean over the data-parallel group, overlapped with backward LAYER BY LAYER.

    Two producers feed the buckets:

    * ``handoff`` -- called from inside the coarse backward nodes (``functions.EncoderFn.backward``) as soon as the
      weight gradients of one encoder layer have been enqueued, in reverse layer order.  The reducer installs the tensors
      as ``p.grad`` itself (the node returns ``None`` for them), so what it averages in place is exactly what the
      optimiser reads.  Only the gradients of the last layer processed plus the embeddings are still in flight when
      backward returns.
    * ``register_post_accumulate_grad_hook`` for every other parameter (heads, prefix, scorer).

    Per bucket, on a side stream: ONE ``modcr_multi_copy`` launch per hand-off gathers (and casts to ``comm_dtype``) the
    gradients into the flat bucket, NCCL averages inside the collective, ONE launch per 64 tensors scatters the means
    back.  Parameters that never receive a gradient never enter a bucket.

    ``windows=True``: the collectives are not issued when a bucket fills but when the backward program calls
    ``open_window()`` -- right before a long kernel with many more CTAs than SMs (the attention backward of the next
    layer), on high-priority streams.  A collective kernel holds its SMs until every rank has arrived; the persistent GEMMs
    partition their tiles statically over all SMs, so one that starts beside a collective cannot finish before it (the
    CTAs that found no SM run last).  Beside a many-CTA kernel the collective only costs its share of the SMs.

    Gradient accumulation (run_PMR_ModCR.py:211-227): wrap every micro-step but the last in ``no_sync()``; the last one
    reduces the accumulated ``.grad``.  ``finish()`` must follow exactly one synchronising backward: a parameter that is
    enqueued twice before ``finish()`` raises."""

    def __init__(self, params, bucket_mb=128, group=None, comm_dtype=None, handoff=True, windows=False):
        """``comm_dtype=torch.bfloat16`` halves the bytes on NVLink (the sum is then formed in bf16, like DDP's bf16
        compression hook); ``None`` reduces in the gradients' own dtype (fp32)."""
        global _ACTIVE
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self._mine = {id(p) for p in self.params}
        self.bucket_bytes = int(bucket_mb * (1 << 20))
        self.comm_dtype = comm_dtype
        self.sync = True
        self.windows = bool(windows) and handoff
        self._seen = {}
        self._bucket = None            # dict(flat=, used=, grads=[...]) being filled
        self._flats, self._free = [], []   # flat buffers, recycled step after step (stable addresses for graph replay)
        self._pending = []             # gloo: (work, flat, used, grads)
        self._handles = []
        self.stream = None
        self.launched = 0
        self._sm_reserved = False
        if self.world > 1:
            for p in self.params:
                self._handles.append(p.register_post_accumulate_grad_hook(self._on_grad))
            if handoff:
                _ACTIVE = self

    # ---------------------------------------------------------------- producers
    def no_sync(self):
        """Context manager of the accumulation micro-steps that do not end in a reduction."""
        return _NoSync(self)

    def owns(self, p):
        return id(p) in self._mine

    def handoff(self, pairs):
        """``pairs``: (parameter, gradient tensor) of one layer, complete on the current stream.  Installs / accumulates
        ``p.grad`` and, unless inside ``no_sync()``, enqueues the gradients for reduction."""
        ps = []
        for p, g in pairs:
            if p.grad is None:
                p.grad = g
            else:
                p.grad.add_(g)
            ps.append(p)
        if self.sync and self.world > 1:
            self._enqueue(ps, "hand-off")

    def _on_grad(self, p):
        if p.grad is None or not self.sync:
            return
        # autograd still runs AccumulateGrad (and this hook) for an input whose backward returned None: a parameter the
        # node handed off itself is already on its way
        if self._seen.get(id(p)) == "hand-off":
            return
        self._enqueue([p])

    # ---------------------------------------------------------------- buckets
    def _side(self, cuda):
        if not cuda:
            return _Null()
        if self.stream is None:
            self.stream = torch.cuda.Stream(priority=-1 if COMM_HIGH_PRIORITY else 0)
        ev = torch.cuda.Event()
        ev.record()                      # these gradients are complete on the compute stream here
        self.stream.wait_event(ev)
        return torch.cuda.stream(self.stream)

    def _new_flat(self, numel, like):
        dt = self.comm_dtype if self.comm_dtype is not None else like.dtype
        for i, f in enumerate(self._free):
            if f.numel() >= numel and f.dtype == dt and f.device == like.device:
                return self._free.pop(i)
        f = torch.empty(numel, dtype=dt, device=like.device)
        self._flats.append(f)
        return f

    def _enqueue(self, ps, via="hook"):
        for p in ps:
            if id(p) in self._seen:
                raise RuntimeError("GradReducer: a parameter (shape %s, first enqueued by %s, now by %s) was enqueued twice "
                                   "before finish() -- finish() must follow exactly one synchronising backward (use "
                                   "no_sync() for accumulation micro-steps)" % (tuple(p.shape), self._seen[id(p)], via))
            self._seen[id(p)] = via
        grads = [p.grad for p in ps]
        need = sum((g.numel() + 7) & ~7 for g in grads)          # 16-byte aligned slots
        cuda = grads[0].is_cuda
        if cuda and not self._sm_reserved and COMM_SMS > 0:
            # from the first bucket of the step until finish(): the persistent GEMM / attention kernels leave COMM_SMS
            # SMs to the collectives instead of queueing a whole kernel behind them
            from . import _lib
            _lib.set_sm_limit(torch.cuda.get_device_properties(grads[0].device).multi_processor_count - COMM_SMS)
            self._sm_reserved = True
        with self._side(cuda):
            b = self._bucket
            esz = 2 if self.comm_dtype in (torch.bfloat16, torch.float16) else grads[0].element_size()
            cap = max(self.bucket_bytes // esz, 8)
            if b is not None and b["used"] + need > b["flat"].numel():
                self._launch()
                b = None
            if b is None:
                b = self._bucket = dict(flat=self._new_flat(max(cap, need), grads[0]), used=0, grads=[], views=[])
            views = []
            for g in grads:
                views.append(b["flat"][b["used"]:b["used"] + g.numel()].view_as(g))
                b["used"] += (g.numel() + 7) & ~7
            _copy_many(grads, views)
            b["grads"] += grads
            b["views"] += views
            if b["used"] * esz >= self.bucket_bytes and not self.windows:
                self._launch()

    def open_window(self):
        """Called by a backward program right before it launches a long many-CTA kernel: issue the collective of what has
        been handed off so far (``windows=True`` only), ordered behind the compute stream *here*."""
        b = self._bucket
        if not self.windows or b is None or not b["grads"]:
            return
        with self._side(b["flat"].is_cuda):
            self._launch()

    def _launch(self):
        """All-reduce the bucket being filled (caller holds the side-stream context)."""
        b, self._bucket = self._bucket, None
        if b is None or not b["grads"]:
            return
        flat = b["flat"][:b["used"]]
        cuda = flat.is_cuda
        avg = cuda and dist.get_backend(self.group) == "nccl"    # NCCL averages inside the collective
        if _SKIP_COLLECTIVE:              # timing diagnostic only (gather + scatter without the exchange): wrong gradients
            work = None
        else:
            work = dist.all_reduce(flat, op=dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM, group=self.group,
                                   async_op=True)
        self.launched += 1
        if cuda:
            if work is not None:
                work.wait()               # stream-level: the side stream waits for NCCL, the host does not
            _copy_many(b["views"], b["grads"], 1.0 if avg else 1.0 / self.world)
            self._free.append(b["flat"])  # stream-ordered reuse on the side stream
        else:
            self._pending.append((work, b))

    # ---------------------------------------------------------------- per step
    def finish(self):
        """Flush the partial bucket and make the compute stream wait for every reduction; afterwards ``.grad`` holds the
        mean over ranks."""
        if self.world <= 1:
            self._seen.clear()
            return
        cuda = self.stream is not None
        if self._bucket is not None:
            with (torch.cuda.stream(self.stream) if self._bucket["flat"].is_cuda else _Null()):
                self._launch()
        for work, b in self._pending:
            work.wait()
            _copy_many(b["views"], b["grads"], 1.0 / self.world)
            self._free.append(b["flat"])
        self._pending = []
        if cuda:
            torch.cuda.current_stream().wait_stream(self.stream)
        if self._sm_reserved:
            from . import _lib
            _lib.set_sm_limit(0)
            self._sm_reserved = False
        self._seen.clear()

    def remove(self):
        global _ACTIVE
        for h in self._handles:
            h.remove()
        self._handles = []
        if _ACTIVE is self:
            _ACTIVE = None


def _copy_many(srcs, dsts, scale=1.0):
    """dst[i] = scale * src[i].  CUDA: ``modcr_multi_copy`` (csrc/optim.cu, one launch per 64 tensors).  Host tensors
    (the world-size-2 gloo tests of this plumbing) use torch's multi-tensor copy."""
    if srcs[0].is_cuda:
        from . import ops
        ops.multi_copy([s.contiguous() for s in srcs], dsts, scale)
        return
    torch._foreach_copy_(dsts, srcs)
    if scale != 1.0:
        torch._foreach_mul_(dsts, scale)


class _NoSync:
    def __init__(self, red):
        self.red = red

    def __enter__(self):
        self.prev, self.red.sync = self.red.sync, False
        return self.red

    def __exit__(self, *a):
        self.red.sync = self.prev
        return False


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def allreduce_mean_scalar(t, group=None):
    """Mean of a scalar tensor over ranks (loss reporting)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return t
    t = t.detach().clone()
    dist.all_reduce(t, group=group)
    return t / dist.get_world_size(group)
