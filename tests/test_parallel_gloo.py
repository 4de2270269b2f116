"""world_size-2 gloo test (CPU) of the bucketed, hook-driven gradient mean used for data parallelism."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import modcr_b200  # noqa: F401


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from modcr_b200 import parallel
    r, _, w = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4))
    unused = torch.nn.Parameter(torch.zeros(3))                      # never receives a gradient
    red = parallel.GradReducer(list(net.parameters()) + [unused], bucket_mb=0.0002)  # 209-byte buckets -> several launches
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(8, 16, generator=g)
    net(x).pow(2).mean().backward()
    local = [p.grad.clone() for p in net.parameters()]
    red.finish()
    # reference: plain all_reduce of the local grads
    ok = True
    for p, lg in zip(net.parameters(), local):
        ref = lg.clone()
        dist.all_reduce(ref)
        ok = ok and torch.allclose(p.grad, ref / world, atol=1e-7)
    ok = ok and unused.grad is None and red.launched >= 2
    # second step re-uses the reducer
    for p in net.parameters():
        p.grad = None
    net(x * 2).pow(2).mean().backward()
    red.finish()
    ok = ok and all(p.grad is not None for p in net.parameters())
    # ---- hand-off path (what functions.EncoderFn.backward does per layer) mixed with hook-driven parameters
    red.remove()
    w = [torch.nn.Parameter(torch.zeros(5, 7)), torch.nn.Parameter(torch.zeros(11)), torch.nn.Parameter(torch.zeros(3, 3))]
    red2 = parallel.GradReducer(w, bucket_mb=0.0001, comm_dtype=None)
    ok = ok and parallel.active_reducer() is red2
    gs = [torch.full_like(p, float(rank + 1) * (i + 1)) for i, p in enumerate(w)]
    class _Node(torch.autograd.Function):      # a backward node that hands its parameter gradients off itself
        @staticmethod
        def forward(ctx, x, a, b):
            ctx.ab = (a, b)
            return x * 1.0

        @staticmethod
        def backward(ctx, dy):
            a, b = ctx.ab
            red2.handoff([(a, gs[0].clone()), (b, gs[1].clone())])   # installed as .grad by the reducer
            return dy, None, None                                    # autograd still fires their hooks (ignored)
    x = torch.ones(2, requires_grad=True)
    (_Node.apply(x, w[0], w[1]).sum() + (w[2] * gs[2]).sum()).backward()   # w[2] arrives through the post-accumulate hook
    red2.finish()
    for i, p in enumerate(w):
        ok = ok and torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1)), atol=1e-6)
    # ---- gradient accumulation: the micro-step inside no_sync() only accumulates, the last one reduces the sum
    for p in w:
        p.grad = None
    before = red2.launched
    with red2.no_sync():
        red2.handoff([(w[0], gs[0].clone())])
        (w[2] * gs[2]).sum().backward()
    launched = red2.launched - before
    red2.handoff([(w[0], gs[0].clone())])
    (w[2] * gs[2]).sum().backward()
    red2.finish()
    ok = ok and launched == 0 and torch.allclose(w[0].grad, torch.full_like(w[0], 3.0), atol=1e-6)
    ok = ok and torch.allclose(w[2].grad, torch.full_like(w[2], 9.0), atol=1e-6) and w[1].grad is None
    # ---- a second synchronising backward before finish() is an error, not a silent double average
    for p in w:
        p.grad = None
    red2.handoff([(w[0], gs[0].clone())])
    try:
        red2.handoff([(w[0], gs[0].clone())])
        ok = False
    except RuntimeError:
        pass
    red2.finish()
    red2.remove()
    ok = ok and parallel.active_reducer() is None
    # ---- windows: nothing is reduced when a bucket fills, only when the backward program opens a window (or at finish)
    for p in w:
        p.grad = None
    red3 = parallel.GradReducer(w, bucket_mb=0.0001, comm_dtype=None, windows=True)
    red3.handoff([(w[0], gs[0].clone())])
    ok = ok and red3.launched == 0
    red3.open_window()
    ok = ok and red3.launched == 1
    red3.open_window()                               # nothing new: no collective
    ok = ok and red3.launched == 1
    red3.handoff([(w[1], gs[1].clone())])
    red3.handoff([(w[2], gs[2].clone())])            # does not fit behind w[1]: the full bucket goes first
    red3.finish()
    ok = ok and red3.launched == 3
    for i, p in enumerate(w):
        ok = ok and torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1)), atol=1e-6)
    red3.remove()
    loss = parallel.allreduce_mean_scalar(torch.tensor(float(rank)))
    ok = ok and abs(float(loss) - 0.5) < 1e-6
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_grad_reducer_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
