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
