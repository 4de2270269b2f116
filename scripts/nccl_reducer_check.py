"""2-GPU check of parallel.GradReducer over NCCL (bf16 buckets, in-collective averaging)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import modcr_b200
from modcr_b200 import parallel

rank, local, world = parallel.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
g = torch.Generator().manual_seed(5)
shapes = [(768, 768), (3072, 768), (768,), (30567, 768), (2,), (1536, 768)]
base = [torch.randn(*s, generator=g) for s in shapes]
for comm in (torch.bfloat16, None):
    params = [torch.nn.Parameter(torch.zeros(*s, device=dev)) for s in shapes]
    red = parallel.GradReducer(params, bucket_mb=8, comm_dtype=comm)
    loss = sum((p * (b.to(dev) * (rank + 1))).sum() for p, b in zip(params, base))   # d/dp = base * (rank + 1)
    loss.backward()
    red.finish()
    torch.cuda.synchronize()
    mean_factor = sum(r + 1 for r in range(world)) / world
    worst = max(float((p.grad.cpu() - b * mean_factor).abs().max() / b.abs().max()) for p, b in zip(params, base))
    tol = 1e-2 if comm is not None else 1e-6
    if rank == 0:
        print("GradReducer nccl comm=%s buckets=%d worst rel err %.2e (tol %.0e) %s" % (
            comm, red.launched, worst, tol, "OK" if worst < tol else "FAIL"), flush=True)
    assert worst < tol
    red.remove()
dist.barrier()
dist.destroy_process_group()
