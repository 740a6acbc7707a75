"""Exchange kernels alone, ranks in lock-step (their own in-kernel barriers line the ranks up): time per launch of
   the pull form (eg_allreduce_symm, multimem / peer) and of the exposed half of the push form (eg_exchange_reduce_bcast)
   on the flat gradient buffer of N Gaussians.   torchrun --nproc-per-node G scripts/exchange_micro.py [--n 500000]"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edgegaussians_b200.layout import grad_numel
from edgegaussians_b200.parallel import SymmetricExchange

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=500_000)
ap.add_argument("--reps", type=int, default=40)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
numel = grad_numel(a.n)
out = {"n": a.n, "world": world, "bytes": 4 * numel}


def timed(fn):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / a.reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]) * 1e3   # us


grids = [int(x) for x in os.environ.get("EG_MICRO_GRIDS", "148").split(",")]
for mc in (True, False):
    for grid in grids:
        ex = SymmetricExchange(numel, dev, grid=grid, multicast=mc)
        tag = ("mc" if ex.multicast_ptr else "p2p") + f"_g{grid}"
        if mc and not ex.multicast_ptr:
            continue
        out[f"pull_{tag}_us"] = timed(ex.allreduce_)
        out[f"barriers_only_{tag}_us"] = timed(lambda: ex.allreduce_(count=4 * world))   # two rank barriers, 16 bytes per rank
        ex.enable_push(a.n)
        out[f"push_reduce_bcast_{tag}_us"] = timed(ex.reduce_bcast_)
        del ex
flat = torch.zeros(numel, device=dev)
out["nccl_us"] = timed(lambda: dist.all_reduce(flat))
if rank == 0:
    print(json.dumps(out), flush=True)
dist.destroy_process_group()
