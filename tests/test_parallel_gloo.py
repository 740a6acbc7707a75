"""N>1 host logic on CPU: world_size-2 gloo processes exercise the view sharding and the flat-gradient
all-reduce of edgegaussians_b200/parallel.py (the GPU kernels are not involved: gradients here are a
deterministic function of the view id)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from edgegaussians_b200 import parallel
from edgegaussians_b200.layout import grad_layout, grad_numel, split_grads


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_view_grad(view, n):
    g = torch.Generator().manual_seed(1234 + view)
    flat = torch.zeros(grad_numel(n))
    for t in split_grads(flat, n):      # the pad floats between the segments are never written
        t.copy_(torch.randn(t.shape, generator=g, dtype=torch.float32))
    return flat


def _worker(rank, world, port, n, n_views, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        perm = parallel.view_permutation(n_views, epoch=3, seed=7)
        total = torch.zeros(grad_numel(n))
        absg = torch.zeros(n)
        for step in range(parallel.steps_per_epoch(n_views, world)):
            view = parallel.views_for_step(perm, step, world)[rank]
            flat = _fake_view_grad(view, n) if view is not None else torch.zeros(grad_numel(n))
            inc = torch.full((n,), float(view + 1)) if view is not None else torch.zeros(n)
            parallel.allreduce_gradients(flat, inc)
            total += flat
            absg += inc
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), total=total.numpy(), absg=absg.numpy(), perm=np.array(perm))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_views", [6, 5])
def test_view_sharded_allreduce_equals_sum_over_views(tmp_path, n_views):
    world, n = 2, 37
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, n_views, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (np.load(tmp_path / f"rank{r}.npz") for r in range(world))
    np.testing.assert_array_equal(r0["perm"], r1["perm"])           # rank-independent shuffle
    assert sorted(r0["perm"].tolist()) == list(range(n_views))
    np.testing.assert_array_equal(r0["total"], r1["total"])         # replicas stay identical
    expect = sum(_fake_view_grad(v, n) for v in range(n_views)).numpy()
    np.testing.assert_allclose(r0["total"], expect, rtol=1e-5, atol=1e-5)
    assert r0["absg"][0] == sum(v + 1 for v in range(n_views))      # every view counted exactly once


def test_views_for_step_and_layout():
    perm = [4, 2, 0, 3, 1]
    assert parallel.views_for_step(perm, 0, 2) == [4, 2]
    assert parallel.views_for_step(perm, 2, 2) == [1, None]
    assert parallel.steps_per_epoch(5, 2) == 3 and parallel.steps_per_epoch(8, 8) == 1
    for n in (5, 8, 37, 1001):            # odd N included: every segment stays 16-byte aligned
        om, os_, oq, oo, total = grad_layout(n)
        assert (om, total) == (0, grad_numel(n)) and total % 4 == 0
        assert all(o % 4 == 0 for o in (os_, oq, oo)) and os_ >= 3 * n and oq - os_ >= 3 * n and oo - oq >= 4 * n and total - oo >= n
        flat = torch.arange(total, dtype=torch.float32)
        vm, vs, vq, vo = parallel.flat_grad_views(flat, n)
        assert vm.shape == (n, 3) and vs.shape == (n, 3) and vq.shape == (n, 4) and vo.shape == (n, 1)
        assert vs[0, 0] == os_ and vq[0, 0] == oq and vo[0, 0] == oo
        vm[0, 0] = -1.0
        assert flat[0] == -1.0                                      # views, not copies
        assert (vq.data_ptr() - flat.data_ptr()) % 16 == 0          # the float4 gradient stores of the kernels
    assert parallel.view_permutation(10, 1, 3) == parallel.view_permutation(10, 1, 3)
    assert parallel.view_permutation(10, 1, 3) != parallel.view_permutation(10, 2, 3)
