"""View-sharded multi-GPU exchange on real GPUs: needs >= 2 CUDA devices, skipped otherwise
(the host-side sharding logic is covered on CPU by tests/test_parallel_gloo.py).

  * eg_allreduce_symm (this library's kernel over symmetric memory: multimem switch reduction, and the peer
    load/store variant) == sum over ranks, repeated calls (self-resetting rank barriers), inside a CUDA graph;
  * GraphedRasterStep with the exchange captured in the graph -- the push form fused into the backward's stores
    (eg_splat_bwd_push / eg_project_bwd_push + eg_exchange_reduce_bcast), the pull form, NCCL: every rank ends with
    the SUM of the ranks' single-GPU per-view gradients (computed again, view by view, on every rank) -- the
    multi-GPU parity definition of SURVEY.md section 4 / 8e; a ragged step with idle ranks gives rank 0's gradients.
"""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N_GPUS = torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _init(rank, world, port):
    import datetime
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank),
                            timeout=datetime.timedelta(seconds=120))
    return dist


def _exchange_worker(rank, world, port, out_dir):
    dist = _init(rank, world, port)
    try:
        from edgegaussians_b200.layout import grad_numel
        from edgegaussians_b200.parallel import SymmetricExchange
        dev = torch.device("cuda", rank)
        res = {}
        for mc in (True, False):
            for n in (1001, 250_003):
                numel = grad_numel(n)
                ex = SymmetricExchange(numel, dev, multicast=mc)
                res[f"kind_{mc}"] = np.array([1 if ex.multicast_ptr else 0])
                gen = torch.Generator(device=dev).manual_seed(100 * rank + 7)
                expect = None
                for it in range(3):    # repeated calls: the in-kernel barriers reset themselves
                    mine = torch.randn(numel, generator=gen, device=dev)
                    allv = [torch.empty_like(mine) for _ in range(world)]
                    dist.all_gather(allv, mine)
                    expect = torch.stack(allv).double().sum(0)
                    ex.buf.copy_(mine)
                    ex.allreduce_()
                    torch.cuda.synchronize()
                    err = float((ex.buf.double() - expect).abs().max())
                    res[f"err_mc{int(mc)}_n{n}_it{it}"] = np.array([err, float(expect.abs().max())])
                # inside a CUDA graph, replayed twice
                src = torch.randn(numel, generator=gen, device=dev)
                allv = [torch.empty_like(src) for _ in range(world)]
                dist.all_gather(allv, src)
                expect = torch.stack(allv).double().sum(0)
                g = torch.cuda.CUDAGraph()
                torch.cuda.synchronize()
                with torch.cuda.graph(g):
                    ex.buf.copy_(src)
                    ex.allreduce_()
                for _ in range(2):
                    g.replay()
                torch.cuda.synchronize()
                res[f"graph_err_mc{int(mc)}_n{n}"] = np.array([float((ex.buf.double() - expect).abs().max()), float(expect.abs().max())])
                dist.barrier()
        np.savez(os.path.join(out_dir, f"ex{rank}.npz"), **res)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(N_GPUS < 2, reason="needs >= 2 GPUs")
@pytest.mark.timeout(240)
def test_symmetric_allreduce_equals_sum(tmp_path):
    import torch.multiprocessing as mp
    world = min(N_GPUS, 8)
    mp.spawn(_exchange_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        res = np.load(tmp_path / f"ex{r}.npz")
        for key in res.files:
            if "err" in key:
                err, scale = res[key]
                assert err <= 1e-5 * scale, (r, key, err, scale)   # fp32 sum of `world` terms, any order
        print(f"rank {r}: multicast used = {bool(res['kind_True'][0])}")


def _step_worker(rank, world, port, out_dir, exchange):
    dist = _init(rank, world, port)
    try:
        from edgegaussians_b200 import synth
        from edgegaussians_b200.cameras import OpenCVCamera
        from edgegaussians_b200.edge_gs import EdgeGaussianSplatting
        from edgegaussians_b200.graph_step import GraphedRasterStep
        dev = torch.device("cuda", rank)
        N, W, H = 30_001, 640, 480       # odd N: padded layout through the exchange
        m, q, s, o = synth.make_gaussians(N, "trained", 3, base_scale=0.006)
        vms, Ks = synth.make_cameras(world, W, H)
        gts = [synth.make_edge_map_u8(W, H, v) for v in range(world)]
        model = EdgeGaussianSplatting(device=dev)
        model.set_params(m, s, q, o, viewcams=[OpenCVCamera.from_matrices(H, W, Ks[v], vms[v]).to(dev) for v in range(world)])
        ranges = 1
        if exchange.endswith("-tiles"):    # the tile pipeline's backward (eg_project_bwd_push) instead of eg_splat_bwd_push
            exchange = exchange[:-len("-tiles")]
            model.pipeline = "tiles"
        if "-ranged" in exchange:   # e.g. "symm-ranged4": backward in 4 Gaussian ranges, exchanges on a side stream
            exchange, r = exchange.split("-ranged")
            ranges = int(r)
        step = GraphedRasterStep(model, W, H, n_slots=1, allreduce=True, exchange=exchange, exchange_ranges=ranges)
        step.set_view(0, torch.from_numpy(vms[rank]), torch.from_numpy(Ks[rank]), torch.from_numpy(gts[rank]), non_blocking=False)
        step.calibrate()
        for _ in range(3):      # graph replays: forward + backward + exchange
            ws = step.replay(0)
        torch.cuda.synchronize()
        got = ws.grads.clone()
        dist.barrier()
        # every rank recomputes all views alone (eager, no exchange) and sums them
        acc = torch.zeros_like(got, dtype=torch.float64)
        for v in range(world):
            step.set_view(0, torch.from_numpy(vms[v]), torch.from_numpy(Ks[v]), torch.from_numpy(gts[v]), non_blocking=False)
            w = step._enqueue(0, accumulate_absgrad=False)
            torch.cuda.synchronize()
            acc += w.grads.double()
        scale = float(acc.abs().max())
        err = (got.double() - acc).abs()
        bad = int((err > 1e-5 * acc.abs() + 1e-6 * scale).sum())
        # ragged last step: only rank 0 has a view, the others contribute zeros (GraphedRasterStep.idle_step)
        step.set_view(0, torch.from_numpy(vms[rank]), torch.from_numpy(Ks[rank]), torch.from_numpy(gts[rank]), non_blocking=False)
        torch.cuda.synchronize()
        dist.barrier()
        ws = step.replay(0) if rank == 0 else step.idle_step()
        torch.cuda.synchronize()
        got1 = ws.grads.clone()
        dist.barrier()
        step.set_view(0, torch.from_numpy(vms[0]), torch.from_numpy(Ks[0]), torch.from_numpy(gts[0]), non_blocking=False)
        w = step._enqueue(0, accumulate_absgrad=False)
        torch.cuda.synchronize()
        ref1 = w.grads.double()
        bad_idle = int(((got1.double() - ref1).abs() > 1e-5 * ref1.abs() + 1e-6 * float(ref1.abs().max())).sum())
        np.savez(os.path.join(out_dir, f"st{rank}.npz"), bad=np.array([bad]), err=np.array([float(err.max()), scale]),
                 factor=np.array([float(model.absgrads_normalize_factor)]), kind=np.array([step.exchange_name()]),
                 bad_idle=np.array([bad_idle]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(N_GPUS < 2, reason="needs >= 2 GPUs")
@pytest.mark.timeout(240)
@pytest.mark.parametrize("exchange", ["push", "push-p2p", "push-tiles", "symm", "symm-p2p", "nccl", "symm-ranged4", "symm-p2p-ranged3"])
def test_sharded_step_gradients_equal_sum_of_views(tmp_path, exchange):
    import torch.multiprocessing as mp
    world = min(N_GPUS, 8)
    mp.spawn(_step_worker, args=(world, _free_port(), str(tmp_path), exchange), nprocs=world, join=True)
    for r in range(world):
        res = np.load(tmp_path / f"st{r}.npz")
        print(f"rank {r}: {res['kind'][0]}: max err {res['err'][0]:.3e} of {res['err'][1]:.3e}")
        assert int(res["bad"][0]) == 0
        assert int(res["bad_idle"][0]) == 0        # a step in which only rank 0 had a view
        assert float(res["factor"][0]) == 5.0      # four steps advanced the abs-grad normaliser on every rank (idle ones too)
        if exchange.startswith("push"):
            assert "push form" in str(res["kind"][0])
        if "ranged" in exchange:
            assert "Gaussian ranges" in str(res["kind"][0])
