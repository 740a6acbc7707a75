"""CUDA-graph execution of the fused raster iteration (one graph per view slot).

The whole iteration (3 memsets + 6 kernels, see edge_gs.enqueue_raster_step) is launch-latency
sensitive at the reference's sizes (tens of microseconds of device work per kernel), so steady-state
training replays a captured graph instead of re-issuing launches from Python.  Everything the graph
reads that changes from step to step -- camera matrices and the edge map of the view -- lives in
static per-slot device buffers that are refreshed by (async) copies before the replay.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .edge_gs import EdgeGaussianSplatting, RasterStepWorkspace
from .engine import get_engine


class GraphedRasterStep:
    def __init__(self, model: EdgeGaussianSplatting, width: int, height: int, n_slots: int, gt_dtype=torch.uint8,
                 loss_weight: float = 1.0, accumulate_absgrad: bool = True, allreduce_group=None, allreduce: bool = False,
                 allreduce_chunks: int = 1, native_allreduce: bool = False):
        dev = model.means.device
        self.model, self.W, self.H, self.n_slots = model, width, height, n_slots
        self.loss_weight, self.accumulate_absgrad = loss_weight, accumulate_absgrad
        self.viewmats = torch.zeros((n_slots, 4, 4), dtype=torch.float32, device=dev)
        self.Ks = torch.zeros((n_slots, 3, 3), dtype=torch.float32, device=dev)
        self.gts = torch.zeros((n_slots, height, width), dtype=gt_dtype, device=dev)
        # view-sharded data parallelism: the NCCL all-reduce of the flat gradient buffer is issued on the same
        # stream right behind the replay.  (Capturing the collective inside the graph was tried and hung on this
        # stack -- torch 2.11 / NCCL 2.28.9 -- so it is not done.)
        self.allreduce, self.allreduce_group = allreduce, allreduce_group
        self.allreduce_in_graph = False
        # Gaussian-major backward: gradients become final range by range, so the backward is launched in
        # `allreduce_chunks` Gaussian ranges (outside the graph) and the all-reduce of a finished range runs on a
        # side stream while the next range is computed; only the last range's collective is exposed.
        self.allreduce_chunks = max(1, int(allreduce_chunks))
        self.comm_stream = torch.cuda.Stream(device=dev) if allreduce else None
        self.chunked = False
        self.native_comm = None   # parallel.NativeComm: ranged backward + collectives as one C call
        self.python_ranges = False  # True: drive the ranged backward + torch.distributed collectives from Python (A/B)
        # opt-in: the one all-reduce per step through libedgegs' own communicator, enqueued straight on the compute
        # stream (torch.distributed routes through an internal stream; measured identical at 2 GPUs: 0.379 ms)
        self.native_allreduce = native_allreduce
        self.graphs: Dict[int, torch.cuda.CUDAGraph] = {}
        self.ws: Optional[RasterStepWorkspace] = None
        self._capacity = None

    def set_view(self, slot: int, viewmat: torch.Tensor, K: torch.Tensor, gt: torch.Tensor, non_blocking=True):
        """Refresh a slot from host (pinned) or device tensors."""
        self.viewmats[slot].copy_(viewmat.reshape(4, 4), non_blocking=non_blocking)
        self.Ks[slot].copy_(K.reshape(3, 3), non_blocking=non_blocking)
        self.gts[slot].copy_(gt, non_blocking=non_blocking)

    def _distributed(self) -> bool:
        if not self.allreduce:
            return False
        import torch.distributed as dist
        return dist.is_initialized() and dist.get_world_size(self.allreduce_group) > 1

    def _enqueue(self, slot: int, stage_cb=None, parts="all"):
        return self.model.enqueue_raster_step(self.viewmats[slot], self.Ks[slot], self.W, self.H, self.gts[slot],
                                              loss_weight=self.loss_weight, accumulate_absgrad=self.accumulate_absgrad,
                                              capacity=self._capacity, stage_cb=stage_cb, parts=parts)

    def calibrate(self, slots=None, margin: float = 1.3) -> int:
        """Eager runs (with a host read of the status words) that size the intersection capacity for
        the given slots; must be called before :meth:`capture`."""
        need = 0
        for slot in (range(self.n_slots) if slots is None else slots):
            while True:
                ws = self._enqueue(slot)
                hs = ws.status.cpu()
                n_isects = int(hs[0])
                need = max(need, n_isects)
                eng = get_engine(self.model.means.device)
                eng.max_tile = max(eng.max_tile, self.model.max_tile_load(ws, hs))
                self.model.note_status(hs, ws.T)
                if not int(hs[1]):
                    break
                self._capacity = int(n_isects * margin) + 1024
        if self._distributed():
            # every rank must run the same pipeline: the collective pattern (one all-reduce vs. ranged ones) and
            # the buffer sizes follow from it.  Agree on the most conservative choice and the largest capacity.
            import torch.distributed as dist
            order = ["splat", "tiles+splat", "tiles"]
            t = torch.tensor([order.index(self.model.current_pipeline()), need], dtype=torch.int64,
                             device=self.model.means.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.allreduce_group)
            if self.model.pipeline == "auto":
                self.model._auto_pipeline = order[int(t[0])]
            need = int(t[1])
        self._capacity = max(int(need * margin) + 1024, 1 << 16)
        self.ws = self.model._workspace(self.W, self.H, self._capacity)
        self.model.install_grads(self.ws)
        self.graphs.clear()
        return need

    def capture(self, slot: int, stage_cb=None) -> torch.cuda.CUDAGraph:
        if self.ws is None:
            self.calibrate([slot])
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        self.chunked = (self._distributed() and self.allreduce_chunks > 1
                        and self.model.current_pipeline() != "tiles" and stage_cb is None)
        if (self._distributed() and self.native_comm is None and not self.python_ranges
                and (self.chunked or self.native_allreduce)):
            from .parallel import NativeComm
            self.native_comm = NativeComm(self.model.means.device, self.allreduce_group)
        with torch.cuda.graph(g):
            ws = self._enqueue(slot, stage_cb=stage_cb, parts="forward" if self.chunked else "all")
        assert ws is self.ws, "workspace changed during capture"
        if stage_cb is None:
            self.graphs[slot] = g
        return g

    def replay(self, slot: int):
        g = self.graphs.get(slot)
        if g is None:
            g = self.capture(slot)
        g.replay()
        if self._distributed():
            import torch.distributed as dist
            from . import parallel
            ws = self.ws
            if self.chunked and self.native_comm is not None:
                self.model.enqueue_backward_allreduce(ws, self.native_comm, self.allreduce_chunks)
            elif self.chunked:
                main = torch.cuda.current_stream()
                for g0, g1 in parallel.gaussian_ranges(ws.N, self.allreduce_chunks):
                    self.model.enqueue_backward_range(ws, g0, g1)
                    ev = torch.cuda.Event()
                    ev.record(main)
                    self.comm_stream.wait_event(ev)
                    with torch.cuda.stream(self.comm_stream):
                        parallel.allreduce_range(ws.grads, ws.N, g0, g1, self.allreduce_group)
                done = torch.cuda.Event()
                done.record(self.comm_stream)
                main.wait_event(done)   # the optimizer / next step needs the reduced gradients
            elif self.native_comm is not None and self.native_allreduce:
                self.native_comm.allreduce_(ws.grads)
            else:
                dist.all_reduce(ws.grads, group=self.allreduce_group)
        return self.ws

    def poll_policy(self) -> bool:
        """Host read of the last step's status words (one small D2H copy: call it every few hundred steps, not per
        step).  Feeds the model's pipeline policy; when the policy moves (e.g. the scene has become opaque enough
        that most tiles need the sorted fallback) the captured graphs are dropped and re-captured on next use.
        Returns True when that happened.  In a distributed run every rank must call it at the same step."""
        before = self.model.current_pipeline()
        hs = self.ws.status.cpu()
        self.model.note_status(hs, self.ws.T)
        code = torch.tensor([["splat", "tiles+splat", "tiles"].index(self.model.current_pipeline())],
                            dtype=torch.int64, device=self.ws.status.device)
        if self._distributed():
            import torch.distributed as dist
            dist.all_reduce(code, op=dist.ReduceOp.MAX, group=self.allreduce_group)
            if self.model.pipeline == "auto":
                self.model._auto_pipeline = ["splat", "tiles+splat", "tiles"][int(code[0])]
        changed = self.model.current_pipeline() != before
        if changed:
            self.graphs.clear()
        return changed

    def loss(self) -> torch.Tensor:
        return (self.ws.loss_sum[0] / float(self.W * self.H)).float()
