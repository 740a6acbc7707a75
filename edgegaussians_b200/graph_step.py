"""CUDA-graph execution of the fused raster iteration (one graph per view slot).

The whole iteration (one memset + 4..6 kernels, see edge_gs.enqueue_raster_step) is launch-latency
sensitive at the reference's sizes (tens of microseconds of device work per kernel), so steady-state
training replays a captured graph instead of re-issuing launches from Python.  Everything the graph
reads that changes from step to step -- camera matrices and the edge map of the view -- lives in
static per-slot device buffers that are refreshed by (async) copies before the replay.

View-sharded multi-GPU: the gradient exchange runs on kernels of this library over
symmetric memory (parallel.SymmetricExchange) and is captured INSIDE the graph, so one replay = forward +
backward + exchange.  Default ("auto" / "push"): the PUSH form -- the backward kernel itself stores every Gaussian's
gradients into its owner rank's staging slot over NVLink (the reduce-scatter rides on the backward), and one
kernel behind it sums the slots and broadcasts the result (eg_exchange_reduce_bcast).  ``"symm"`` / ``"symm-p2p"``:
the pull form, one all-reduce kernel behind the backward (eg_allreduce_symm).  ``exchange="nccl"`` keeps round 1's
torch.distributed all-reduce issued behind the replay as the A/B baseline (capturing NCCL inside the graph hangs
on this stack).
"""
from __future__ import annotations

import warnings
from typing import Dict, Optional

import torch

from . import _lib
from .edge_gs import EdgeGaussianSplatting, RasterStepWorkspace
from .engine import get_engine
from .layout import grad_numel

PIPELINES = ["splat", "tiles+splat", "tiles"]


class GraphedRasterStep:
    def __init__(self, model: EdgeGaussianSplatting, width: int, height: int, n_slots: int, gt_dtype=torch.uint8,
                 loss_weight: float = 1.0, accumulate_absgrad: bool = True, allreduce_group=None, allreduce: bool = False,
                 exchange: str = "auto", loss_mode: str = "whole", exchange_ranges: int = 1):
        dev = model.means.device
        self.model, self.W, self.H, self.n_slots = model, width, height, n_slots
        self.loss_weight, self.accumulate_absgrad = loss_weight, accumulate_absgrad
        self.loss_mode = loss_mode
        self.viewmats = torch.zeros((n_slots, 4, 4), dtype=torch.float32, device=dev)
        self.Ks = torch.zeros((n_slots, 3, 3), dtype=torch.float32, device=dev)
        self.gts = torch.zeros((n_slots, height, width), dtype=gt_dtype, device=dev)
        self.allreduce, self.allreduce_group = allreduce, allreduce_group
        if exchange not in ("auto", "push", "push-p2p", "symm", "symm-p2p", "nccl", "native-nccl"):
            raise ValueError(f"unknown exchange {exchange!r}")
        self.exchange_mode = exchange
        # Gaussian-major backward: gradients become final range by range, so with exchange_ranges > 1 the backward is
        # launched in that many Gaussian ranges and the exchange kernel of a finished range runs on a (high-priority)
        # side stream while the next range is computed -- all inside the captured graph; only the last range's
        # exchange is exposed.  (Round 1 tried this with NCCL collectives and lost: an NCCL kernel per range costs its
        # fixed latency and takes the SMs it wants; the library kernel is one small launch per range.)
        self.exchange_ranges = max(1, int(exchange_ranges))
        self.comm_stream = torch.cuda.Stream(device=dev, priority=-1) if allreduce else None
        self.ranged = False
        self.exchange = None      # parallel.SymmetricExchange (owns ws.grads) when the library kernel carries the sum
        self.native_comm = None   # parallel.NativeComm for exchange="native-nccl"
        self.graphs: Dict[int, torch.cuda.CUDAGraph] = {}
        self.ws: Optional[RasterStepWorkspace] = None
        self._capacity = None
        self._version = None      # model._resize_version the workspace / graphs were built for

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

    def _enqueue(self, slot: int, stage_cb=None, accumulate_absgrad=None, parts="all", push=None):
        acc = self.accumulate_absgrad if accumulate_absgrad is None else accumulate_absgrad
        return self.model.enqueue_raster_step(self.viewmats[slot], self.Ks[slot], self.W, self.H, self.gts[slot],
                                              loss_weight=self.loss_weight, accumulate_absgrad=acc,
                                              capacity=self._capacity, stage_cb=stage_cb, loss_mode=self.loss_mode,
                                              view_slot=slot, parts=parts, push=push)

    # ------------------------------------------------------------------ exchange
    def _setup_exchange(self) -> None:
        """(Re)build the exchange for the model's current N.  Collective over the group."""
        if not self._distributed():
            return
        dev, n = self.model.means.device, grad_numel(self.model.num_points)
        mode = self.exchange_mode
        if mode in ("auto", "push", "push-p2p", "symm", "symm-p2p"):
            if self.exchange is None or self.exchange.numel != n:
                from .parallel import SymmetricExchange
                try:
                    import torch.distributed as dist
                    # "auto": broadcast through the NVSwitch multicast object from 3 ranks up; between two GPUs plain
                    # peer stores are faster (measured: profiles/r2_exchange.md)
                    mc = not mode.endswith("-p2p") and not (mode == "auto" and dist.get_world_size(self.allreduce_group) <= 2)
                    self.exchange = SymmetricExchange(n, dev, self.allreduce_group, multicast=mc)
                    # default: the push form -- the backward's own stores carry the reduce-scatter, one reduce +
                    # broadcast kernel behind it; "symm*": the pull form (one all-reduce kernel behind the backward)
                    if mode in ("auto", "push", "push-p2p") and self.exchange_ranges <= 1:
                        self.exchange.enable_push(self.model.num_points)
                except Exception as e:  # symmetric memory unavailable on this box: say so, fall back to NCCL
                    if mode != "auto":
                        raise
                    warnings.warn(f"symmetric-memory exchange unavailable ({e!r}); using the NCCL all-reduce")
                    self.exchange, self.exchange_mode = None, "nccl"
            if self.exchange is not None:
                self.model._external_grads = self.exchange.buf   # the fused step writes its gradients into it
        elif mode == "native-nccl" and self.native_comm is None:
            from .parallel import NativeComm
            self.native_comm = NativeComm(dev, self.allreduce_group)

    def _ranged(self) -> bool:
        return (self.exchange is not None and self.exchange.push is None and self.exchange_ranges > 1
                and self.model.current_pipeline() != "tiles")

    def exchange_name(self) -> str:
        if not self._distributed():
            return "none"
        if self.exchange is not None and self.exchange.push is not None:
            bc = "multimem.st broadcast" if self.exchange.multicast_ptr else "peer-store broadcast"
            return (f"push form inside the graph: the backward stores each Gaussian's gradients into its owner rank's staging slot "
                    f"over NVLink (eg_splat_bwd_push / eg_project_bwd_push), then eg_exchange_reduce_bcast "
                    f"({bc}, {self.exchange.grid} CTAs)")
        if self.exchange is not None:
            how = (f"{self.exchange_ranges} Gaussian ranges, each exchanged on a side stream while the next range's backward runs"
                   if self.ranged else "one launch behind the backward")
            return f"eg_allreduce_symm inside the graph ({self.exchange.kind}, {self.exchange.grid} CTAs; {how})"
        if self.native_comm is not None:
            return "ncclAllReduce through libedgegs' communicator, after the replay"
        return "torch.distributed all_reduce (NCCL), after the replay"

    # ------------------------------------------------------------------ calibration / capture
    def calibrate(self, slots=None, margin: float = 1.3) -> int:
        """Eager runs (with a host read of the status words) that size the intersection capacity for
        the given slots; must be called before :meth:`capture`.  Does not touch the abs-grad statistics."""
        self.graphs.clear()
        self._setup_exchange()
        need = 0
        eng = get_engine(self.model.means.device)
        for slot in (range(self.n_slots) if slots is None else slots):
            while True:
                ws = self._enqueue(slot, accumulate_absgrad=False)
                hs = ws.status.cpu()
                n_keys = max(int(hs[_lib.EG_ST_NISECT]), int(hs[_lib.EG_ST_NKEYS]))
                need = max(need, n_keys)
                eng.max_tile = max(eng.max_tile, self.model.max_tile_load(ws, hs))
                self.model.note_status(hs, ws.T)
                if not int(hs[_lib.EG_ST_OVERFLOW]):
                    break
                self._capacity = int(n_keys * margin) + 1024
        if self._distributed():
            # every rank must run the same pipeline with the same buffer sizes: agree on the most conservative
            # pipeline and the largest capacity / tile load
            import torch.distributed as dist
            t = torch.tensor([PIPELINES.index(self.model.current_pipeline()), need, eng.max_tile], dtype=torch.int64,
                             device=self.model.means.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.allreduce_group)
            if self.model.pipeline == "auto":
                self.model._auto_pipeline = PIPELINES[int(t[0])]
            need, eng.max_tile = int(t[1]), int(t[2])
        self._capacity = max(int(need * margin) + 1024, 1 << 16)
        self.ws = self.model._workspace(self.W, self.H, self._capacity)
        self.model.install_grads(self.ws)
        self._version = self.model._resize_version
        return need

    def capture(self, slot: int, stage_cb=None) -> torch.cuda.CUDAGraph:
        if self.ws is None or self._version != self.model._resize_version:
            self.calibrate()
        g = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        self.ranged = self._ranged() and stage_cb is None
        with torch.cuda.graph(g):
            if self.ranged:
                ws = self._enqueue(slot, parts="forward")
                main, side = torch.cuda.current_stream(), self.comm_stream
                for g0, g1 in self.exchange.gaussian_ranges(ws.N, self.exchange_ranges):
                    self.model.enqueue_backward_range(ws, g0, g1)
                    ev = torch.cuda.Event()
                    ev.record(main)
                    side.wait_event(ev)
                    with torch.cuda.stream(side):
                        self.exchange.allreduce_range_(ws.N, g0, g1, grid=self.exchange.grid_ranged)
                done = torch.cuda.Event()
                done.record(side)
                main.wait_event(done)   # the optimizer / next step needs the reduced gradients
            elif self.exchange is not None and self.exchange.push is not None and stage_cb is None:
                ws = self._enqueue(slot, push=self.exchange.push)   # gradients go to the owners' staging slots ...
                self.exchange.reduce_bcast_()                       # ... and come back summed in ws.grads
            else:
                ws = self._enqueue(slot, stage_cb=stage_cb)
                if self.exchange is not None and stage_cb is None:
                    self.exchange.allreduce_()
        assert ws is self.ws, "workspace changed during capture"
        if stage_cb is None:
            self.graphs[slot] = g
        return g

    def replay(self, slot: int):
        """One iteration on the slot's view: forward + backward (+ exchange).  Re-calibrates and re-captures by itself
        after the model was resized (cull / duplicate): the old graphs hold pointers to the old parameters."""
        if self._version != self.model._resize_version:
            self.calibrate()
        g = self.graphs.get(slot)
        if g is None:
            g = self.capture(slot)
        g.replay()
        if self._distributed() and self.exchange is None:
            if self.native_comm is not None:
                self.native_comm.allreduce_(self.ws.grads)
            else:
                import torch.distributed as dist
                dist.all_reduce(self.ws.grads, group=self.allreduce_group)
        m = self.model
        if self.accumulate_absgrad:
            m.absgrads_normalize_factor += 1   # update_absgrads bookkeeping (edge_gs.py:607-613), once per step
        m.step += 1
        m.last_size = (self.H, self.W)
        return self.ws

    def idle_step(self):
        """A rank without a view in a ragged last step: zero gradients, same exchange, same bookkeeping."""
        if self._version != self.model._resize_version:
            self.calibrate()
        self.ws.grads.zero_()
        if self._distributed():
            if self.exchange is not None and self.exchange.push is not None:
                self.exchange.push_zero_()
                self.exchange.reduce_bcast_()
            elif self._ranged():   # the same launches (grid, barriers) as the working ranks' captured graphs
                for g0, g1 in self.exchange.gaussian_ranges(self.ws.N, self.exchange_ranges):
                    self.exchange.allreduce_range_(self.ws.N, g0, g1, grid=self.exchange.grid_ranged)
            elif self.exchange is not None:
                self.exchange.allreduce_()
            elif self.native_comm is not None:
                self.native_comm.allreduce_(self.ws.grads)
            else:
                import torch.distributed as dist
                dist.all_reduce(self.ws.grads, group=self.allreduce_group)
        if self.accumulate_absgrad:
            self.model.absgrads_normalize_factor += 1
        self.model.step += 1
        return self.ws

    def poll_policy(self) -> bool:
        """Host read of the last step's status words (one small D2H copy: call it every few hundred steps, not per
        step).  Feeds the model's pipeline policy and checks the overflow word: when the policy moves (e.g. the
        scene has become opaque enough that most tiles need the sorted fallback) or a key buffer overflowed (the
        step's kernels were no-ops and its gradients stale), the workspace is re-calibrated and the graphs are
        re-captured on next use.  Returns True when that happened.  In a distributed run every rank must call it at
        the same step; the decision is agreed with a MAX reduction."""
        before = self.model.current_pipeline()
        hs = self.ws.status.cpu()
        self.model.note_status(hs, self.ws.T)
        overflow = int(hs[_lib.EG_ST_OVERFLOW])
        code = torch.tensor([PIPELINES.index(self.model.current_pipeline()), overflow], dtype=torch.int64,
                            device=self.ws.status.device)
        if self._distributed():
            import torch.distributed as dist
            dist.all_reduce(code, op=dist.ReduceOp.MAX, group=self.allreduce_group)
            if self.model.pipeline == "auto":
                self.model._auto_pipeline = PIPELINES[int(code[0])]
        changed = self.model.current_pipeline() != before or int(code[1]) != 0
        if changed:
            if int(code[1]):
                self._capacity = None   # let calibrate() find the new size
            self.calibrate()
        return changed

    def loss(self) -> torch.Tensor:
        return self.model.loss_from_workspace(self.ws, self.loss_mode)
