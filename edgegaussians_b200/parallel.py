"""View-sharded data parallelism for the raster iteration.

The reference is strictly single-GPU, one view per step (train_gaussians.py:71,311).  The path shards
naturally by VIEW: parameters are replicated, rank r renders view ``perm[step * G + r]`` and the
per-view gradients (additive over views) are summed with ONE all-reduce over the flat fp32 gradient
buffer laid out  means | scales | quats | opacities  (11 N floats) -- optionally followed by the [N]
abs-grad statistics.  One process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def view_permutation(n_views: int, epoch: int, seed: int = 0) -> List[int]:
    """Rank-independent shuffle of the views of one epoch (the reference shuffles with an unseeded
    DataLoader, train_gaussians.py:311; every rank must draw the same permutation)."""
    g = torch.Generator()
    g.manual_seed(seed * 1_000_003 + epoch)
    return torch.randperm(n_views, generator=g).tolist()


def views_for_step(perm: Sequence[int], step: int, world_size: int) -> List[Optional[int]]:
    """The views rendered at ``step`` by ranks 0..G-1 (None = the rank idles: ragged last step)."""
    base = step * world_size
    return [perm[base + r] if base + r < len(perm) else None for r in range(world_size)]


def steps_per_epoch(n_views: int, world_size: int) -> int:
    return (n_views + world_size - 1) // world_size


def flat_grad_views(flat: torch.Tensor, n: int):
    """(v_means [N,3], v_scales [N,3], v_quats [N,4], v_opacities [N,1]) views of the flat buffer."""
    return (flat[0:3 * n].view(n, 3), flat[3 * n:6 * n].view(n, 3), flat[6 * n:10 * n].view(n, 4),
            flat[10 * n:11 * n].view(n, 1))


def allreduce_gradients(flat: torch.Tensor, absgrad_increment: Optional[torch.Tensor] = None, group=None,
                        average: bool = False) -> None:
    """Sum (or average) the per-view gradients of all ranks in place.  An idle rank passes zeros."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if absgrad_increment is not None:
        dist.all_reduce(absgrad_increment, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat.div_(dist.get_world_size(group))


def gaussian_ranges(n: int, chunks: int, align: int = 128) -> List[tuple]:
    """Split [0, n) into at most ``chunks`` contiguous ranges whose boundaries are multiples of ``align`` (the
    Gaussians one CTA of eg_splat_bwd owns)."""
    chunks = max(1, int(chunks))
    per = -(-n // chunks)
    per = -(-per // align) * align
    return [(b, min(n, b + per)) for b in range(0, n, per)]


def range_slices(flat: torch.Tensor, n: int, g0: int, g1: int):
    """The four slices of the flat gradient buffer (means | scales | quats | opacities) that hold the
    Gaussians [g0, g1)."""
    return [flat[3 * g0:3 * g1], flat[3 * n + 3 * g0:3 * n + 3 * g1], flat[6 * n + 4 * g0:6 * n + 4 * g1],
            flat[10 * n + g0:10 * n + g1]]


def allreduce_range(flat: torch.Tensor, n: int, g0: int, g1: int, group=None) -> None:
    """Sum the gradients of the Gaussians [g0, g1) over all ranks in place (issued on the current stream).  The
    Gaussian-major backward finishes its gradients range by range, so the collective of one range runs while
    the next range is still being computed.  On NCCL the four slices go out as one grouped launch."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    parts = range_slices(flat, n, g0, g1)
    if flat.is_cuda and hasattr(dist, "_coalescing_manager"):
        try:
            with dist._coalescing_manager(group=group, device=flat.device):
                for t in parts:
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            return
        except (RuntimeError, TypeError, NotImplementedError):
            pass
    for t in parts:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


class NativeComm:
    """Communicator of libedgegs.so (eg_comm_*, NCCL underneath) over the ranks of ``group``: what
    ``eg_splat_bwd_allreduce`` exchanges gradients through.  torch.distributed only carries the 128-byte
    rendezvous id.  Collective: every rank of the group must construct it at the same point."""

    def __init__(self, device: torch.device, group=None):
        import ctypes
        from . import _lib
        self.lib = _lib.load()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        id_t = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (ctypes.c_char * 128)()
            _lib.check(self.lib.eg_comm_unique_id(ctypes.cast(buf, ctypes.c_void_p)), "eg_comm_unique_id")
            id_t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        carrier = id_t.to(device) if dist.get_backend(group) == "nccl" else id_t
        dist.broadcast(carrier, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = bytes(carrier.cpu().numpy().tobytes())
        self.handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(self.lib.eg_comm_init(ctypes.c_char_p(raw), self.rank, self.world, ctypes.byref(self.handle)),
                       "eg_comm_init")
        # high priority: the collective of a finished range must get SMs while the next range's backward is running
        self.stream = torch.cuda.Stream(device=device, priority=-1)

    def allreduce_(self, flat: torch.Tensor) -> None:
        """In-place fp32 sum over the ranks, enqueued directly on the current stream."""
        import ctypes
        from . import _lib
        _lib.check(self.lib.eg_comm_allreduce(ctypes.c_void_p(flat.data_ptr()), flat.numel(), self.handle,
                                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "eg_comm_allreduce")

    def close(self):
        if getattr(self, "handle", None):
            self.lib.eg_comm_destroy(self.handle)
            self.handle = None


def sync_absgrads(model, group=None) -> None:
    """Sum the per-rank abs-grad statistics (model.absgrads, accumulated locally by the fused step) over all
    ranks.  Needed only where the reference reads them: at densification (edge_gs.py:544-576), i.e. once per
    epoch boundary, not per step."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(model.absgrads, op=dist.ReduceOp.SUM, group=group)


class ViewShardedStep:
    """Drives ``EdgeGaussianSplatting.raster_step`` on this rank's view and all-reduces the result.

    ``model`` holds replicated parameters; ``gts`` maps view id -> device edge map.  The abs-grad
    statistics are accumulated from all views (the reference accumulates one view per step)."""

    def __init__(self, model, gts, group=None):
        self.model, self.gts, self.group = model, gts, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def step(self, perm: Sequence[int], step: int, loss_weight: float = 1.0) -> torch.Tensor:
        view = views_for_step(perm, step, self.world)[self.rank]
        model = self.model
        before = model.absgrads.clone()
        if view is None:
            ws = model._ws
            ws.grads.zero_()
            loss = torch.zeros((), device=ws.grads.device)
        else:
            loss = model.raster_step(view, self.gts[view], loss_weight=loss_weight)
        ws = model._ws
        inc = model.absgrads - before
        allreduce_gradients(ws.grads, inc, self.group)
        model.absgrads.copy_(before + inc)
        if self.world > 1:
            dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=self.group)
        return loss
