"""View-sharded data parallelism for the raster iteration.

The reference is strictly single-GPU, one view per step (train_gaussians.py:71,311).  The path shards
naturally by VIEW: parameters are replicated, rank r renders view ``perm[step * G + r]`` and the
per-view gradients (additive over views) are summed with ONE all-reduce over the flat fp32 gradient
buffer laid out  means | scales | quats | opacities  (layout.grad_layout: 11 x N rounded up to a multiple of 4 floats) -- optionally followed by the [N]
abs-grad statistics.  One process per GPU, ``torch.distributed`` (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from .layout import grad_layout, grad_numel, split_grads


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
    vm, vs, vq, vo = split_grads(flat, n)
    return vm, vs, vq, vo.view(n, 1)


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


class SymmetricExchange:
    """The gradient exchange of the view-sharded step through libedgegs' own kernel (eg_allreduce_symm).

    Owns a SYMMETRIC fp32 buffer of ``numel`` floats (``.buf``; the fused step writes its gradients straight into
    it -- it is handed to the workspace as ``ws.grads``) plus the flag area of the in-kernel rank barriers.
    torch.distributed._symmetric_memory supplies the allocation, the peer mappings and the NVSwitch multicast
    mapping (plumbing); the data path is one kernel of this library: switch-side ``multimem.ld_reduce`` of this
    rank's slice + ``multimem.st`` broadcast (or 128-bit peer loads / stores where the fabric has no multicast
    object).  ``allreduce_()`` only enqueues that kernel on the current stream: no host sync, CUDA-graph capturable.
    Collective: every rank of the group constructs it at the same point."""

    def __init__(self, numel: int, device: torch.device, group=None, grid: int = 0, multicast: bool = True):
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self.lib = _lib.load()
        group = group if group is not None else dist.group.WORLD
        self._group = group
        self.push = None       # eg_push_target once enable_push() has run
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.numel = int(numel)
        assert self.numel % 4 == 0, "the flat gradient buffer is padded to 16 bytes (layout.grad_numel)"
        # one CTA per SM by default (nothing else runs at the tail of a step); EG_AR_GRID overrides for tuning
        self.grid = int(grid) if grid > 0 else int(os.environ.get("EG_AR_GRID", "148"))
        # the ranged exchange shares the SMs with the backward of the next range: fewer CTAs
        self.grid_ranged = min(self.grid, int(os.environ.get("EG_AR_GRID_RANGED", "74")))
        self.buf = symm_mem.empty(self.numel, dtype=torch.float32, device=device)
        self.buf.zero_()
        self._hdl = symm_mem.rendezvous(self.buf, group)
        n_flags = int(self.lib.eg_allreduce_flag_words(self.grid))
        self.flags = symm_mem.empty(n_flags, dtype=torch.int32, device=device)
        self.flags.zero_()
        self._fhdl = symm_mem.rendezvous(self.flags, group)
        torch.cuda.synchronize(device)
        dist.barrier(group)   # every rank's flags are zero before the first kernel signals anybody

        def ptrs(hdl, local):
            # the handle reports the BASE of each rank's block; the tensor may sit at an offset inside it
            delta = local.data_ptr() - int(hdl.buffer_ptrs[self.rank])
            return (ctypes.c_void_p * self.world)(*[int(p) + delta for p in hdl.buffer_ptrs]), delta
        (self._bufs, delta), (self._flagp, _) = ptrs(self._hdl, self.buf), ptrs(self._fhdl, self.flags)
        mc = int(getattr(self._hdl, "multicast_ptr", 0) or 0) if multicast else 0
        if mc:
            mc += delta
        self.multicast_ptr = mc
        self.kind = "multimem (switch-side reduction)" if mc else "peer loads/stores"

    # ------------------------------------------------------------------ push form (eg_push_target)
    def enable_push(self, n_gaussians: int) -> None:
        """Allocate the symmetric staging area of the push form for ``n_gaussians`` (collective): afterwards
        ``self.push`` is the eg_push_target the backward kernels store through (eg_splat_bwd_push /
        eg_project_bwd_push) and :meth:`reduce_bcast_` finishes the sum.  The backward's gradient stores then ARE
        the reduce-scatter: they travel over NVLink while the backward is still running."""
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        n = int(n_gaussians)
        assert grad_numel(n) == self.numel, "the exchange was sized for another Gaussian count"
        per = int(self.lib.eg_exchange_push_per(n, self.world))
        floats = int(self.lib.eg_exchange_stage_floats(n, self.world))
        self.stage = symm_mem.empty(floats, dtype=torch.float32, device=self.buf.device)
        self.stage.zero_()     # rows past N in the last owner's range are never written: they must add zeros
        self._shdl = symm_mem.rendezvous(self.stage, self._group)
        torch.cuda.synchronize(self.buf.device)
        dist.barrier(self._group)
        delta = self.stage.data_ptr() - int(self._shdl.buffer_ptrs[self.rank])
        tgt = _lib.EgPushTarget()
        for r in range(self.world):
            tgt.stage[r] = int(self._shdl.buffer_ptrs[r]) + delta
        tgt.per, tgt.rank, tgt.world = per, self.rank, self.world
        self.push, self.push_n = tgt, n

    def reduce_bcast_(self) -> None:
        """Second half of the push form, enqueued on the current stream behind the pushing backward: rank barrier, sum
        of this rank's ``world`` local slots, broadcast into every rank's ``buf`` (multimem.st / peer stores)."""
        import ctypes
        from . import _lib
        _lib.check(self.lib.eg_exchange_reduce_bcast(ctypes.byref(self.push), self._bufs, ctypes.c_void_p(self.multicast_ptr or None),
                                                     self._flagp, self.push_n, self.grid,
                                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "eg_exchange_reduce_bcast")

    def push_zero_(self) -> None:
        """What a rank without a view does instead of a pushing backward: zeros into its slot of every owner."""
        import ctypes
        from . import _lib
        _lib.check(self.lib.eg_exchange_push_zero(ctypes.byref(self.push), self.push_n,
                                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "eg_exchange_push_zero")

    @staticmethod
    def gaussian_ranges(n: int, n_ranges: int, align: int = 128):
        """[0, n) in at most ``n_ranges`` contiguous Gaussian ranges whose inner boundaries are multiples of ``align``
        (the Gaussians one CTA of eg_splat_bwd owns; also keeps every slice 16-byte aligned)."""
        n_ranges = max(1, int(n_ranges))
        per = -(-n // n_ranges)
        per = -(-per // align) * align
        return [(b, min(n, b + per)) for b in range(0, n, per)]

    def allreduce_range_(self, n: int, g0: int, g1: int, grid: int = 0) -> None:
        """In-place sum over the ranks of the gradients of the Gaussians [g0, g1) only: the four slices of the flat buffer
        (means | scales | quats | opacities, layout.grad_layout) as ONE launch (eg_allreduce_symm_segs).  ``g0`` must be a
        multiple of 4; a range that ends at ``n`` is extended over the layout's zero padding."""
        import ctypes
        from . import _lib
        from .layout import padded
        assert g0 % 4 == 0 and (g1 == n or g1 % 4 == 0), "range boundaries must keep the slices 16-byte aligned"
        offs = grad_layout(n)
        g1p = padded(n) if g1 == n else g1
        seg_off = (ctypes.c_int64 * 4)(*[offs[k] + w * g0 for k, w in enumerate((3, 3, 4, 1))])
        seg_cnt = (ctypes.c_int64 * 4)(*[w * (g1p - g0) for w in (3, 3, 4, 1)])
        _lib.check(self.lib.eg_allreduce_symm_segs(self._bufs, ctypes.c_void_p(self.multicast_ptr or None), self._flagp, 4,
                                                   seg_off, seg_cnt, self.rank, self.world, int(grid) if grid > 0 else self.grid,
                                                   ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "eg_allreduce_symm_segs")

    def allreduce_(self, count: Optional[int] = None) -> None:
        """In-place sum of the first ``count`` floats (default: all) over the ranks, enqueued on the current stream."""
        import ctypes
        from . import _lib
        n = self.numel if count is None else int(count)
        _lib.check(self.lib.eg_allreduce_symm(self._bufs, ctypes.c_void_p(self.multicast_ptr or None), self._flagp, n,
                                              self.rank, self.world, self.grid,
                                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "eg_allreduce_symm")


class NativeComm:
    """Communicator of libedgegs.so (eg_comm_*, NCCL underneath) over the ranks of ``group``: what
    ``exchange="native-nccl"`` (A/B baseline) all-reduces through.  torch.distributed only carries the 128-byte
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
            # idle rank of a ragged last step: zero gradients, but the same bookkeeping as the working ranks -- the
            # abs-grad normaliser and the step counter feed the densification thresholds (edge_gs.py:544-576), which
            # every rank must evaluate identically
            ws = model._ws
            ws.grads.zero_()
            loss = torch.zeros((), device=ws.grads.device)
            model.absgrads_normalize_factor += 1
            model.step += 1
        else:
            loss = model.raster_step(view, self.gts[view], loss_weight=loss_weight)
        ws = model._ws
        inc = model.absgrads - before
        allreduce_gradients(ws.grads, inc, self.group)
        model.absgrads.copy_(before + inc)
        if self.world > 1:
            dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=self.group)
        return loss


def broadcast_parameters(model, group=None, src: int = 0) -> None:
    """Make every rank's replica identical to rank ``src``'s after a step that draws random numbers per rank
    (the jitter of duplicated Gaussians, edge_gs.py:466): parameters and abs-grad statistic.  Call it after
    densification in a view-sharded run (the masks themselves are computed from all-reduced statistics and are
    already identical)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    src_g = dist.get_global_rank(group, src) if group is not None else src
    for p in model.gauss_params.values():
        dist.broadcast(p.data, src=src_g, group=group)
    dist.broadcast(model.absgrads, src=src_g, group=group)
