"""Layout of the flat fp32 gradient buffer of the fused iteration (the buffer that is all-reduced in the
view-sharded multi-GPU step):

    means [N,3] | scales [N,3] | quats [N,4] | opacities [N]

Every segment starts on a 16-byte boundary: the segment offsets use N rounded up to a multiple of 4, so the
128-bit quaternion-gradient stores of eg_splat_bwd / eg_project_bwd are aligned for ANY N (odd N after a cull or a
duplication included).  The pad floats between the segments are never written by the kernels and stay zero.
The same arithmetic is exported from C as ``eg_grad_layout`` (include/edgegs.h).
"""
from __future__ import annotations

from typing import Tuple


def padded(n: int) -> int:
    return (int(n) + 3) // 4 * 4


def grad_layout(n: int) -> Tuple[int, int, int, int, int]:
    """(offset of means, scales, quats, opacities, total length), in floats."""
    p = padded(n)
    return 0, 3 * p, 6 * p, 10 * p, 11 * p


def grad_numel(n: int) -> int:
    return 11 * padded(n)


def split_grads(flat, n: int):
    """(v_means [N,3], v_scales [N,3], v_quats [N,4], v_opacities [N]) views of the flat buffer."""
    om, os_, oq, oo, _ = grad_layout(n)
    # reshape of a contiguous slice is a view for torch tensors and numpy arrays alike
    return (flat[om:om + 3 * n].reshape(n, 3), flat[os_:os_ + 3 * n].reshape(n, 3), flat[oq:oq + 4 * n].reshape(n, 4),
            flat[oo:oo + n])
