"""edgegaussians_b200 -- B200-native differentiable edge-Gaussian splat path.

Drop-in for the one third-party call on the reference's hot path:
    from gsplat import rasterization          (/root/reference/edgegaussians/models/edge_gs.py:8)
->  from edgegaussians_b200 import rasterization
See DESIGN.md / INTEGRATION.md.  CUDA only; there is no CPU fallback.
"""
from .rasterization import rasterization  # noqa: F401

__all__ = ["rasterization"]
