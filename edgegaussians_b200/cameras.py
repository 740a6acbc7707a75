"""Host-side mirror of the reference camera interface (projection inputs of the splat path).

Mirrors /root/reference/edgegaussians/cameras/cameras.py:
  Camera        (COLMAP qvec / tvec)            cameras.py:64-101
  OpenCVCamera  (K, R, t)                       cameras.py:103-140
  get_K() -> [1,3,3] fp32, get_viewmat() -> [1,4,4] fp32, .width/.height ints, .to(device)
and the EMAP frame conversion of data/dataparsers.py:107-122 (``from_emap_frame``).
No projection arithmetic lives here (it is inside the kernels); these objects only hold K / viewmat
on the device so a training step never copies them.
"""
from __future__ import annotations

import numpy as np
import torch


def qvec2rotmat(qvec):
    """COLMAP quaternion (w,x,y,z) -> rotation matrix (utils/colmap_read_write_model.py:454-465)."""
    w, x, y, z = [float(v) for v in qvec]
    return np.array([
        [1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * w * z, 2 * z * x + 2 * w * y],
        [2 * x * y + 2 * w * z, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * w * x],
        [2 * z * x - 2 * w * y, 2 * y * z + 2 * w * x, 1 - 2 * x * x - 2 * y * y]])


class BaseCamera:
    width: int
    height: int
    K: torch.Tensor
    viewmat: torch.Tensor
    R: torch.Tensor
    t: torch.Tensor
    device = "cpu"

    def _compose(self):
        bottom = torch.tensor([[0.0, 0.0, 0.0, 1.0]], dtype=self.R.dtype)
        self.viewmat = torch.cat((torch.cat((self.R, self.t.reshape(-1, 1)), dim=1), bottom), dim=0).float()

    def get_K(self) -> torch.Tensor:
        return self.K.reshape(1, 3, 3)

    def get_viewmat(self) -> torch.Tensor:
        return self.viewmat.reshape(1, 4, 4)

    def to(self, device):
        self.device = device
        self.K = self.K.to(device)
        self.viewmat = self.viewmat.to(device)
        return self

    def scale_translation(self, scaling_factor):
        """cameras.py:23-26."""
        self.t = self.t * scaling_factor
        self._compose()


class Camera(BaseCamera):
    """cameras.py:64-101."""

    def __init__(self, height, width, fx, fy, cx, cy, quat, trans, device="cpu", scaling_factor: float = 1.0):
        self.height = int(np.ceil(height * scaling_factor))
        self.width = int(np.ceil(width * scaling_factor))
        self.fx, self.fy = fx * scaling_factor, fy * scaling_factor
        self.cx, self.cy = cx * scaling_factor, cy * scaling_factor
        self.quat = quat
        self.device = device
        self.t = torch.as_tensor(np.asarray(trans)).float()
        self.K = torch.tensor([[self.fx, 0, self.cx], [0, self.fy, self.cy], [0, 0, 1]]).float()
        self.R = torch.from_numpy(qvec2rotmat(quat)).float()
        self._compose()


class OpenCVCamera(BaseCamera):
    """cameras.py:103-140."""

    def __init__(self, height, width, K, R, t):
        self.height, self.width = int(height), int(width)
        K = torch.as_tensor(np.asarray(K)).float() if not isinstance(K, torch.Tensor) else K.float()
        self.K = K[:3, :3].contiguous()
        self.fx, self.fy = float(K[0, 0]), float(K[1, 1])
        self.cx, self.cy = float(K[0, 2]), float(K[1, 2])
        self.R = torch.as_tensor(np.asarray(R)).float() if not isinstance(R, torch.Tensor) else R.float()
        self.t = torch.as_tensor(np.asarray(t)).float() if not isinstance(t, torch.Tensor) else t.float()
        self._compose()

    @classmethod
    def from_emap_frame(cls, height, width, cam_to_world, intrinsics):
        """EMAP meta_data.json frame -> camera (data/dataparsers.py:107-122)."""
        c2w = np.asarray(cam_to_world, dtype=np.float64)
        R_w2c = c2w[:3, :3].T
        t_w2c = -R_w2c @ c2w[:3, 3].reshape(-1, 1)
        return cls(height=height, width=width, K=np.asarray(intrinsics, dtype=np.float64), R=R_w2c, t=t_w2c)

    @classmethod
    def from_matrices(cls, height, width, K, viewmat):
        vm = np.asarray(viewmat, dtype=np.float32)
        return cls(height=height, width=width, K=np.asarray(K, dtype=np.float32), R=vm[:3, :3], t=vm[:3, 3])
