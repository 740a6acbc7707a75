"""PLY hand-off format (reference utils/io_utils.py:4-39): layout, header and round trip (host-side, no GPU)."""
import numpy as np

from edgegaussians_b200 import io_utils


def test_ply_layout_and_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    n = 257
    means, scales = rng.normal(size=(n, 3)).astype(np.float32), rng.uniform(1e-3, 1, (n, 3)).astype(np.float32)
    quats, opac = rng.normal(size=(n, 4)).astype(np.float32), rng.uniform(0, 1, (n, 1)).astype(np.float32)
    path = tmp_path / "gaussians_all.ply"
    io_utils.write_gaussian_params_as_ply(means, scales, quats, opac, path)
    raw = path.read_bytes()
    header = io_utils.ply_header(n)
    assert raw.startswith(b"ply\nformat binary_little_endian 1.0\nelement vertex 257\nproperty float x\n")
    assert header.endswith(b"property float opacity\nend_header\n") and raw[:len(header)] == header
    assert len(raw) == len(header) + n * 11 * 4                      # eleven little-endian float32 per Gaussian
    rec = np.frombuffer(raw, "<f4", offset=len(header)).reshape(n, 11)
    np.testing.assert_array_equal(rec, np.concatenate([means, scales, quats, opac], 1))   # x y z s1-3 q1-4 opacity
    pos, s, q, o = io_utils.read_gaussian_params_from_ply(path)
    for got, exp in ((pos, means), (s, scales), (q, quats), (o, opac)):
        np.testing.assert_array_equal(got, exp)
    assert o.shape == (n, 1)


def test_export_as_ply_writes_activated_values(tmp_path):
    import torch
    from edgegaussians_b200.edge_gs import EdgeGaussianSplatting
    m = EdgeGaussianSplatting(device="cpu")
    n = 5
    m.set_params(np.arange(3 * n, dtype=np.float32).reshape(n, 3), np.log(np.full((n, 3), 0.004, np.float32)),
                 np.tile(np.array([1, 0, 0, 0], np.float32), (n, 1)), np.zeros((n, 1), np.float32))
    m.export_as_ply(tmp_path / "g.ply")
    pos, s, q, o = io_utils.read_gaussian_params_from_ply(tmp_path / "g.ply")
    np.testing.assert_allclose(s, 0.004, rtol=1e-6)                   # exp(log-scales)
    np.testing.assert_allclose(o, 0.5)                                # sigmoid(0)
    np.testing.assert_array_equal(pos, m.means.detach().numpy())
