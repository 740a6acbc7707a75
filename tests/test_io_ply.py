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


# The exact file `plyfile` writes for this dtype, spelled out byte by byte (known-answer test; the package itself is
# absent from this image, so the expectation is a literal: PlyData([PlyElement.describe(vertex, 'vertex')]).write()
# emits "ply", "format binary_little_endian 1.0", one "element vertex N", one "property float <name>" per 'f4' field
# in dtype order, "end_header", every line terminated by "\n" (no "\r", no comment or obj_info lines), followed by the
# packed little-endian records -- utils/io_utils.py:4-25).
GOLDEN_HEADER = (b"ply\n"
                 b"format binary_little_endian 1.0\n"
                 b"element vertex 2\n"
                 b"property float x\n"
                 b"property float y\n"
                 b"property float z\n"
                 b"property float scale1\n"
                 b"property float scale2\n"
                 b"property float scale3\n"
                 b"property float quat1\n"
                 b"property float quat2\n"
                 b"property float quat3\n"
                 b"property float quat4\n"
                 b"property float opacity\n"
                 b"end_header\n")
GOLDEN_PAYLOAD = bytes.fromhex(
    "0000803f" "000000c0" "00004040"      # x y z            = 1, -2, 3
    "0000003f" "0000803e" "0000003e"      # scale1-3         = 0.5, 0.25, 0.125
    "0000803f" "00000000" "00000000" "00000000"   # quat1-4  = 1, 0, 0, 0
    "0000403f"                            # opacity          = 0.75
    "000020c1" "0000a041" "0000f0c1"      # x y z            = -10, 20, -30
    "00000040" "00008040" "00000041"      # scale1-3         = 2, 4, 8
    "00000000" "0000803f" "00000000" "00000000"   # quat1-4  = 0, 1, 0, 0
    "0000803e")                           # opacity          = 0.25


def test_ply_bytes_known_answer(tmp_path):
    means = np.array([[1, -2, 3], [-10, 20, -30]], np.float32)
    scales = np.array([[0.5, 0.25, 0.125], [2, 4, 8]], np.float32)
    quats = np.array([[1, 0, 0, 0], [0, 1, 0, 0]], np.float32)
    opac = np.array([[0.75], [0.25]], np.float32)
    path = tmp_path / "two.ply"
    io_utils.write_gaussian_params_as_ply(means, scales, quats, opac, path)
    raw = path.read_bytes()
    assert raw == GOLDEN_HEADER + GOLDEN_PAYLOAD
    assert len(GOLDEN_PAYLOAD) == 2 * 11 * 4
    # float64 / [N] inputs are converted like the reference's structured-array assignment does
    io_utils.write_gaussian_params_as_ply(means.astype(np.float64), scales, quats, opac, path)
    assert path.read_bytes() == GOLDEN_HEADER + GOLDEN_PAYLOAD
    # and the reader returns the reference's shapes
    pos, s, q, o = io_utils.read_gaussian_params_from_ply(path)
    assert pos.shape == (2, 3) and s.shape == (2, 3) and q.shape == (2, 4) and o.shape == (2, 1)
