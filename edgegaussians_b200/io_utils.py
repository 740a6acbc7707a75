"""PLY hand-off format of the reference: the file `fit_edges.py` reads.

Mirrors write_gaussian_params_as_ply / read_gaussian_params_from_ply of
/root/reference/edgegaussians/utils/io_utils.py:4-39: one `vertex` element with eleven little-endian float32
properties x, y, z, scale1-3, quat1-4, opacity holding ACTIVATED values (exp of the log-scales, sigmoid of the
logits; edge_gs.py:635-642).  The reference goes through the third-party `plyfile` package, which is absent from
this image, so the bytes are written directly: PLY 1.0 binary_little_endian with exactly the header `plyfile` emits
for that dtype ("property float <name>" per 'f4' field, "\\n" line ends, no comment lines) and the packed records.
tests/test_io_ply.py holds the expected file byte by byte (known-answer test)."""
from __future__ import annotations

import numpy as np

FIELDS = ("x", "y", "z", "scale1", "scale2", "scale3", "quat1", "quat2", "quat3", "quat4", "opacity")
VERTEX_DTYPE = np.dtype([(f, "<f4") for f in FIELDS])


def ply_header(n: int) -> bytes:
    lines = ["ply", "format binary_little_endian 1.0", f"element vertex {int(n)}"]
    lines += [f"property float {f}" for f in FIELDS]
    lines.append("end_header")
    return ("\n".join(lines) + "\n").encode("ascii")


def write_gaussian_params_as_ply(means, scales, quats, opacities, ply_path) -> None:
    means, scales, quats = (np.asarray(a, np.float32) for a in (means, scales, quats))
    opacities = np.asarray(opacities, np.float32).reshape(means.shape[0], -1)
    vertex = np.zeros(means.shape[0], dtype=VERTEX_DTYPE)
    for j, f in enumerate(("x", "y", "z")):
        vertex[f] = means[:, j]
    for j in range(3):
        vertex[f"scale{j + 1}"] = scales[:, j]
    for j in range(4):
        vertex[f"quat{j + 1}"] = quats[:, j]
    vertex["opacity"] = opacities[:, 0]
    with open(ply_path, "wb") as fh:
        fh.write(ply_header(vertex.shape[0]))
        fh.write(vertex.tobytes())


def read_gaussian_params_from_ply(ply_path):
    """(pos [N,3], scales [N,3], quats [N,4], opacities [N,1]) as the reference's reader returns them."""
    with open(ply_path, "rb") as fh:
        raw = fh.read()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    header = raw[:end].decode("ascii").split("\n")
    if header[0] != "ply" or not header[1].startswith("format binary_little_endian"):
        raise ValueError("expected a binary little-endian PLY file")
    n = next(int(l.split()[2]) for l in header if l.startswith("element vertex"))
    props = [l.split()[2] for l in header if l.startswith("property ")]
    if tuple(props) != FIELDS:
        raise ValueError(f"unexpected vertex properties {props}")
    data = np.frombuffer(raw, dtype=VERTEX_DTYPE, count=n, offset=end)
    col = lambda names: np.stack([data[f] for f in names], -1).astype(np.float32)
    return (col(("x", "y", "z")), col(("scale1", "scale2", "scale3")), col(("quat1", "quat2", "quat3", "quat4")),
            col(("opacity",)))
