"""In-tree build of libedgegs.so (hand-written sm_100a CUDA behind include/edgegs.h).

    python -m edgegaussians_b200.build [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU; the .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_C")
LIB = os.path.join(OUT_DIR, "libedgegs.so")
SOURCES = ["eg_api.cu", "eg_project_fwd.cu", "eg_bin.cu", "eg_raster_fwd.cu", "eg_raster_bwd.cu", "eg_splat_bwd.cu", "eg_splat_fwd.cu", "eg_comm.cu", "eg_allreduce.cu",
           "eg_project_bwd.cu", "eg_reg.cu", "eg_knn.cu", "eg_adam.cu", "eg_visibility.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]
NVCC_FLAGS += os.environ.get("EG_NVCC_EXTRA", "").split()  # tuning experiments, e.g. -DEG_ROWS_PER_ITEM=4


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "edgegs.h"))
    objs, log = [], []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        if not os.path.exists(sp):
            continue
        obj = os.path.join(OUT_DIR, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [sp] + headers):
            cmd = [_nvcc()] + NVCC_FLAGS + ["-c", sp, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            log.append(f"$ {' '.join(cmd)}\n{r.stdout}{r.stderr}")
            if r.returncode != 0:
                sys.stderr.write(log[-1])
                raise RuntimeError(f"nvcc failed on {src}")
            if verbose:
                print(log[-1])
    if force or _stale(LIB, objs):
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    if log:
        with open(os.path.join(OUT_DIR, "build.log"), "w") as f:
            f.write("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
