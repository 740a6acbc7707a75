#!/usr/bin/env python
"""BASELINE.json configs 3-5 through bench.py (run on the GPU box):  python scripts/sweep.py --gpus G [--quick]
   config 3: 200k Gaussians x 1600x1200;  config 4: 500k x 1200x680;  config 5: 10k -> 2M x 1920x1080.
   One bench.py process (or torchrun group) per row; rows go to gpurun_out/r2_sweep_g<G>.jsonl and a table to stdout."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--quick", action="store_true", help="fewer sweep sizes (multi-GPU boxes are charged per GPU)")
ap.add_argument("--regime", default="both")
a = ap.parse_args()
rows = [("cfg3 DTU-like", 200_000, 1600, 1200), ("cfg4 Replica-like", 500_000, 1200, 680)]
sizes = [100_000, 1_000_000, 2_000_000] if a.quick else [10_000, 100_000, 300_000, 1_000_000, 2_000_000]
rows += [(f"cfg5 sweep", n, 1920, 1080) for n in sizes]
out_path = os.path.join(ROOT, "gpurun_out", f"r2_sweep_g{a.gpus}.jsonl")
os.makedirs(os.path.dirname(out_path), exist_ok=True)
open(out_path, "w").close()
print(f"| config | Gaussians | image | GPUs | regime | pipeline | ms/step | it/s | step roofline | dominant kernel (frac) |")
print("|---|---|---|---|---|---|---|---|---|---|")
for name, n, w, h in rows:
    cmd = [sys.executable]
    if a.gpus > 1:
        cmd += ["-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}", "--master-addr", "127.0.0.1",
                "--master-port", "29571"]
    cmd += [os.path.join(ROOT, "bench.py"), "--gpus", str(a.gpus), "--gaussians", str(n), "--width", str(w), "--height", str(h),
            "--steps", "30", "--warmup", "5", "--no-cpu-baseline", "--no-aux", "--regime", a.regime]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
        d = json.loads(line)
    except Exception as e:
        print(f"| {name} | {n} | {w}x{h} | {a.gpus} | FAILED: {e!r} |")
        continue
    with open(out_path, "a") as f:
        f.write(json.dumps({"config": name, "n": n, "width": w, "height": h, "line": d}) + "\n")
    for reg, res in d["regimes"].items():
        rf = res["roofline"]
        print(f"| {name} | {n} | {w}x{h} | {d['n_gpus']} | {reg} | {res['pipeline']} | {res['ms_per_step']:.4f} | {res['value']:.0f} | "
              f"{res['roofline_step']['frac']:.3f} | {rf['kernel']} ({rf['frac']:.3f}) |", flush=True)
