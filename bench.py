#!/usr/bin/env python
"""bench.py -- headline benchmark of the fused forward+backward edge-Gaussian raster iteration.

Metric (BASELINE.json): train iters/sec (fwd+bwd raster) @ 500k Gaussians x 1600x1200, with the
achieved fraction of the B200 HBM roofline.  One "step" = one view: activations + projection +
tile binning + per-tile sort + compositing + "whole" L1 edge-map loss + both backward kernels +
abs-grad accumulation; the optimizer step and KNN are excluded.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line on rank 0 (contract in the task statement).  Synthetic data (synth-v1).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "train iters/sec (fwd+bwd raster) @500k Gaussians x 1600x1200"
UNIT = "iters/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=500_000, help="Gaussians")
    ap.add_argument("--width", type=int, default=1600)
    ap.add_argument("--height", type=int, default=1200)
    ap.add_argument("--regime", default="init", choices=["init", "trained"])
    ap.add_argument("--views", type=int, default=8, help="distinct synthetic views cycled per rank")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed iterations")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lazy-sort", action="store_true", help="always sort every tile (gsplat order) in the tile pipelines")
    ap.add_argument("--pipeline", default="auto", choices=["auto", "splat", "tiles+splat", "tiles"],
                    help="fused-step pipeline (edge_gs.enqueue_raster_step); auto = what training would run")
    ap.add_argument("--allreduce-chunks", type=int, default=1,
                    help="N > 1: Gaussian ranges of the backward whose all-reduce overlaps the next range")
    ap.add_argument("--native-allreduce", action="store_true",
                    help="N > 1: all-reduce through the library's own communicator (eg_comm_allreduce) instead of torch.distributed")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0)
    return ap.parse_args()


def algorithmic_bytes(N, I, P):
    """A = 228 N + 92 I + 20 P, split per stage."""
    stages = {
        "project_fwd": 76 * N,
        "bin": 12 * I,
        "raster_fwd": (24 + 28) * I + 8 * P,   # sort (one read + one write of 12 B) + compositing
        "raster_bwd": 28 * I + 12 * P + 32 * N,
        "project_bwd": 120 * N,
    }
    total = sum(stages.values())
    # fused kernels carry the algorithmic bytes of the stages they replace
    stages["splat_bwd"] = stages["raster_bwd"] + stages["project_bwd"]
    stages["splat_fwd"] = stages["bin"] + stages["raster_fwd"]
    stages["memset"] = 0
    return stages, total


def workload_name(args):
    return f"{args.n} Gaussians x {args.width}x{args.height}, 1 view/iter/GPU, regime={args.regime}"


def ncu_traffic_bytes(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full
    summary (profiles/r1_splat_ncu_full_summary.txt, written by scripts/summarise_profiles.py); None if absent."""
    path = os.path.join(ROOT, "profiles", "r1_splat_ncu_full_summary.txt")
    if not os.path.exists(path):
        return None
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    cur, rd, wr, best = None, None, None, None
    for line in open(path):
        parts = line.split()
        if not parts:
            continue
        if parts[0] == "Kernel" and len(parts) > 2:
            cur, rd, wr = line, None, None
        elif parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and cur and kernel in cur:
            try:
                val = float(parts[1].replace(",", "")) * mult.get(parts[2], 1.0)
            except (ValueError, IndexError):
                continue
            if parts[0].startswith("dram__bytes_read"):
                rd = val
            else:
                wr = val
            if rd is not None and wr is not None:
                best = rd + wr
    return best


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
def cpu_step_time(args, n_sample, reps, threads=None):
    """Time the CPU oracle's full iteration (projection .. loss .. backward) on synth-v1."""
    from edgegaussians_b200 import synth
    from oracle import oracle
    if threads:
        oracle.set_num_threads(threads)
    m, q, s, o = synth.make_gaussians(n_sample, args.regime, 0)
    vms, Ks = synth.make_cameras(max(args.views, 2), args.width, args.height)
    gt = synth.make_edge_map(args.width, args.height, 0)
    times = []
    for r in range(reps):
        t0 = time.perf_counter()
        oracle.edge_step(m, q, s, o, vms[r % len(vms)], Ks[r % len(vms)], args.width, args.height, gt)
        times.append(time.perf_counter() - t0)
    return times, oracle.num_threads()


def run_reference(args):
    """--impl reference: the reference path's CPU implementation (the oracle port: the reference's own
    splat lives in CUDA-only gsplat, absent here) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    oracle.build()
    cores = oracle.num_threads()
    budget = 150.0
    t_probe, _ = cpu_step_time(args, min(args.n, 50_000), 1)
    est_full = t_probe[0] * args.n / min(args.n, 50_000)
    total_steps = args.steps + args.warmup
    frac = min(1.0, budget / max(est_full * total_steps, 1e-9))
    n_sample = max(1000, int(args.n * frac))
    times, cores = cpu_step_time(args, n_sample, total_steps)
    timed = times[args.warmup:]
    t_step = sum(timed) / len(timed)
    scale = args.n / n_sample
    value = 1.0 / (t_step * scale)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_step * scale, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (synth-v1)",
        "config": {"workload": workload_name(args)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n_sample} of {args.n} Gaussians per step at full resolution, time scaled x{scale:.2f} (linear in N)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from edgegaussians_b200 import synth
    from edgegaussians_b200.cameras import OpenCVCamera
    from edgegaussians_b200.edge_gs import EdgeGaussianSplatting
    from edgegaussians_b200.graph_step import GraphedRasterStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    N, W, H, V = args.n, args.width, args.height, args.views
    P = W * H
    m, q, s, o = synth.make_gaussians(N, args.regime, 0)
    vms, Ks = synth.make_cameras(V * world, W, H)
    my_views = [rank + world * i for i in range(V)]            # view-sharded: rank r renders views r, r+G, ...
    gts_u8 = [synth.make_edge_map_u8(W, H, v) for v in my_views]
    model = EdgeGaussianSplatting(device=dev)
    cams = [OpenCVCamera.from_matrices(H, W, Ks[v], vms[v]).to(dev) for v in my_views]
    model.set_params(m, s, q, o, viewcams=cams)
    model.lazy_sort = False if args.no_lazy_sort else "auto"
    model.pipeline = args.pipeline

    step = GraphedRasterStep(model, W, H, n_slots=V, gt_dtype=torch.uint8, allreduce=world > 1,
                             allreduce_chunks=args.allreduce_chunks, native_allreduce=args.native_allreduce)
    host_vm = [torch.from_numpy(vms[v]).pin_memory() for v in my_views]
    host_K = [torch.from_numpy(Ks[v]).pin_memory() for v in my_views]
    host_gt = [torch.from_numpy(g).pin_memory() for g in gts_u8]
    for i in range(V):
        step.set_view(i, host_vm[i], host_K[i], host_gt[i])
    torch.cuda.synchronize()
    n_isects_max = step.calibrate()
    for i in range(V):
        step.capture(i)
    ws = step.ws
    grads = ws.grads

    flush_buf = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def flush():
        if flush_buf is not None:
            flush_buf.fill_(1)

    def one_step(i):
        step.replay(i % V)   # includes the NCCL all-reduce of the gradient buffer when world > 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing (`value`) ----------------
    for i in range(args.warmup):
        flush(); one_step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush()
        ev0[i].record()
        one_step(i)
        ev1[i].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    per_step_ms = [a.elapsed_time(b) for a, b in zip(ev0, ev1)]
    total_ms = sum(per_step_ms)
    # informational: the same K steps back to back WITHOUT the L2 flush (parameters / records stay L2-resident,
    # as they would between optimizer steps); not the headline
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        one_step(i)
    e1.record()
    barrier()
    hot_ms = e0.elapsed_time(e1)
    # the timed region lasts only tens of milliseconds: keep the same load running (untimed) for about a second so
    # that the nvidia-smi sampler (100 ms period) sees the clocks / throttle reasons this workload runs at
    # (a fixed iteration COUNT agreed by all ranks: every step holds a collective, so a time-based loop would
    # run a different number of all-reduces per rank and dead-lock)
    n_cont = torch.tensor([max(32, min(20000, int(1200.0 * args.steps / max(total_ms, 1e-3))))], device=dev)
    if world > 1:
        dist.broadcast(n_cont, src=0)
    for i in range(int(n_cont[0])):
        flush(); one_step(i)
        if i % 64 == 63:
            torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    n_isects = int(ws.status[0])
    overflow = int(ws.status[1])

    # ---------------- per-kernel breakdown (events between stages, eager launches) ----------------
    names, acc = None, {}
    isect_sum = 0
    reps = min(args.steps, 16)
    for i in range(reps):
        evs = {}

        def cb(name, evs=evs):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            evs[name] = e
        flush()  # also lets the host run ahead of the device so events see no launch gaps
        step._enqueue(i % V, stage_cb=cb)
        torch.cuda.synchronize()
        if names is None:
            names = [k for k in evs if k != "begin"]   # stage order as enqueued (dicts keep insertion order)
            acc = {k: 0.0 for k in names}
        prev = "begin"
        for k in names:
            acc[k] += evs[prev].elapsed_time(evs[k])
            prev = k
        isect_sum += int(ws.status[0])
    kern_ms = {k: acc[k] / reps for k in names}
    I_mean = isect_sum / reps

    # ---------------- end-to-end through the public API with host buffers (`e2e`) ----------------
    loss_host = torch.zeros(1, dtype=torch.float64).pin_memory()
    copy_stream = torch.cuda.Stream()
    h2d = host_gt[0].numel() * host_gt[0].element_size() + 16 * 4 + 9 * 4
    d2h = 8

    def upload(slot, i):
        with torch.cuda.stream(copy_stream):
            step.set_view(slot, host_vm[i % V], host_K[i % V], host_gt[i % V])
            e = torch.cuda.Event()
            e.record(copy_stream)
        return e

    def e2e_loop(n):
        # slot ping-pong: the copy of step i+1's inputs overlaps step i's kernels (all inside the timed region)
        cur = torch.cuda.current_stream()
        done = [None, None]  # per slot: recorded after the last replay that read it
        pend = upload(0, 0)
        for i in range(n):
            slot = i % 2
            cur.wait_event(pend)
            if i + 1 < n:
                nslot = (i + 1) % 2
                if done[nslot] is not None:
                    copy_stream.wait_event(done[nslot])
                pend = upload(nslot, i + 1)
            step.replay(slot)
            loss_host.copy_(ws.loss_sum, non_blocking=True)
            e = torch.cuda.Event()
            e.record(cur)
            done[slot] = e
        torch.cuda.synchronize()

    e2e_loop(max(2, args.warmup))
    barrier()
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    barrier()
    t_e2e = time.perf_counter() - t0
    # restore slot contents for any later use
    for i in range(V):
        step.set_view(i, host_vm[i], host_K[i], host_gt[i])
    torch.cuda.synchronize()

    # ---------------- reductions over ranks ----------------
    vals = torch.tensor([total_ms, t_e2e * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(vals[0]), float(vals[1])

    if rank == 0:
        peak, peak_src = load_peaks()
        stages, A = algorithmic_bytes(N, I_mean, P)
        ms_per_step = total_ms / args.steps
        value = world * args.steps / (total_ms * 1e-3)
        kernels = [k for k in names if k != "memset"]
        dom = max(kernels, key=lambda k: kern_ms[k])
        achieved = stages[dom] / (kern_ms[dom] * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic (synth-v1: uniform Gaussians in [-1,1]^3, Fibonacci-sphere cameras, random-segment edge maps)",
            "config": {"workload": workload_name(args), "views_cycled_per_gpu": V,
                       "n_isects": I_mean, "isect_per_gaussian": I_mean / N, "overflow": overflow,
                       "l2": "flushed between timed iterations (256 MiB fill)" if flush_buf is not None else "not flushed",
                       "pipeline": ws.pipeline + {"splat": " (Gaussian-major forward + backward; tiles near the transmittance stop threshold redone sorted)",
                                                  "tiles+splat": " (tile binning + per-tile sort/compositing, Gaussian-major backward)",
                                                  "tiles": " (tile binning + per-tile sort/compositing, tile-major backward)"}[ws.pipeline],
                       "stopped_tiles": int(ws.status[5]),
                       "tile_sort": ("n/a" if ws.pipeline == "splat" else "lazy (only tiles near the transmittance stop threshold)" if model._use_lazy() else "every tile"),
                       "execution": f"CUDA graph replay per iteration (1 memset + {ws.n_kernels} kernels; stages: {', '.join(kernels)})" + ((", + NCCL all-reduce of the 11N fp32 gradient buffer " + (f"in {step.allreduce_chunks} Gaussian ranges on a side stream, overlapped with the backward of the next range" if step.chunked else ("issued on the compute stream after the replay (libedgegs communicator)" if step.native_comm is not None else "issued after the replay (torch.distributed)"))) if world > 1 else ""),
                       "parallelism": f"view-sharded dp{world}" if world > 1 else "single GPU"},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic_bytes(dom + "_kernel"), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": stages[dom], "kernel_ms": kern_ms[dom]},
            "roofline_step": {"algorithmic_bytes": A, "achieved": A / (ms_per_step * 1e-3) / 1e9,
                              "frac": A / (ms_per_step * 1e-3) / 1e9 / peak, "formula": "228 N + 92 I + 20 P"},
            "kernel_ms": kern_ms,
            "e2e": {"value": world * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "api": "GraphedRasterStep.set_view(pinned host) + replay + loss readback"},
            "value_no_l2_flush": world * args.steps / (hot_ms * 1e-3),
            "gpu_launches": ws.n_kernels * args.steps,
            "clocks": clocks,
            "wall_s_timed_region": t_wall,
        }
        if not args.no_cpu_baseline and world == 1:
            try:
                t_probe, cores = cpu_step_time(args, min(N, 50_000), 1)
                est = t_probe[0] * N / min(N, 50_000)
                n_s = max(1000, int(N * min(1.0, args.cpu_budget_s / max(2 * est, 1e-9))))
                ts, cores = cpu_step_time(args, n_s, 2)
                t_cpu = min(ts) * N / n_s
                line["cpu_baseline"] = {"value": 1.0 / t_cpu, "unit": UNIT, "cores": cores, "kind": "port",
                                        "sample": f"{n_s} of {N} Gaussians at full resolution, best of 2, time scaled linearly in N"}
            except Exception as e:  # the oracle is only a reported baseline
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
