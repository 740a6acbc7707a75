#!/usr/bin/env python
"""bench.py -- headline benchmark of the fused forward+backward edge-Gaussian raster iteration.

Metric (BASELINE.json): train iters/sec (fwd+bwd raster) @ 500k Gaussians x 1600x1200, with the
achieved fraction of the B200 HBM roofline.  One "step" = one view: activations + projection +
(tile binning + per-tile sort) + compositing + "whole" L1 edge-map loss + backward + abs-grad
accumulation (+ the gradient exchange when N > 1); the optimizer step, the
regularisers and KNN are excluded from the step and reported separately (`aux_ms`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--regime both|init|trained]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line on rank 0 (contract in the task statement).  Synthetic data (synth-v1).  The headline
`value` is the init regime (reference initialisation: isotropic 0.004, opacity 0.08); the default run also
measures the trained regime (5:1 anisotropy, opacity U(0.05, 0.9)) and reports both under `regimes`, each
with its own roofline and its own full-size parity check against the CPU oracle.
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
    ap.add_argument("--n", "--gaussians", dest="n", type=int, default=500_000,
                    help="Gaussians (use --gaussians under torchrun: its own parser trips over the prefix --n)")
    ap.add_argument("--width", type=int, default=1600)
    ap.add_argument("--height", type=int, default=1200)
    ap.add_argument("--regime", default="both", choices=["both", "init", "trained"])
    ap.add_argument("--views", type=int, default=8, help="distinct synthetic views cycled per rank")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed iterations")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle (no parity block, no cpu_baseline)")
    ap.add_argument("--no-aux", action="store_true", help="skip the optimizer / regulariser / KNN timings")
    ap.add_argument("--no-lazy-sort", action="store_true", help="always sort every tile (gsplat order) in the tile pipelines")
    ap.add_argument("--no-cull", action="store_true", help="emit keys to every tile of gsplat's rectangle (no footprint culling)")
    ap.add_argument("--no-front-sort", action="store_true", help="sort whole tile lists (no depth-sliced early stop)")
    ap.add_argument("--no-morton", action="store_true",
                    help="keep the Gaussians in creation order (default: Morton-ordered once before the run, as the trainer does after "
                         "populate / densify: EdgeGaussianSplatting.sort_gaussians_morton)")
    ap.add_argument("--pipeline", default="auto", choices=["auto", "splat", "tiles+splat", "tiles"],
                    help="fused-step pipeline (edge_gs.enqueue_raster_step); auto = what training would run")
    ap.add_argument("--exchange-ranges", type=int, default=1,
                    help="N > 1: Gaussian ranges of the backward whose exchange overlaps the next range (library exchange only)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "push", "push-p2p", "symm", "symm-p2p", "nccl", "native-nccl"],
                    help="N > 1: gradient exchange -- push form fused into the backward's stores (default), the library's "
                         "all-reduce kernel over symmetric memory (symm*), or NCCL (A/B)")
    return ap.parse_args()


def algorithmic_bytes(N, I, P):
    """A = 228 N + 92 I + 20 P, split per stage (I = gsplat's n_isects)."""
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


def workload_name(args, regime="init"):
    return f"{args.n} Gaussians x {args.width}x{args.height}, 1 view/iter/GPU, regime={regime}"


def ncu_traffic_bytes(kernel, regime="init"):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the newest committed
    `ncu --set full` summary of that regime under profiles/ (written by scripts/summarise_profiles.py from a capture
    of the same workload, scripts/profile_step.py); None if no capture of that kernel is committed.  (ncu cannot run
    inside the timed bench: a number measured under a profiler is never a bench value.)"""
    pdir = os.path.join(ROOT, "profiles")
    want = "trained_ncu_full_summary.txt" if regime == "trained" else "ncu_full_summary.txt"
    cands = sorted((f for f in os.listdir(pdir) if f.endswith(want) and (regime == "trained" or "trained" not in f)),
                   reverse=True) if os.path.isdir(pdir) else []
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for fn in cands:
        cur, rd, wr, best = None, None, None, None
        for line in open(os.path.join(pdir, fn)):
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
        if best is not None:
            return best, fn
    return None, None


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
# CPU legs (the oracle is test infrastructure: it is only ever the checker or the reported baseline)
# ----------------------------------------------------------------------------------------------
def cpu_all_cores():
    """Use every host core, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    from oracle import oracle
    oracle.build()
    oracle.set_num_threads(os.cpu_count() or 1)
    return oracle.num_threads()


def cpu_full_step(args, regime, view):
    """One full-size iteration of the CPU oracle on synth-v1 (same inputs as the GPU rank-0 slot `view`).
    Returns (result dict, seconds)."""
    from edgegaussians_b200 import synth
    from oracle import oracle
    m, q, s, o = synth.make_gaussians(args.n, regime, 0)
    vms, Ks = synth.make_cameras(args.views * max(args.gpus, 1), args.width, args.height)
    gt = synth.make_edge_map_u8(args.width, args.height, view).astype(np.float32) / np.float32(255.0)
    t0 = time.perf_counter()
    ref = oracle.edge_step(m, q, s, o, vms[view], Ks[view], args.width, args.height, gt)
    return ref, time.perf_counter() - t0


def run_reference(args):
    """--impl reference: the reference path's CPU implementation (the oracle port: the reference's own splat lives in
    CUDA-only gsplat, absent here) on ALL host cores, full size, no sub-sampling.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = cpu_all_cores()
    regime = "init" if args.regime == "both" else args.regime
    n_views = args.views * max(args.gpus, 1)
    times = []
    for i in range(args.warmup + args.steps):
        _, dt = cpu_full_step(args, regime, i % n_views)
        times.append(dt)
        if sum(times) > 240.0 and i + 1 >= args.warmup + 3:   # bounded wall time; says so in `sample`
            break
    timed = times[args.warmup:]
    t_step = sum(timed) / len(timed)
    value = 1.0 / t_step
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(timed),
        "warmup": args.warmup, "ms_per_step": 1e3 * t_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (synth-v1)",
        "config": {"workload": workload_name(args, regime)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{len(timed)} full-size iterations ({args.n} Gaussians, {args.width}x{args.height}), "
                                   f"no sub-sampling, {cores} threads (os.cpu_count() = {os.cpu_count()})"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def parity_block(ref, loss, grads_flat, absgrad, n_isects, N):
    """Full-size parity of one GPU iteration against the CPU oracle on identical inputs: tolerances of
    tests/test_gpu_parity.py (gradients |d| <= 1e-3 |ref| + 2e-5 max|ref|, at most max(2, 5e-4 n) violations)."""
    from edgegaussians_b200.layout import split_grads
    gm, gs, gq, go = split_grads(grads_flat, N)
    out = {"loss_gpu": float(loss), "loss_oracle": float(ref["loss"]), "loss_abs_err": abs(float(loss) - float(ref["loss"])),
           "n_isects_gpu": int(n_isects), "n_isects_oracle": int(ref["state"]["n_isects"]),
           "n_isects_equal": int(n_isects) == int(ref["state"]["n_isects"]), "grad_violations": {}, "grad_max_rel_err": {}}
    ok = out["n_isects_equal"] and out["loss_abs_err"] <= 1e-5
    for key, got in (("v_means", gm), ("v_log_scales", gs), ("v_quats", gq), ("v_logit_opacities", go), ("absgrad_norm", absgrad)):
        exp = ref[key].reshape(-1)
        got = got.reshape(-1)
        scale = float(np.abs(exp).max())
        err = np.abs(got - exp)
        # isotropic Gaussians (init regime): the true quaternion gradient is a cancellation to ~0 of O(|v_means|) terms
        floor = 1e-6 * float(np.abs(ref["v_means"]).max()) if key == "v_quats" else 0.0
        bad = int((err > 1e-3 * np.abs(exp) + 2e-5 * scale + floor).sum())
        out["grad_violations"][key] = bad
        out["grad_max_rel_err"][key] = float(err.max() / (scale + floor + 1e-30))
        ok = ok and bad <= max(2, int(5e-4 * err.size))
    out["ok"] = bool(ok)
    return out


# ----------------------------------------------------------------------------------------------
def bench_regime(args, regime, ctx):
    """Everything measured for one regime on this rank; rank 0 returns the summary dict."""
    import torch
    import torch.distributed as dist
    from edgegaussians_b200 import synth
    from edgegaussians_b200.cameras import OpenCVCamera
    from edgegaussians_b200.edge_gs import EdgeGaussianSplatting
    from edgegaussians_b200.graph_step import GraphedRasterStep

    world, rank, dev = ctx["world"], ctx["rank"], ctx["dev"]
    N, W, H, V = args.n, args.width, args.height, args.views
    P = W * H
    m, q, s, o = synth.make_gaussians(N, regime, 0)
    vms, Ks = synth.make_cameras(V * world, W, H)
    # view-sharded: at step i rank r renders view i * G + r (slot i of that rank)
    my_views = [i * world + rank for i in range(V)]
    gts_u8 = [synth.make_edge_map_u8(W, H, v) for v in my_views]
    model = EdgeGaussianSplatting(device=dev)
    cams = [OpenCVCamera.from_matrices(H, W, Ks[v], vms[v]).to(dev) for v in my_views]
    model.set_params(m, s, q, o, viewcams=cams)
    model.lazy_sort = False if args.no_lazy_sort else "auto"
    model.pipeline = args.pipeline
    model.cull_tiles = not args.no_cull
    model.front_sort = not args.no_front_sort
    perm = None
    if not args.no_morton:
        perm = model.sort_gaussians_morton()   # once, like after a densify; parity below maps back through `perm`

    step = GraphedRasterStep(model, W, H, n_slots=V, gt_dtype=torch.uint8, allreduce=world > 1, exchange=args.exchange,
                             exchange_ranges=args.exchange_ranges)
    host_vm = [torch.from_numpy(vms[v]).pin_memory() for v in my_views]
    host_K = [torch.from_numpy(Ks[v]).pin_memory() for v in my_views]
    host_gt = [torch.from_numpy(g).pin_memory() for g in gts_u8]
    for i in range(V):
        step.set_view(i, host_vm[i], host_K[i], host_gt[i])
    torch.cuda.synchronize()
    step.calibrate()
    for i in range(V):
        step.capture(i)
    ws = step.ws

    flush_buf = ctx["flush_buf"]

    def flush():
        if flush_buf is not None:
            flush_buf.fill_(1)

    def one_step(i):
        step.replay(i % V)   # includes the gradient exchange when world > 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- parity at full size: this rank-0 iteration against the CPU oracle ----------------
    parity, cpu_s, cpu_cores = None, None, None
    if rank == 0 and not args.no_cpu_baseline:
        cpu_cores = cpu_all_cores()
        ws_e = step._enqueue(0, accumulate_absgrad=False)   # eager, no exchange: one view on one GPU
        torch.cuda.synchronize()
        hs = ws_e.status.cpu()
        grads_host = ws_e.grads.cpu().numpy().copy()
        loss_gpu = float(step.loss())
        # abs-grad norm of this view alone
        absg = torch.zeros(N, device=dev)
        saved = model.absgrads
        model.absgrads = absg
        step._enqueue(0, accumulate_absgrad=True)
        torch.cuda.synchronize()
        model.absgrads = saved
        ref, cpu_s = cpu_full_step(args, regime, my_views[0])
        absg_h = absg.cpu().numpy()
        if perm is not None:   # the model holds the Gaussians in Morton order: undo it for the comparison
            from edgegaussians_b200.layout import grad_layout
            inv = np.empty(N, np.int64)
            inv[perm.cpu().numpy()] = np.arange(N)
            offs = grad_layout(N)
            g2 = grads_host.copy()
            for off, w in zip(offs[:4], (3, 3, 4, 1)):
                g2[off:off + w * N] = grads_host[off:off + w * N].reshape(N, w)[inv].reshape(-1)
            grads_host, absg_h = g2, absg_h[inv]
        parity = parity_block(ref, loss_gpu, grads_host, absg_h, int(hs[0]), N)
        parity["stopped_tiles"] = int(hs[5])
    barrier()

    # ---------------- N > 1: the exchanged buffer equals the sum of the ranks' single-GPU gradients ----------------
    sum_check = None
    if world > 1:
        one_step(0)
        torch.cuda.synchronize()
        exchanged = ws.grads.clone()
        barrier()
        if rank == 0:
            acc = torch.zeros_like(exchanged, dtype=torch.float64)
            tmp_gt = [synth.make_edge_map_u8(W, H, r) for r in range(world)]   # slot 0 of rank r is view r
            for r in range(world):
                step.set_view(0, torch.from_numpy(vms[r]), torch.from_numpy(Ks[r]), torch.from_numpy(tmp_gt[r]), non_blocking=False)
                w_r = step._enqueue(0, accumulate_absgrad=False)
                torch.cuda.synchronize()
                acc += w_r.grads.double()
            step.set_view(0, host_vm[0], host_K[0], host_gt[0], non_blocking=False)
            scale = float(acc.abs().max())
            err = (exchanged.double() - acc).abs()
            bad = int((err > 1e-5 * acc.abs() + 1e-6 * scale).sum())
            sum_check = {"ranks": world, "max_abs_err": float(err.max()), "max_abs": scale, "violations": bad,
                         "ok": bad == 0, "tolerance": "|d| <= 1e-5 |sum| + 1e-6 max|sum|"}
        barrier()

    # ---------------- device-resident timing (`value`) ----------------
    for i in range(args.warmup):
        flush(); one_step(i)
    barrier()
    sampler = ClockSampler(ctx["local"])
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
    # this rank's mean step time per view slot: the spread is what a view-sharded step pays as rank skew (max over ranks)
    slot_ms = [float(np.mean(per_step_ms[v::V])) for v in range(min(V, args.steps))]
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
    # (a fixed iteration COUNT agreed by all ranks: every step holds a collective)
    n_cont = torch.tensor([max(32, min(20000, int(1200.0 * args.steps / max(total_ms, 1e-3))))], device=dev)
    if world > 1:
        dist.broadcast(n_cont, src=0)
    for i in range(int(n_cont[0])):
        flush(); one_step(i)
        if i % 64 == 63:
            torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    overflow = int(ws.status[1])

    # ---------------- per-kernel breakdown (events between stages, eager launches) ----------------
    names, acc_t = None, {}
    isect_sum = keys_sum = 0
    reps = min(args.steps, 16)
    for i in range(reps):
        evs = {}

        def cb(name, evs=evs):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            evs[name] = e
        flush()  # also lets the host run ahead of the device so events see no launch gaps
        step._enqueue(i % V, stage_cb=cb, accumulate_absgrad=False)
        ex0 = ex1 = None
        if step.exchange is not None:
            # exchange kernel alone (all ranks enter together: its own in-kernel barriers line the ranks up)
            ex0, ex1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ex0.record()
            if step.exchange.push is not None:
                step.exchange.reduce_bcast_()   # the part of the push form that is not hidden behind the backward
            else:
                step.exchange.allreduce_()
            ex1.record()
        torch.cuda.synchronize()
        if names is None:
            names = [k for k in evs if k != "begin"]   # stage order as enqueued (dicts keep insertion order)
            acc_t = {k: 0.0 for k in names}
            if ex0 is not None:
                acc_t["exchange"] = 0.0
        prev = "begin"
        for k in names:
            acc_t[k] += evs[prev].elapsed_time(evs[k])
            prev = k
        if ex0 is not None:
            acc_t["exchange"] += ex0.elapsed_time(ex1)
        isect_sum += int(ws.status[0])
        keys_sum += int(ws.status[6])
    kern_ms = {k: acc_t[k] / reps for k in acc_t}
    I_mean = isect_sum / reps

    # ---------------- end-to-end through the public API with host buffers (`e2e`) ----------------
    loss_host = torch.zeros(1, dtype=torch.float64).pin_memory()
    copy_stream = ctx["copy_stream"]
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

    n_e2e = 4 * args.steps   # a longer region than the device-timed one: wall-clock timing needs it
    e2e_loop(max(2, args.warmup))
    barrier()
    t0 = time.perf_counter()
    e2e_loop(n_e2e)
    barrier()
    t_e2e = time.perf_counter() - t0
    for i in range(V):   # restore slot contents
        step.set_view(i, host_vm[i], host_K[i], host_gt[i])
    torch.cuda.synchronize()

    # ---------------- reductions over ranks ----------------
    vals = torch.tensor([total_ms, t_e2e * 1e3, hot_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, hot_ms = float(vals[0]), float(vals[1]), float(vals[2])
    if rank != 0:
        return None

    peak, peak_src = load_peaks()
    stages, A = algorithmic_bytes(N, I_mean, P)
    ms_per_step = total_ms / args.steps
    value = world * args.steps / (total_ms * 1e-3)
    kernels = [k for k in names if k != "memset"]
    dom = max(kernels, key=lambda k: kern_ms[k])
    achieved = stages[dom] / (kern_ms[dom] * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic_bytes(dom + "_kernel", regime)
    n_launch = ws.n_kernels + (1 if step.exchange is not None else 0)
    out = {
        "workload": workload_name(args, regime), "value": value, "unit": UNIT, "ms_per_step": ms_per_step,
        "n_isects": I_mean, "isect_per_gaussian": I_mean / N, "keys_emitted": keys_sum / reps, "overflow": overflow,
        "pipeline": ws.pipeline, "stopped_tiles": int(ws.status[5]),
        "tile_sort": ("n/a" if ws.pipeline == "splat" else "lazy (only tiles near the transmittance stop threshold)" if model._use_lazy() else
                      ("front-to-back depth slices, early stop" if model.front_sort else "every tile")),
        "tile_culling": bool(model.cull_tiles), "morton_order": not args.no_morton,
        "execution": f"CUDA graph replay per iteration (1 memset + {n_launch} kernels; stages: {', '.join(kernels)}"
                     + (", exchange" if step.exchange is not None else "") + ")",
        "exchange": step.exchange_name(),
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": stages[dom], "kernel_ms": kern_ms[dom]},
        "roofline_step": {"algorithmic_bytes": A, "achieved": A / (ms_per_step * 1e-3) / 1e9,
                          "frac": A / (ms_per_step * 1e-3) / 1e9 / peak, "formula": "228 N + 92 I + 20 P"},
        "kernel_ms": kern_ms,
        "e2e": {"value": world * n_e2e / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": n_e2e, "l2": "not flushed (back-to-back steps, as in training)",
                "api": "GraphedRasterStep.set_view(pinned host) + replay + loss readback"},
        "value_no_l2_flush": world * args.steps / (hot_ms * 1e-3),
        "ms_per_view_slot_rank0": slot_ms,
        "gpu_launches": n_launch * args.steps,
        "clocks": clocks, "wall_s_timed_region": t_wall,
    }
    if parity is not None:
        out["parity"] = parity
        out["cpu_baseline"] = {"value": 1.0 / cpu_s, "unit": UNIT, "cores": cpu_cores, "kind": "port",
                               "sample": f"1 full-size iteration ({N} Gaussians, {W}x{H}, regime={regime}) of the C/OpenMP oracle, "
                                         f"{cpu_s:.2f} s, no sub-sampling; the same run is the parity check"}
    if sum_check is not None:
        out["exchange_sum_check"] = sum_check
    return out


def aux_timings(args, ctx):
    """Optimizer step, regularisers and KNN at N Gaussians: this library's kernels next to the reference's own code
    (torch.optim.Adam x 4 on the same GPU -- the reference's optimizer, utils/train_utils.py:48-65; the numpy ports of
    compute_direction_loss / compute_ratio_loss pinned to the reference's outputs; sklearn NearestNeighbors exactly as
    k_nearest_sklearn calls it, edge_gs.py:135-151).  SURVEY.md section 8d: reported separately from the step."""
    import torch
    from edgegaussians_b200 import synth
    from edgegaussians_b200.edge_gs import EdgeGaussianSplatting
    from edgegaussians_b200.knn import knn_indices
    from edgegaussians_b200.optim import NAMES, FusedAdamGroup
    from edgegaussians_b200.regularisers import _run as reg_run
    dev, N = ctx["dev"], args.n
    m, q, s, o = synth.make_gaussians(N, "trained", 0)
    model = EdgeGaussianSplatting(device=dev)
    model.set_params(m, s, q, o)
    out = {"n": N}

    def timed(fn, reps=10):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # ---- KNN (k = 5 as configs/DTU.json:72)
    k = 5
    pts = model.means.data
    out["knn_ms"] = timed(lambda: knn_indices(pts, k), reps=3)
    nn_idx = knn_indices(pts, k)
    # ---- regularisers fwd + bwd in one pass
    out["reg_fwd_bwd_ms"] = timed(lambda: reg_run(model.means.data, model.quats.data, model.scales.data, nn_idx, k, False, 1.0, 1.0))
    # ---- Adam: one launch (eg_adam_multi) vs the reference's four torch.optim.Adam on the same GPU
    for nm in NAMES:
        model.gauss_params[nm].grad = torch.randn_like(model.gauss_params[nm]) * 1e-3
    group = FusedAdamGroup(model, {nm: 1e-3 for nm in NAMES})
    out["adam_ms"] = timed(lambda: group.step(zero_grad=False))
    ref_params = [torch.nn.Parameter(model.gauss_params[k].detach().clone()) for k in NAMES]
    for p in ref_params:
        p.grad = torch.randn_like(p) * 1e-3
    opts = [torch.optim.Adam([p], lr=1e-3) for p in ref_params]

    def torch_adam():
        for op in opts:
            op.step()
    out["adam_ms_reference_torch_x4_gpu"] = timed(torch_adam)
    if not args.no_cpu_baseline:
        from oracle import reference_ports as ports
        try:
            from sklearn.neighbors import NearestNeighbors
            t0 = time.perf_counter()
            nn_model = NearestNeighbors(n_neighbors=k + 2, algorithm="auto", metric="euclidean").fit(m)
            _, ind = nn_model.kneighbors(m)
            out["knn_s_reference_sklearn_cpu"] = time.perf_counter() - t0
            ref_idx = ind[:, 2:]
            out["knn_mismatch_rows_vs_sklearn"] = int((nn_idx.cpu().numpy() != ref_idx).any(axis=1).sum())
        except Exception as e:  # sklearn missing on the box: the timing is informational
            out["knn_s_reference_sklearn_cpu"] = f"unavailable: {e}"
            ref_idx = nn_idx.cpu().numpy()
        t0 = time.perf_counter()
        ports.direction_loss(m, q, s, ref_idx, k)
        ports.ratio_loss(s)
        out["reg_fwd_bwd_s_reference_port_cpu"] = time.perf_counter() - t0
        out["cpu_cores"] = os.cpu_count()
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    ctx = {"world": world, "rank": rank, "local": local, "dev": dev,
           "flush_buf": None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev),
           "copy_stream": torch.cuda.Stream()}
    regimes = ["init", "trained"] if args.regime == "both" else [args.regime]
    results = {}
    for regime in regimes:
        results[regime] = bench_regime(args, regime, ctx)
        torch.cuda.synchronize()
    aux = None
    if rank == 0 and world == 1 and not args.no_aux:
        try:
            aux = aux_timings(args, ctx)
        except Exception as e:   # informational block: never takes the headline down
            aux = {"failed": repr(e)}
    if rank == 0:
        head = results[regimes[0]]
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic (synth-v1: uniform Gaussians in [-1,1]^3, Fibonacci-sphere cameras, random-segment edge maps)",
            "config": {"workload": head["workload"], "views_cycled_per_gpu": args.views,
                       "n_isects": head["n_isects"], "isect_per_gaussian": head["isect_per_gaussian"], "overflow": head["overflow"],
                       "l2": "flushed between timed iterations (256 MiB fill)" if ctx["flush_buf"] is not None else "not flushed",
                       "pipeline": head["pipeline"], "stopped_tiles": head["stopped_tiles"], "tile_sort": head["tile_sort"],
                       "execution": head["execution"], "exchange": head["exchange"],
                       "parallelism": f"view-sharded dp{world}" if world > 1 else "single GPU"},
            "roofline": head["roofline"], "roofline_step": head["roofline_step"], "kernel_ms": head["kernel_ms"],
            "e2e": head["e2e"], "value_no_l2_flush": head["value_no_l2_flush"], "gpu_launches": head["gpu_launches"],
            "clocks": head["clocks"], "wall_s_timed_region": head["wall_s_timed_region"],
        }
        for k in ("parity", "cpu_baseline", "exchange_sum_check"):
            if k in head:
                line[k] = head[k]
        line["regimes"] = {r: {k: v for k, v in res.items() if k not in ("clocks",)} for r, res in results.items()}
        if aux is not None:
            line["aux_ms"] = aux
        print(json.dumps(line), flush=True)
        bad = [r for r, res in results.items() if ("parity" in res and not res["parity"]["ok"])
               or ("exchange_sum_check" in res and not res["exchange_sum_check"]["ok"])]
        if bad:
            sys.stderr.write(f"bench.py: PARITY VIOLATION in regime(s) {bad}: the numbers above are invalid\n")
            if world > 1:
                dist.destroy_process_group()
            sys.exit(3)
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
