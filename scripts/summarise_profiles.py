"""Summarise gpurun_out/{launches,prof}_<tag> into profiles/<tag>_*.{csv,txt} (tracked).
   python scripts/summarise_profiles.py r2            # init regime (launch list of bench.py, capture of profile_step.py)
   python scripts/summarise_profiles.py r2_trained    # trained regime (both from scripts/profile_step.py --regime trained)"""
import csv
import io
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
os.makedirs(pr, exist_ok=True)
trained = "trained" in tag
cmd_launch = ("python scripts/profile_step.py --regime trained --iters 4 (--profile-from-start off: steady-state iterations only)"
              if trained else "python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aux")
cmd_full = f"python scripts/profile_step.py {'--regime trained ' if trained else ''}--iters 1"
hot = ("raster_fwd", "raster_bwd") if trained else ("splat_fwd", "splat_bwd")

# ---- launch list: per-kernel totals and shares ----
lpath = os.path.join(go, f"launches_{tag}.csv")
if os.path.exists(lpath):
    rows = list(csv.reader(open(lpath)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[h]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = {}
    with open(os.path.join(pr, f"{tag}_launches.csv"), "w") as f:
        f.write("id,kernel,gpu__time_duration_ns\n")
        for r in rows[h + 1:]:
            if len(r) <= vi:
                continue
            name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:70]
            ns = float(r[vi].replace(",", ""))
            f.write(f"{r[0]},{name},{ns:.0f}\n")
            agg.setdefault(name, [0, 0.0])
            agg[name][0] += 1
            agg[name][1] += ns
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(pr, f"{tag}_launch_shares.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, command: {cmd_launch}\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"{'kernel':72s} {'launches':>8s} {'total_us':>10s} {'avg_us':>9s} {'share':>7s}\n")
        for name in sorted(agg, key=lambda n: -agg[n][1]):
            c, ns = agg[name]
            f.write(f"{name:72s} {c:8d} {ns / 1e3:10.1f} {ns / 1e3 / c:9.2f} {100 * ns / tot:6.1f}%\n")

# ---- full-set capture: key metrics per kernel ----
rep = os.path.join(go, f"prof_{tag}.ncu-rep")
if os.path.exists(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rws = list(csv.reader(io.StringIO(out)))
    Hh = rws[0]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__grid_size", "launch__block_size",
            "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_issued.avg.per_cycle_active",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
    idx = [(w, Hh.index(w)) for w in want if w in Hh]
    units = rws[1]
    with open(os.path.join(pr, f"{tag}_ncu_full_summary.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on --profile-from-start off, command: {cmd_full}\n")
        f.write(f"# (eager launches of one steady-state fused raster iteration, N=500k x 1600x1200, {'trained' if trained else 'init'} regime)\n")
        for r in rws[2:]:
            f.write("\n")
            for w, i in idx:
                f.write(f"{w:75s} {r[i][:90]} {units[i]}\n")
    for kern in hot:
        o = subprocess.run([sys.executable, os.path.join(root, "scripts", "ncu_lines.py"), rep, kern + "_kernel", "40", "ins"],
                           capture_output=True, text=True).stdout
        with open(os.path.join(pr, f"{tag}_{kern}_source_hotspots.txt"), "w") as f:
            f.write(f"# per-source-line share of warp-level instructions executed / stall samples ({kern}_kernel)\n" + o)
print("written:", sorted(x for x in os.listdir(pr) if x.startswith(tag)))
