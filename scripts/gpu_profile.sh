#!/bin/bash
# Run on the GPU box (under gpurun): ncu launch list of the bench command + one full-set capture of the
# raster kernels; results land in gpurun_out/ and are summarised into profiles/ by scripts/summarise_profiles.py.
set -x
TAG=${1:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"project_fwd_kernel|splat_fwd_kernel|splat_resolve4_kernel|emit_flagged_kernel|raster_fwd_flagged_kernel|splat_bwd_kernel" -s 6 -c 6 \
    -o gpurun_out/prof_${TAG} python scripts/profile_step.py --iters 3 > gpurun_out/prof_${TAG}.log 2>&1
tail -2 gpurun_out/prof_${TAG}.log
