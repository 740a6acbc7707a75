#!/bin/bash
# Trained-regime evidence: launch list of the bench command + full-set capture of one steady-state iteration.
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r2_trained.csv \
    python scripts/profile_step.py --regime trained --iters 4 > gpurun_out/launches_r2_trained.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"project_fwd_kernel|scan_kernel|raster_fwd_kernel|raster_bwd_kernel|project_bwd_kernel" -c 5 \
    -o gpurun_out/prof_r2_trained -f python scripts/profile_step.py --regime trained --iters 1 > gpurun_out/prof_r2_trained.log 2>&1
tail -n 2 gpurun_out/prof_r2_trained.log
