#!/bin/bash
# Round-2 evidence batch (run under gpurun, one GPU): GPU test suite, the default bench line, the ncu launch list of the
# bench command and one full-set capture of the shipped kernels in both regimes.  scripts/summarise_profiles.py r2 digests it.
mkdir -p gpurun_out
if ! timeout 200 python -c 'import __graft_entry__ as g; g.smoke()'; then echo GATE FAILED; exit 1; fi
(timeout 1200 python -m pytest tests -m gpu -q --maxfail=4 -p no:cacheprovider 2>&1 | tail -30) > gpurun_out/r2_tests.log 2>&1
tail -3 gpurun_out/r2_tests.log
(timeout 400 python bench.py 2> gpurun_out/r2_bench_n1.err | tail -1) > gpurun_out/r2_bench_n1.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/bench_under_ncu_r2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:"project_fwd_kernel|splat_fwd_chunks_kernel|splat_resolve4_kernel|emit_flagged_kernel|raster_fwd_flagged_kernel|splat_bwd_kernel" -c 6 \
    -o gpurun_out/prof_r2 -f python scripts/profile_step.py --iters 1 > gpurun_out/prof_r2.log 2>&1
tail -2 gpurun_out/prof_r2.log
timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:"project_fwd_kernel|scan_kernel|raster_fwd_kernel|raster_bwd_kernel|project_bwd_kernel" -c 5 \
    -o gpurun_out/prof_r2_trained -f python scripts/profile_step.py --regime trained --iters 1 > gpurun_out/prof_r2_trained.log 2>&1
tail -2 gpurun_out/prof_r2_trained.log
