#!/bin/bash
# Final single-GPU evidence of round 2: smoke, GPU test suite, default bench line, reference arm, BASELINE configs 3-5
# sweep, ncu launch lists + full-set captures of both regimes.
run() { local t=$1; shift; setsid timeout --kill-after=3 $t "$@" & local pid=$!; wait $pid; local rc=$?; kill -9 -- -$pid 2>/dev/null; return $rc; }
mkdir -p gpurun_out
run 200 python -c 'import __graft_entry__ as g; g.smoke()' || { echo SMOKE FAILED; exit 1; }
run 600 python -m pytest tests -m gpu -q --maxfail=4 -p no:cacheprovider > gpurun_out/r2_tests_final.log 2>&1; tail -3 gpurun_out/r2_tests_final.log
run 300 python bench.py 2> gpurun_out/r2_bench_n1.err | tail -1 > gpurun_out/r2_bench_n1.json; python scripts/show_bench.py gpurun_out/r2_bench_n1.json
run 200 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/r2_bench_reference.err | tail -1 > gpurun_out/r2_bench_reference.json; cut -c1-300 gpurun_out/r2_bench_reference.json
run 400 python scripts/sweep.py --gpus 1 > gpurun_out/r2_sweep_g1.md 2> gpurun_out/r2_sweep_g1.err; cat gpurun_out/r2_sweep_g1.md
run 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aux --regime init > gpurun_out/bench_under_ncu_r2.log 2>&1
run 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"project_fwd_kernel|splat_fwd_chunks_kernel|splat_resolve4_kernel|emit_flagged_kernel|raster_fwd_flagged_kernel|splat_bwd_kernel" -c 6 \
    -o gpurun_out/prof_r2 -f python scripts/profile_step.py --iters 1 > gpurun_out/prof_r2.log 2>&1
run 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r2_trained.csv \
    python scripts/profile_step.py --regime trained --iters 4 > gpurun_out/launches_r2_trained.log 2>&1
run 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"project_fwd_kernel|scan_kernel|raster_fwd_kernel|raster_bwd_kernel|project_bwd_kernel" -c 5 \
    -o gpurun_out/prof_r2_trained -f python scripts/profile_step.py --regime trained --iters 1 > gpurun_out/prof_r2_trained.log 2>&1
tail -n 1 gpurun_out/prof_r2.log gpurun_out/prof_r2_trained.log
