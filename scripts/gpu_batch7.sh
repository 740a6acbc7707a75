#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider 2>&1 | tail -30) > gpurun_out/r2_tests_7.log 2>&1
tail -3 gpurun_out/r2_tests_7.log
(timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1) > gpurun_out/r2_b7_default.json
(timeout 200 python bench.py --steps 20 --warmup 5 --no-aux --no-cpu-baseline --no-morton 2>&1 | tail -1) > gpurun_out/r2_b7_nomorton.json
(EG_BWD_OPTS=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-aux --no-cpu-baseline --regime init 2>&1 | tail -1) > gpurun_out/r2_b7_noadaptive.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"raster_fwd_kernel|raster_bwd_kernel" -c 2 -o gpurun_out/prof_r2_trained -f python scripts/profile_step.py --regime trained --iters 3 > gpurun_out/prof_r2_trained.log 2>&1
tail -2 gpurun_out/prof_r2_trained.log
