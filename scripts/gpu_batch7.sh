#!/bin/bash
mkdir -p gpurun_out
# quick gate: stop the batch when the basics are broken (a sticky CUDA error makes everything after it meaningless)
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k 'odd or cfg1' -p no:cacheprovider 2>&1 | tail -5 || true
if ! timeout 200 python -c 'import __graft_entry__ as g; g.smoke()'; then echo GATE FAILED; exit 1; fi
(timeout 900 python -m pytest tests -m gpu -q --maxfail=4 -p no:cacheprovider 2>&1 | tail -30) > gpurun_out/r2_tests_7.log 2>&1
tail -3 gpurun_out/r2_tests_7.log
(timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1) > gpurun_out/r2_b7_default.json
(timeout 200 python bench.py --steps 20 --warmup 5 --no-aux --no-cpu-baseline --no-morton 2>&1 | tail -1) > gpurun_out/r2_b7_nomorton.json
(EG_BWD_OPTS=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-aux --no-cpu-baseline --regime init 2>&1 | tail -1) > gpurun_out/r2_b7_noadaptive.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"raster_fwd_kernel|raster_bwd_kernel" -c 2 -o gpurun_out/prof_r2_trained -f python scripts/profile_step.py --regime trained --iters 3 > gpurun_out/prof_r2_trained.log 2>&1
tail -2 gpurun_out/prof_r2_trained.log
