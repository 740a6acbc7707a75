#!/bin/bash
# A/B of the Gaussian-major backward variants on one B200 (run under gpurun); results in gpurun_out/r2_b4_*.
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-aux --no-cpu-baseline --regime init"
for o in 1 2 3; do (EG_BWD_OPTS=$o timeout 200 $B 2>&1 | tail -1) > gpurun_out/r2_b4_opts$o.json; done
(EG_BWD_MODE=gauss timeout 200 $B 2>&1 | tail -1) > gpurun_out/r2_b4_gauss.json
(EG_BWD_MODE=gauss timeout 200 $B --morton 2>&1 | tail -1) > gpurun_out/r2_b4_gauss_morton.json
(EG_BWD_OPTS=3 timeout 200 $B --morton 2>&1 | tail -1) > gpurun_out/r2_b4_opts3_morton.json
(timeout 200 python bench.py --steps 20 --warmup 5 --no-aux --regime trained 2>&1 | tail -1) > gpurun_out/r2_b4_trained.json
(EG_BWD_MODE=gauss timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_raster_step or odd or ranges" -p no:cacheprovider 2>&1 | tail -3) > gpurun_out/r2_b4_gauss_tests.log
(EG_BWD_OPTS=3 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_raster_step or odd or ranges or backward_parity" -p no:cacheprovider 2>&1 | tail -3) > gpurun_out/r2_b4_opts3_tests.log
cat gpurun_out/r2_b4_*tests.log
