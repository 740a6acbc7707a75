#!/bin/bash
# N=1 A/B of compile-time variants of eg_splat_bwd (scripts/build_variant.sh): parity suite on the default build, then
# the init-regime bench line of every build.
mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4) > gpurun_out/r2_parity_e.log 2>&1
tail -2 gpurun_out/r2_parity_e.log
for v in default $*; do
  LIB=""; [ "$v" != default ] && LIB=edgegaussians_b200/_C/libedgegs_$v.so
  EG_LIB=$LIB timeout 200 python bench.py --steps 40 --warmup 5 --regime init --no-aux --no-cpu-baseline 2> gpurun_out/r2_ab_$v.err | grep '^{' > gpurun_out/r2_ab_$v.json
  python scripts/show_bench.py gpurun_out/r2_ab_$v.json
done
