#!/bin/bash
# On the GPU box: rebuild with each -D<NAME>=<v> and run the headline bench.  scripts/sweep_define.sh EG_ROWS_PER_ITEM 1 2 4
NAME=$1; shift
for v in "$@"; do
  EG_NVCC_EXTRA="-D${NAME}=${v}" python -m edgegaussians_b200.build --force > /dev/null 2>&1 || { echo "build failed for $v"; continue; }
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline ${BENCH_ARGS} 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('${NAME}=${v}', round(d['value'],1), round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['kernel_ms'].items()})"
done
python -m edgegaussians_b200.build --force > /dev/null 2>&1
