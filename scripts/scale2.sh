#!/bin/bash
# On a 2-GPU box: single-GPU reference point + 2-GPU view-sharded runs with different all-reduce chunkings.
show() { tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.readline()); print('$1', 'gpus', d['n_gpus'], 'it/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['config']['execution'][-90:])
except Exception as e: print('$1 failed', e)"; }
python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>&1 | show n1
for k in ${CHUNKS:-1 2 4 8}; do
  timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + k)) bench.py --gpus 2 --steps 40 --warmup 5 --allreduce-chunks $k 2>&1 | tee gpurun_out/n2_chunks$k.log | grep '^{' | show "n2/chunks=$k"
done
