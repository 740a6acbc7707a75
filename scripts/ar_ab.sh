#!/bin/bash
# A/B of the per-step gradient all-reduce: libedgegs communicator on the compute stream vs torch.distributed.
N=${1:-2}
for mode in "--native-allreduce" ""; do
  timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 40 --warmup 5 $mode > gpurun_out/ar_ab.log 2>&1
  grep '^{' gpurun_out/ar_ab.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.readline()); print('N=$N', '${mode:-torch}', 'it/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1))
except Exception as e: print('${mode:-torch} failed')" || tail -5 gpurun_out/ar_ab.log
done
tail -3 gpurun_out/ar_ab.log | cut -c1-200 | grep -v '^{'
