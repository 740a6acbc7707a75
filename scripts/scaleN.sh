#!/bin/bash
# On an N-GPU box: the driver's launch line for the view-sharded bench.   scripts/scaleN.sh 4
N=${1:-2}
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 40 --warmup 5 2>&1 | tee gpurun_out/scale_n$N.log | grep '^{' > gpurun_out/scale_n$N.json
python -c "
import json; d = json.load(open('gpurun_out/scale_n$N.json')); print('gpus', d['n_gpus'], 'it/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1))" || tail -5 gpurun_out/scale_n$N.log
