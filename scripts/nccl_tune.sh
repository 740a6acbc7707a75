#!/bin/bash
# On an N-GPU box: the view-sharded bench under a few NCCL settings.   scripts/nccl_tune.sh 4
N=${1:-4}
run() { tag=$1; shift; env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/nccl_$tag.log 2>&1
  grep '^{' gpurun_out/nccl_$tag.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.readline()); print('$tag', 'it/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4))
except Exception as e: print('$tag failed')"; }
run default NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING
grep -E "NVLS|Channel|algo|Algo|nChannels|comm .* nranks" gpurun_out/nccl_default.log | grep -v "Ring\b.*->" | head -12
run minch32 NCCL_MIN_NCHANNELS=32
run nvls NCCL_ALGO=NVLS
run ring_simple NCCL_ALGO=Ring NCCL_PROTO=Simple
