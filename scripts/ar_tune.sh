#!/bin/bash
# Exchange-kernel tuning on N GPUs (run under gpurun --gpus N): bytes in flight (unroll), grid, multicast vs peer, NCCL.
N=${1:-2}
mkdir -p gpurun_out
run() {  # tag, env..., exchange
  tag=$1; ex=$2; shift 2
  (env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus $N --steps 30 --warmup 5 --regime init --no-aux --no-cpu-baseline --exchange $ex $EXTRA 2>gpurun_out/r2_ar_${N}_$tag.err | grep "^{") > gpurun_out/r2_ar_${N}_$tag.json
  tail -c 200 gpurun_out/r2_ar_${N}_$tag.err | grep -i "error\|Traceback" || true
}
(timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -p no:cacheprovider -k "ranged" 2>&1 | tail -4) > gpurun_out/r2_ar_${N}_tests.log; cat gpurun_out/r2_ar_${N}_tests.log
EXTRA="--exchange-ranges 2" run p2p_r2 symm-p2p X=1
EXTRA="--exchange-ranges 4" run p2p_r4 symm-p2p X=1
EXTRA="--exchange-ranges 4" run p2p_r4_g32 symm-p2p EG_AR_GRID_RANGED=32
EXTRA="--exchange-ranges 4" run symm_r4 symm X=1
EXTRA="--exchange-ranges 8" run p2p_r8 symm-p2p X=1
run p2p symm-p2p X=1
run nccl nccl X=1
ls gpurun_out | grep r2_ar_${N}_ | grep json
