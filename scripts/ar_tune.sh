#!/bin/bash
# Exchange-kernel tuning on N GPUs (run under gpurun --gpus N): bytes in flight (unroll), grid, multicast vs peer, NCCL.
N=${1:-2}
mkdir -p gpurun_out
run() {  # tag, env..., exchange
  tag=$1; ex=$2; shift 2
  (env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus $N --steps 30 --warmup 5 --regime init --no-aux --no-cpu-baseline --exchange $ex 2>gpurun_out/r2_ar_${N}_$tag.err | grep "^{") > gpurun_out/r2_ar_${N}_$tag.json
  tail -c 200 gpurun_out/r2_ar_${N}_$tag.err | grep -i "error\|Traceback" || true
}
run symm_u8_g148 symm EG_AR_UNROLL=8 EG_AR_GRID=148
run symm_u4_g148 symm EG_AR_UNROLL=4 EG_AR_GRID=148
run symm_u8_g64 symm EG_AR_UNROLL=8 EG_AR_GRID=64
run symm_u8_g296 symm EG_AR_UNROLL=8 EG_AR_GRID=296
run p2p_u8_g148 symm-p2p EG_AR_UNROLL=8 EG_AR_GRID=148
run nccl nccl X=1
ls gpurun_out | grep r2_ar_${N}_ | grep json
