#!/bin/bash
# Multi-GPU evidence on an N-GPU box: scripts/gpu_scale.sh N [full]
#   exchange micro-benchmark, init-regime A/B of the exchange forms, the driver's default bench line, BASELINE config 4.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
timeout 240 $TR scripts/exchange_micro.py 2> gpurun_out/r2_exmicro_$N.err | grep '^{' > gpurun_out/r2_exmicro_$N.json; cat gpurun_out/r2_exmicro_$N.json
for ex in push push-p2p nccl; do
  timeout 200 $TR bench.py --gpus $N --steps 40 --warmup 5 --regime init --no-aux --no-cpu-baseline --exchange $ex 2> gpurun_out/r2_n${N}_$ex.err | grep '^{' > gpurun_out/r2_n${N}_$ex.json
  python scripts/show_bench.py gpurun_out/r2_n${N}_$ex.json
done
timeout 300 $TR bench.py --gpus $N --steps 40 --warmup 5 2> gpurun_out/r2_scale_n$N.err | grep '^{' > gpurun_out/r2_scale_n$N.json
python scripts/show_bench.py gpurun_out/r2_scale_n$N.json
if [ "$2" == "full" ]; then
  timeout 200 $TR bench.py --gpus $N --width 1200 --height 680 --steps 40 --warmup 5 --no-aux --no-cpu-baseline 2> gpurun_out/r2_cfg4_n$N.err | grep '^{' > gpurun_out/r2_cfg4_n$N.json
  python scripts/show_bench.py gpurun_out/r2_cfg4_n$N.json
  timeout 200 $TR bench.py --gpus $N --gaussians 2000000 --width 1920 --height 1080 --steps 30 --warmup 5 --no-aux --no-cpu-baseline 2> gpurun_out/r2_cfg5_2m_n$N.err | grep '^{' > gpurun_out/r2_cfg5_2m_n$N.json
  python scripts/show_bench.py gpurun_out/r2_cfg5_2m_n$N.json
fi
