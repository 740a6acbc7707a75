#!/bin/bash
# Final multi-GPU evidence on an N-GPU box: scripts/gpu_r2_final2.sh N [tests]
run() { local t=$1; shift; setsid timeout --kill-after=3 $t "$@" & local pid=$!; wait $pid; local rc=$?; kill -9 -- -$pid 2>/dev/null; return $rc; }
N=${1:-2}
mkdir -p gpurun_out
if [ "$2" == "tests" ]; then
  run 400 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_multi_tests_final_n$N.log 2>&1; tail -3 gpurun_out/r2_multi_tests_final_n$N.log
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
run 250 $TR bench.py --gpus $N --steps 40 --warmup 5 2> gpurun_out/r2_scale_n$N.err | grep '^{' > gpurun_out/r2_scale_n$N.json
python scripts/show_bench.py gpurun_out/r2_scale_n$N.json
run 150 $TR scripts/exchange_micro.py 2> gpurun_out/r2_exmicro_$N.err | grep '^{' > gpurun_out/r2_exmicro_$N.json; cat gpurun_out/r2_exmicro_$N.json
