#!/bin/bash
mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5) > gpurun_out/r2_push_tests_c.log 2>&1
tail -3 gpurun_out/r2_push_tests_c.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 200 $TR scripts/exchange_micro.py 2> gpurun_out/r2_exmicro_2c.err | grep '^{' > gpurun_out/r2_exmicro_2c.json; cat gpurun_out/r2_exmicro_2c.json
for ex in push-p2p push symm-p2p; do
  timeout 200 $TR bench.py --gpus 2 --steps 40 --warmup 5 --regime init --no-aux --no-cpu-baseline --exchange $ex 2> gpurun_out/r2_n2c_$ex.err | grep '^{' > gpurun_out/r2_n2c_$ex.json
  python scripts/show_bench.py gpurun_out/r2_n2c_$ex.json
done
