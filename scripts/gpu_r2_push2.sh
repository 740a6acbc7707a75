#!/bin/bash
# 2-GPU check of the push exchange: tests, exchange micro-benchmark, bench A/B.
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15) > gpurun_out/r2_push_tests.log 2>&1
tail -4 gpurun_out/r2_push_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 200 $TR scripts/exchange_micro.py 2> gpurun_out/r2_exmicro_2.err | grep '^{' > gpurun_out/r2_exmicro_2.json; cat gpurun_out/r2_exmicro_2.json
for ex in push push-p2p symm-p2p; do
  timeout 200 $TR bench.py --gpus 2 --steps 40 --warmup 5 --regime init --no-aux --exchange $ex 2> gpurun_out/r2_n2_$ex.err | grep '^{' > gpurun_out/r2_n2_$ex.json
  python scripts/show_bench.py gpurun_out/r2_n2_$ex.json
done
timeout 300 $TR bench.py --gpus 2 --steps 40 --warmup 5 --no-aux 2> gpurun_out/r2_n2_default.err | grep '^{' > gpurun_out/r2_n2_default.json
python scripts/show_bench.py gpurun_out/r2_n2_default.json
