#!/bin/bash
mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5) > gpurun_out/r2_push_tests_d.log 2>&1
tail -3 gpurun_out/r2_push_tests_d.log
(timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5) > gpurun_out/r2_parity_d.log 2>&1
tail -3 gpurun_out/r2_parity_d.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 200 $TR scripts/exchange_micro.py 2> gpurun_out/r2_exmicro_2d.err | grep '^{' > gpurun_out/r2_exmicro_2d_leader.json; cat gpurun_out/r2_exmicro_2d_leader.json
EG_AR_BARRIER=cta timeout 200 $TR scripts/exchange_micro.py 2> gpurun_out/r2_exmicro_2d.err | grep '^{' > gpurun_out/r2_exmicro_2d_cta.json; cat gpurun_out/r2_exmicro_2d_cta.json
timeout 200 $TR bench.py --gpus 2 --steps 40 --warmup 5 --regime init --no-aux --no-cpu-baseline --exchange push-p2p 2> gpurun_out/r2_n2d.err | grep '^{' > gpurun_out/r2_n2d_push-p2p.json
python scripts/show_bench.py gpurun_out/r2_n2d_push-p2p.json
timeout 200 python bench.py --steps 40 --warmup 5 --no-aux 2> gpurun_out/r2_n1d.err | grep '^{' > gpurun_out/r2_n1d.json
python scripts/show_bench.py gpurun_out/r2_n1d.json
