#!/bin/bash
# BASELINE.json configs[4]: synthetic sweep 10k -> 2M Gaussians x 1920x1080 (run on the GPU box; one JSON line per size
# into gpurun_out/sweep_n<N>.json).   scripts/sweep_sizes.sh [gpus]
G=${1:-1}
for n in 10000 30000 100000 300000 1000000 2000000; do
  if [ "$G" -gt 1 ]; then
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29577 \
      bench.py --gpus $G --n $n --width 1920 --height 1080 --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | grep '^{' > gpurun_out/sweep_g${G}_n$n.json
  else
    python bench.py --n $n --width 1920 --height 1080 --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/sweep_g1_n$n.json
  fi
  python -c "
import json; d = json.load(open('gpurun_out/sweep_g${G}_n$n.json')); print('N', $n, 'gpus', d['n_gpus'], 'it/s', round(d['value'], 1), 'ms', round(d['ms_per_step'], 4), 'step roofline', round(d['roofline_step']['frac'], 4), d['config']['pipeline'][:12])"
done
