#!/bin/bash
# 8-GPU call (final code): one bench launch = 256 000 atoms per GPU (600 steps) + the configs[3]-size line (131 072 per GPU, 400 steps) + dist_parity
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 420 $TR --master-port 29541 bench.py --gpus $N --steps 600 --warmup 100 2> gpurun_out/c25_n${N}.err | tee gpurun_out/c25_n${N}.json | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); c4=r.get('c4') or {}
print('value %.1f raw %.1f steps/s, %.1f us/step, launches %d; c4: %s box-eq/s %.1f us/step; parity %s' % (r['value'], r['config']['box_steps_per_s'], 1e3*r['ms_per_step'], r['gpu_launches'], c4.get('value'), 1e3*c4.get('ms_per_step',0), {k:(r.get('dist_parity') or {}).get(k) for k in ('dv','dq','dpv','dE')}))" | tee gpurun_out/c25_summary.txt
