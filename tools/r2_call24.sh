#!/bin/bash
# 4-GPU call: the peer-to-peer paths with DISTINCT neighbours below / above (W = 2 has one neighbour on both sides)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S=gpurun_out/c24_summary.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
line() { python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); c4=r.get('c4') or {}
print('value %.1f raw %.1f steps/s, %.1f us/step, launches %d, c4 %s us/step, parity %s' % (r['value'], r['config']['box_steps_per_s'], 1e3*r['ms_per_step'], r['gpu_launches'], 1e3*c4.get('ms_per_step',0), {k:(r.get('dist_parity') or {}).get(k) for k in ('dv','dq','dpv','dE')}))"; }
echo "== dist_check W=4 auto / pull / push" | tee $S
timeout 200 $TR --master-port 29511 tests/dist_check.py 2>&1 | grep "dist_check" | tee -a $S
MDG_DIST_PULL=1 timeout 200 $TR --master-port 29512 tests/dist_check.py 2>&1 | grep "dist_check" | tee -a $S
MDG_DIST_PULL=0 timeout 200 $TR --master-port 29513 tests/dist_check.py 2>&1 | grep "dist_check" | tee -a $S
echo "== N=4 driver style (256k per GPU + c4)" | tee -a $S
timeout 500 $TR --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 2> gpurun_out/c24_n4_20.err | tee gpurun_out/c24_n4_20.json | line | tee -a $S
for nc in 32 40; do
  echo "== N=4 ncell=$nc 600 steps" | tee -a $S
  timeout 400 $TR --master-port 2952$((nc/10)) bench.py --gpus 4 --steps 600 --warmup 100 --ncell $nc --no-c4 2> gpurun_out/c24_n4_${nc}.err | tee gpurun_out/c24_n4_${nc}.json | line | tee -a $S
done
