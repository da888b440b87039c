#!/bin/bash
# 2-GPU call: peer-to-peer step path (NVLink stores + flags) vs the NCCL path - parity and timing
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-2}
S=gpurun_out/d${N}_summary.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== dist_check p2p" | tee $S
timeout 300 $TR --master-port 29511 tests/dist_check.py 2>&1 | grep -v "^W\|^\[W\|NCCL version" | tail -4 | tee -a $S
echo "== dist_check nccl" | tee -a $S
MDG_DIST_P2P=0 timeout 300 $TR --master-port 29512 tests/dist_check.py 2>&1 | grep -v "^W\|^\[W\|NCCL version" | tail -4 | tee -a $S
for nc in 40 32; do
 for p2p in 1 0; do
  echo "== bench gpus=$N ncell=$nc p2p=$p2p" | tee -a $S
  MDG_DIST_P2P=$p2p timeout 400 $TR --master-port 2952$p2p bench.py --gpus $N --steps 600 --warmup 100 --ncell $nc 2> gpurun_out/d${N}_bench_${nc}_${p2p}.err | tee gpurun_out/d${N}_bench_${nc}_${p2p}.json | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.1f box-eq steps/s, raw %.1f steps/s, %.1f us/step, launches %d, parity %s' % (r['value'], r['config']['box_steps_per_s'], 1e3*r['ms_per_step'], r['gpu_launches'], r.get('dist_parity')))" | tee -a $S
 done
done
