#!/bin/bash
# 2-GPU call: pull path vs push path - parity (3x each), timing at 131 072 / 256 000 atoms per GPU, timeline + stamps, driver-style line
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S=gpurun_out/c23_summary.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
line() { python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); c4=r.get('c4') or {}
print('value %.1f raw %.1f steps/s, %.1f us/step, launches %d, c4 %s us/step, parity %s' % (r['value'], r['config']['box_steps_per_s'], 1e3*r['ms_per_step'], r['gpu_launches'], 1e3*c4.get('ms_per_step',0), {k:(r.get('dist_parity') or {}).get(k) for k in ('dv','dq','dpv','dE')}))"; }
echo "== dist_check pull x3" | tee $S
for i in 1 2 3; do timeout 200 $TR --master-port 2951$i tests/dist_check.py 2>&1 | grep "dist_check" | tee -a $S; done
echo "== dist_check push x1" | tee -a $S
MDG_DIST_PULL=0 timeout 200 $TR --master-port 29514 tests/dist_check.py 2>&1 | grep "dist_check" | tee -a $S
for nc in 32 40; do
 for pull in 1 0; do
  echo "== N=2 ncell=$nc pull=$pull 600 steps" | tee -a $S
  MDG_DIST_PULL=$pull timeout 400 $TR --master-port 2952$pull bench.py --gpus 2 --steps 600 --warmup 100 --ncell $nc --no-c4 2> gpurun_out/c23_n2_${nc}_${pull}.err | tee gpurun_out/c23_n2_${nc}_${pull}.json | line | tee -a $S
 done
done
rm -f gpurun_out/tl23_*
MDG_TIMELINE=gpurun_out/tl23_n2_32_ timeout 300 $TR --master-port 29533 bench.py --gpus 2 --steps 60 --warmup 40 --ncell 32 --no-c4 --no-e2e --no-dist-parity > gpurun_out/c23_tl.json 2> gpurun_out/c23_tl.err
grep "^#" gpurun_out/tl23_n2_32_0.txt | tail -2 | tee -a $S
echo "== N=2 driver style" | tee -a $S
timeout 500 $TR --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/c23_n2_20.err | tee gpurun_out/c23_n2_20.json | line | tee -a $S
