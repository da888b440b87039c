#!/bin/bash
# 2-GPU call: in-kernel stamps of the push / wait kernels, push block count sweep
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S=gpurun_out/c21_summary.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
line() { python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.1f raw %.1f steps/s, %.1f us/step, launches %d, parity %s' % (r['value'], r['config']['box_steps_per_s'], 1e3*r['ms_per_step'], r['gpu_launches'], {k:(r.get('dist_parity') or {}).get(k) for k in ('dv','dq','dpv','dE')}))"; }
echo "== dist_check" | tee $S
timeout 300 $TR --master-port 29511 tests/dist_check.py 2>&1 | grep "dist_check" | tee -a $S
rm -f gpurun_out/tl21_*
for pb in 0 4 16; do
  echo "== N=2 ncell=32 push_blocks=$pb 300 steps + stamps" | tee -a $S
  MDG_DIST_PUSH_BLOCKS=$pb MDG_TIMELINE=gpurun_out/tl21_pb${pb}_ MDG_TIMELINE_STEPS=100000:100001 timeout 300 $TR --master-port 2953$((pb%10)) bench.py --gpus 2 --steps 300 --warmup 50 --ncell 32 --no-c4 --no-e2e --no-dist-parity 2> gpurun_out/c21_pb$pb.err | tee gpurun_out/c21_pb$pb.json | line | tee -a $S
  grep "^#" gpurun_out/tl21_pb${pb}_0.txt | tail -3 | tee -a $S
  grep "^#" gpurun_out/tl21_pb${pb}_1.txt | tail -3 | tee -a $S
done
echo "== N=2 ncell=40 (adaptive gate) 600 steps" | tee -a $S
timeout 400 $TR --master-port 29541 bench.py --gpus 2 --steps 600 --warmup 100 --ncell 40 --no-c4 2> gpurun_out/c21_n2_40.err | tee gpurun_out/c21_n2_40.json | line | tee -a $S
