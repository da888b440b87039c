#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for i in 1 2 3; do
timeout 300 $TR --master-port 2951$i tests/dist_check.py > gpurun_out/dd_check_$i.log 2>&1; echo "rc=$?" >> gpurun_out/dd_check_$i.log
grep -h "dist_check\|Error\|error\|rc=" gpurun_out/dd_check_$i.log | head -8
done
for nc in 40 32; do
  MDG_DIST_P2P=1 timeout 400 $TR --master-port 29521 bench.py --gpus $N --steps 600 --warmup 100 --ncell $nc 2> gpurun_out/dd_bench_${nc}.err | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.1f box-eq steps/s, %.1f us/step, parity %s' % (r['value'], 1e3*r['ms_per_step'], {k:r['dist_parity'][k] for k in ('dv','dq','dpv','dE')}))"
done
