#!/bin/bash
# N-GPU call: weak scaling at 256 000 atoms/GPU (driver's SCALE config) and at the north_star's C4 size (131 072 atoms/GPU),
# peer-to-peer step path vs NCCL path, with the dist_parity statement in every line.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
N=${1:-8}
S=gpurun_out/s${N}_summary.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== scale N=$N" | tee $S
for nc in 40 32; do
 for p2p in 1 0; do
  echo "-- ncell=$nc p2p=$p2p" | tee -a $S
  MDG_DIST_P2P=$p2p timeout 500 $TR --master-port 2953$p2p bench.py --gpus $N --steps 600 --warmup 100 --ncell $nc 2> gpurun_out/s${N}_bench_${nc}_${p2p}.err | tee gpurun_out/s${N}_bench_${nc}_${p2p}.json | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.1f box-eq steps/s, raw %.1f steps/s, %.1f us/step, launches %d, parity %s' % (r['value'], r['config']['box_steps_per_s'], 1e3*r['ms_per_step'], r['gpu_launches'], {k:r['dist_parity'][k] for k in ('dv','dq','dpv','dE')}))" | tee -a $S
 done
done
