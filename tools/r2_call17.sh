#!/bin/bash
# 2-GPU call: leaner list builder (phase 2) on one GPU; boundary-layer side stream on two GPUs (on / off), driver-style 20-step lines
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S=gpurun_out/c17_summary.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
line() { python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=r.get('e2e') or {}; c4=r.get('c4') or {}
print('value %.1f raw %.1f steps/s, %.1f us/step, launches %d, e2e %s, force %s us, c4 %s, parity %s' % (r['value'], r['config']['box_steps_per_s'], 1e3*r['ms_per_step'], r['gpu_launches'], e.get('value'), 1e3*((r.get('roofline') or {}).get('kernel_ms') or 0), c4.get('value'), r.get('dist_parity')))"; }
echo "== engine GPU tests" | tee $S
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py -x -q -m gpu 2>&1 | tail -3 | tee -a $S
echo "== N=1 1000 steps" | tee -a $S
timeout 400 python bench.py --steps 1000 --warmup 100 2> gpurun_out/c17_n1_1000.err | tee gpurun_out/c17_n1_1000.json | line | tee -a $S
echo "== N=1 driver style" | tee -a $S
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/c17_n1_20.err | tee gpurun_out/c17_n1_20.json | line | tee -a $S
echo "== builder launch times" | tee -a $S
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_build_fast|k_scan_one|k_cellsort_warp|k_bin|k_scatter' -c 40 --csv --log-file gpurun_out/c17_build_launches.csv python bench.py --steps 30 --warmup 5 --no-c4 > gpurun_out/c17_ncu_bench.log 2>&1
python - <<'PY' | tee -a $S
import csv,collections
t=collections.defaultdict(list)
for r in csv.reader(open('gpurun_out/c17_build_launches.csv')):
    if len(r)>10 and r[0].isdigit(): t[r[4].split('(')[0]].append(float(r[-1])/1e3)
for k,v in t.items(): print("%-24s n=%3d mean %8.2f us min %8.2f" % (k,len(v),sum(v)/len(v),min(v)))
PY
echo "== dist_check p2p (bnd stream on)" | tee -a $S
timeout 300 $TR --master-port 29511 tests/dist_check.py 2>&1 | grep -v "^W\|^\[W\|NCCL version" | tail -4 | tee -a $S
echo "== dist_check nccl (bnd stream on)" | tee -a $S
MDG_DIST_P2P=0 timeout 300 $TR --master-port 29512 tests/dist_check.py 2>&1 | grep -v "^W\|^\[W\|NCCL version" | tail -4 | tee -a $S
echo "== N=2 driver style" | tee -a $S
timeout 500 $TR --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/c17_n2_20.err | tee gpurun_out/c17_n2_20.json | line | tee -a $S
for nc in 32 40; do
 for bnd in 1 0; do
  echo "== N=2 ncell=$nc bnd_stream=$bnd 600 steps" | tee -a $S
  MDG_DIST_BND_STREAM=$bnd timeout 400 $TR --master-port 2952$bnd bench.py --gpus 2 --steps 600 --warmup 100 --ncell $nc --no-c4 2> gpurun_out/c17_n2_${nc}_${bnd}.err | tee gpurun_out/c17_n2_${nc}_${bnd}.json | line | tee -a $S
 done
done
