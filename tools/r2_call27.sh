#!/bin/bash
# 1-GPU call: final default (paired phase 1) - full GPU suite, smoke, bench lines; builder variants with their own parity run
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S=gpurun_out/c27_summary.txt
line() { python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=r.get('e2e') or {}; c4=r.get('c4') or {}
print('value %.1f steps/s, %.1f us/step, launches %d, e2e %s, force %.2f us frac %.3f, c4 %s, clocks %s' % (r['value'], 1e3*r['ms_per_step'], r['gpu_launches'], e.get('value'), 1e3*((r.get('roofline') or {}).get('kernel_ms') or 0), (r.get('roofline') or {}).get('frac') or 0, c4.get('box_steps_per_s'), r.get('clocks')))"; }
echo "== full GPU suite" | tee $S
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee -a $S
echo "== smoke" | tee -a $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1 | tee -a $S
echo "== default, 1000 steps" | tee -a $S
timeout 400 python bench.py --steps 1000 --warmup 100 2> gpurun_out/c27_n1_1000.err | tee gpurun_out/c27_n1_1000.json | line | tee -a $S
echo "== default, driver style" | tee -a $S
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/c27_n1_20.err | tee gpurun_out/c27_n1_20.json | line | tee -a $S
for v in fbp4 fbp64 fbp4w2; do
  echo "== variant $v: engine parity tests + 1000 steps" | tee -a $S
  MDG_LIB_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_zfullsize.py -x -q -m gpu 2>&1 | tail -1 | tee -a $S
  MDG_LIB_VARIANT=$v timeout 300 python bench.py --steps 1000 --warmup 100 --no-c4 --no-e2e 2> gpurun_out/c27_$v.err | tee gpurun_out/c27_$v.json | line | tee -a $S
done
echo "== builder launch times (default)" | tee -a $S
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_build_fast' -c 8 --csv --log-file gpurun_out/c27_build.csv python bench.py --steps 30 --warmup 5 --no-c4 --no-e2e > /dev/null 2>&1
python - <<'PY' | tee -a $S
import csv
t=[float(r[-1])/1e3 for r in csv.reader(open('gpurun_out/c27_build.csv')) if len(r)>10 and r[0].isdigit()]
print("k_build_fast n=%d mean %.2f us min %.2f" % (len(t), sum(t)/len(t), min(t)))
PY
