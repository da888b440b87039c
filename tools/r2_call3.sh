#!/bin/bash
# Round-2 GPU call 3: tile force kernel after the prologue / prefetch rework
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S=gpurun_out/c3_summary.txt
echo "== 1. GPU parity (engine + fullsize)" | tee $S
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_zfullsize.py -m gpu -q -x 2>&1 | tail -5 | tee -a $S
echo "== 2. bench tiles" | tee -a $S
for w in 0 8 10 16; do
  echo "-- MDG_TILE_WARPS=$w" | tee -a $S
  MDG_TILE_WARPS=$w timeout 300 python bench.py --steps 600 --warmup 100 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.1f steps/s force %.2f us' % (r['value'], 1e3*r['roofline']['kernel_ms']))" | tee -a $S
done
echo "== 3. ncu" | tee -a $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 300 --csv --log-file gpurun_out/c3_launches.csv \
    python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_force_tiles" -s 100 -c 1 \
    -o gpurun_out/c3_prof_tiles python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ls -la gpurun_out | tail -5 | tee -a $S
