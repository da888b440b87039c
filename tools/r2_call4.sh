#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S=gpurun_out/c5_summary.txt
echo "== 1. GPU parity (engine + fullsize)" | tee $S
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_zfullsize.py -m gpu -q -x 2>&1 | tail -5 | tee -a $S
echo "== 2. bench tiles (persistent)" | tee -a $S
for cfg in "0 0" "16 0" "8 0" "10 0" "6 0"; do
  set -- $cfg
  echo "-- MDG_TILE_WARPS=$1 MDG_TILE_CTAS=$2" | tee -a $S
  MDG_TILE_WARPS=$1 MDG_TILE_CTAS=$2 timeout 300 python bench.py --steps 600 --warmup 100 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.1f steps/s force %.2f us' % (r['value'], 1e3*r['roofline']['kernel_ms']))" | tee -a $S
done
echo "== 3. ncu" | tee -a $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_force_tiles" -s 100 -c 1 \
    -o gpurun_out/c5_prof_tiles python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ls -la gpurun_out | tail -3 | tee -a $S
