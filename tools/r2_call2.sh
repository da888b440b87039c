#!/bin/bash
# Round-2 GPU call 2: first hardware run of the tile list (tiles.cuh): parity subset, A/B against the row list, ncu.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S=gpurun_out/c2_summary.txt
echo "== 1. GPU parity (engine + fullsize)" | tee $S
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_zfullsize.py tests/test_gpu_api.py -m gpu -q -x 2>&1 | tail -15 | tee -a $S
echo "== 2. bench tiles (default) vs rows" | tee -a $S
timeout 300 python bench.py --steps 1000 --warmup 200 --no-e2e --no-cpu-baseline > gpurun_out/c2_bench_tiles.json 2> gpurun_out/c2_bench_tiles.err
tail -c 1800 gpurun_out/c2_bench_tiles.json | tee -a $S
MDG_TILES=0 timeout 300 python bench.py --steps 1000 --warmup 200 --no-e2e --no-cpu-baseline > gpurun_out/c2_bench_rows.json 2> gpurun_out/c2_bench_rows.err
tail -c 1800 gpurun_out/c2_bench_rows.json | tee -a $S
for w in 8 10 14 16; do
  echo "-- MDG_TILE_WARPS=$w" | tee -a $S
  MDG_TILE_WARPS=$w timeout 300 python bench.py --steps 600 --warmup 100 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.1f steps/s force %.2f us' % (r['value'], 1e3*r['roofline']['kernel_ms']))" | tee -a $S
done
echo "== 3. ncu" | tee -a $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 300 --csv --log-file gpurun_out/c2_launches.csv \
    python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_force_tiles|k_build_tiles" -s 300 -c 4 \
    -o gpurun_out/c2_prof_tiles python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ls -la gpurun_out | tee -a $S
