#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S=gpurun_out/c8_summary.txt
echo "== 1. full GPU parity suite" | tee $S
MDG_TEST_TC=1 timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee -a $S
echo "== 2. row-kernel build variants" | tee -a $S
for v in "" pf pfmb6 mb6 u2mb6; do
  MDG_LIB_VARIANT=$v timeout 200 python bench.py --steps 600 --warmup 60 --no-e2e --no-cpu-baseline 2> /dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('variant %-6s %.1f steps/s force %.2f us' % ('$v', r['value'], 1e3*r['roofline']['kernel_ms']))" | tee -a $S
done
echo "== 3. e2e profile" | tee -a $S
BENCH_PROFILE_E2E=1 timeout 300 python bench.py --steps 1000 --warmup 100 --no-cpu-baseline > gpurun_out/c8_bench.json 2> gpurun_out/c8_e2e_prof.txt
tail -c 600 gpurun_out/c8_bench.json | tee -a $S
echo "== 4. configs c1 c3 c5 (SIMT / TC)" | tee -a $S
timeout 300 python bench.py --config c1 --steps 1000 2>/dev/null | tee gpurun_out/c8_c1.json | tail -c 900 | tee -a $S
timeout 300 python bench.py --config c3 2>/dev/null | tee gpurun_out/c8_c3.json | tail -c 900 | tee -a $S
MDG_GNN_GRAPH=1 timeout 300 python bench.py --config c3 2>/dev/null | tee gpurun_out/c8_c3_graph.json | tail -c 900 | tee -a $S
MDG_SCHNET_TC=1 MDG_GNN_GRAPH=1 timeout 300 python bench.py --config c3 2>/dev/null | tee gpurun_out/c8_c3_tc.json | tail -c 900 | tee -a $S
timeout 600 python bench.py --config c5 2>gpurun_out/c8_c5.err | tee gpurun_out/c8_c5.json | tail -c 1200 | tee -a $S
MDG_SCHNET_TC=1 timeout 600 python bench.py --config c5 2>gpurun_out/c8_c5_tc.err | tee gpurun_out/c8_c5_tc.json | tail -c 1200 | tee -a $S
ls gpurun_out | tail -3
