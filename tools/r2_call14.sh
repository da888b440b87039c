#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S=gpurun_out/c14_summary.txt
echo "== 1. full GPU suite" | tee $S
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee -a $S
echo "== 2. bench default with e2e profile" | tee -a $S
BENCH_PROFILE_E2E=1 timeout 400 python bench.py --steps 1000 --warmup 200 --no-cpu-baseline --no-c4 2>gpurun_out/c14_bench.err > gpurun_out/c14_bench.json
python -c "
import json
r=json.loads(open('gpurun_out/c14_bench.json').read().strip().splitlines()[-1]); print('value %.1f e2e %.1f ratio %.2f' % (r['value'], r['e2e']['value'], r['e2e']['value']/r['value']))" | tee -a $S
echo "== 3. c5 / c3" | tee -a $S
timeout 600 python bench.py --config c5 2>gpurun_out/c14_c5.err | tee gpurun_out/c14_c5.json | cut -c1-150 | tee -a $S
MDG_GNN_GRAPH=1 timeout 300 python bench.py --config c3 2>/dev/null | tee gpurun_out/c14_c3_graph.json | cut -c1-150 | tee -a $S
timeout 300 python bench.py --config c3 2>/dev/null | tee gpurun_out/c14_c3.json | cut -c1-150 | tee -a $S
