#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S=gpurun_out/c11_summary.txt
echo "== 1. SchNet tests (tc v2)" | tee $S
MDG_TEST_TC=1 timeout 900 python -m pytest tests/test_schnet.py tests/test_gpu_observables.py -m gpu -q 2>&1 | tail -12 | tee -a $S
echo "== 2. c5 auto(tc v2) / simt / tc v1" | tee -a $S
timeout 600 python bench.py --config c5 2>gpurun_out/c11_c5.err | tee gpurun_out/c11_c5.json | cut -c1-150 | tee -a $S
MDG_SCHNET_TC=0 timeout 600 python bench.py --config c5 2>gpurun_out/c11_c5_simt.err | tee gpurun_out/c11_c5_simt.json | cut -c1-150 | tee -a $S
echo "== 3. c3" | tee -a $S
MDG_GNN_GRAPH=1 timeout 300 python bench.py --config c3 2>/dev/null | tee gpurun_out/c11_c3_graph.json | cut -c1-150 | tee -a $S
echo "== 4. bench ncell 32 single GPU (C4 base) and default" | tee -a $S
timeout 300 python bench.py --steps 600 --warmup 100 --ncell 32 --no-e2e --no-cpu-baseline 2>/dev/null | tee gpurun_out/c11_bench32.json | cut -c1-150 | tee -a $S
timeout 400 python bench.py --steps 1000 --warmup 200 --no-cpu-baseline 2>/dev/null > gpurun_out/c11_bench.json
python -c "
import json
r=json.loads(open('gpurun_out/c11_bench.json').read().strip().splitlines()[-1]); print('value %.1f e2e %.1f ratio %.2f' % (r['value'], r['e2e']['value'], r['e2e']['value']/r['value']))" | tee -a $S
echo "== 5. launch list c5" | tee -a $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/c11_launches_c5.csv \
    python tools/schnet_md_bench.py --config si --steps 2 > /dev/null 2>&1
