#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S=gpurun_out/c10_summary.txt
echo "== 1. full GPU parity suite" | tee $S
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee -a $S
echo "== 2. bench default (full line)" | tee -a $S
BENCH_PROFILE_E2E=1 timeout 400 python bench.py --steps 1000 --warmup 200 > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err
python -c "
import json
r=json.loads(open('gpurun_out/c10_bench.json').read().strip().splitlines()[-1]); print('value %.1f e2e %.1f force %.2f us frac %.3f cpu %s' % (r['value'], r['e2e']['value'], 1e3*r['roofline']['kernel_ms'], r['roofline']['frac'], r['cpu_baseline']))" | tee -a $S
echo "== 3. configs" | tee -a $S
timeout 300 python bench.py --config c1 --steps 1000 2>/dev/null | tee gpurun_out/c10_c1.json | cut -c1-200 | tee -a $S
timeout 300 python bench.py --config c3 2>/dev/null | tee gpurun_out/c10_c3.json | cut -c1-200 | tee -a $S
MDG_GNN_GRAPH=1 timeout 300 python bench.py --config c3 2>/dev/null | tee gpurun_out/c10_c3_graph.json | cut -c1-200 | tee -a $S
timeout 600 python bench.py --config c5 2>gpurun_out/c10_c5.err | tee gpurun_out/c10_c5.json | cut -c1-200 | tee -a $S
MDG_SCHNET_TC=0 timeout 600 python bench.py --config c5 2>gpurun_out/c10_c5_simt.err | tee gpurun_out/c10_c5_simt.json | cut -c1-200 | tee -a $S
echo "== 4. launch lists" | tee -a $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/c10_launches_c3.csv \
    python tools/schnet_md_bench.py --config water --steps 10 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/c10_launches_c5tc.csv \
    python tools/schnet_md_bench.py --config si --steps 2 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/c10_launches_c1.csv \
    python bench.py --config c1 --steps 150 > /dev/null 2>&1
ls gpurun_out | wc -l
