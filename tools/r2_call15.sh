#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S=gpurun_out/c15_summary.txt
echo "== 1. full GPU suite + smoke" | tee $S
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee -a $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a $S
echo "== 2. bench default full line (driver defaults)" | tee -a $S
timeout 600 python bench.py 2>gpurun_out/c15_bench.err > gpurun_out/c15_bench.json
python -c "
import json
r=json.loads(open('gpurun_out/c15_bench.json').read().strip().splitlines()[-1]); print('value %.1f e2e %.1f ratio %.2f frac %.3f traffic %s c4 %.1f cpu %.4g' % (r['value'], r['e2e']['value'], r['e2e']['value']/r['value'], r['roofline']['frac'], r['roofline']['traffic'], r['c4']['box_steps_per_s'], r['cpu_baseline']['value']))" | tee -a $S
timeout 600 python bench.py --steps 20 --warmup 5 2>/dev/null > gpurun_out/c15_bench_20.json
python -c "
import json
r=json.loads(open('gpurun_out/c15_bench_20.json').read().strip().splitlines()[-1]); print('driver-style 20 steps: value %.1f e2e %.1f' % (r['value'], r['e2e']['value']))" | tee -a $S
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tee gpurun_out/c15_ref.json | cut -c1-300 | tee -a $S
echo "== 3. configs" | tee -a $S
for c in c1 c3 c5; do timeout 600 python bench.py --config $c 2>gpurun_out/c15_$c.err | tee gpurun_out/c15_$c.json | cut -c1-120 | tee -a $S; done
MDG_SCHNET_TC=0 timeout 600 python bench.py --config c5 2>/dev/null | tee gpurun_out/c15_c5_simt.json | cut -c1-120 | tee -a $S
MDG_GNN_GRAPH=0 timeout 600 python bench.py --config c3 2>/dev/null | tee gpurun_out/c15_c3_nograph.json | cut -c1-120 | tee -a $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/c15_launches_c5.csv \
    python tools/schnet_md_bench.py --config si --steps 2 > /dev/null 2>&1
