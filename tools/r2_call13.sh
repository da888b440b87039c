#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S=gpurun_out/c13_summary.txt
echo "== 1. SchNet tests (tc v2b)" | tee $S
timeout 900 python -m pytest tests/test_schnet.py -m gpu -q 2>&1 | tail -4 | tee -a $S
echo "== 2. c5 auto(tc) / simt" | tee -a $S
timeout 600 python bench.py --config c5 2>gpurun_out/c13_c5.err | tee gpurun_out/c13_c5.json | cut -c1-150 | tee -a $S
MDG_SCHNET_TC=0 timeout 600 python bench.py --config c5 2>gpurun_out/c13_c5_simt.err | tee gpurun_out/c13_c5_simt.json | cut -c1-150 | tee -a $S
echo "== 3. bench default (e2e)" | tee -a $S
timeout 400 python bench.py --steps 1000 --warmup 200 --no-cpu-baseline 2>gpurun_out/c13_bench.err > gpurun_out/c13_bench.json
python -c "
import json
r=json.loads(open('gpurun_out/c13_bench.json').read().strip().splitlines()[-1]); print('value %.1f e2e %.1f ratio %.2f c4 %s' % (r['value'], r['e2e']['value'], r['e2e']['value']/r['value'], r.get('c4')))" | tee -a $S
echo "== 4. launch list + ncu c5" | tee -a $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/c13_launches_c5.csv \
    python tools/schnet_md_bench.py --config si --steps 2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_sn_gemm_tc2" -s 8 -c 3 \
    -o gpurun_out/c13_prof_tc python tools/schnet_md_bench.py --config si --steps 1 > /dev/null 2>&1
