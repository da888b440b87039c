#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S=gpurun_out/c16_summary.txt
echo "== 1. engine GPU tests" | tee $S
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_zfullsize.py tests/test_gpu_api.py -m gpu -q 2>&1 | tail -3 | tee -a $S
echo "== 2. bench (slim builder)" | tee -a $S
timeout 400 python bench.py --steps 1000 --warmup 200 --no-cpu-baseline --no-c4 2>/dev/null > gpurun_out/c16_bench.json
python -c "
import json
r=json.loads(open('gpurun_out/c16_bench.json').read().strip().splitlines()[-1]); print('value %.1f e2e %.1f ratio %.2f force %.2f us' % (r['value'], r['e2e']['value'], r['e2e']['value']/r['value'], 1e3*r['roofline']['kernel_ms']))" | tee -a $S
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-c4 2>/dev/null > gpurun_out/c16_bench20.json
python -c "
import json
r=json.loads(open('gpurun_out/c16_bench20.json').read().strip().splitlines()[-1]); print('20-step: value %.1f e2e %.1f' % (r['value'], r['e2e']['value']))" | tee -a $S
echo "== 3. ncu builder" | tee -a $S
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_build_fast" -s 3 -c 1 \
    -o gpurun_out/c16_prof_build python bench.py --steps 30 --warmup 10 --no-cpu-baseline --no-e2e --no-c4 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 300 --csv --log-file gpurun_out/c16_launches.csv \
    python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-e2e --no-c4 > /dev/null 2>&1
