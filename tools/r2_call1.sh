#!/bin/bash
# Round-2 GPU call 1 (1 GPU): first hardware run of everything written after the round-1 budget ran out, the default
# bench line, SchNet MD timings, the first execution of the tcgen05 dense layers, and ncu captures of the CURRENT
# default kernels.  Every stage under its own timeout; all output under gpurun_out/.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
S=gpurun_out/c1_summary.txt
echo "== 1. GPU parity suite" | tee $S
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 | tee -a $S
echo "== 2. bench, default build" | tee -a $S
timeout 400 python bench.py --steps 1000 --warmup 200 > gpurun_out/c1_bench_default.json 2> gpurun_out/c1_bench_default.err
tail -c 3000 gpurun_out/c1_bench_default.json | tee -a $S
echo "== 3. SchNet MD" | tee -a $S
timeout 300 python tools/schnet_md_bench.py --config water --route engine  2>&1 | tail -1 | tee -a $S
timeout 300 python tools/schnet_md_bench.py --config water --route oplevel 2>&1 | tail -1 | tee -a $S
MDG_GNN_GRAPH=1 timeout 300 python tools/schnet_md_bench.py --config water --route engine 2>&1 | tail -1 | sed 's/^/graph: /' | tee -a $S
timeout 600 python tools/schnet_md_bench.py --config si --route engine     2>&1 | tail -1 | tee -a $S
echo "== 4. tcgen05 dense layers: first execution ever, under timeouts" | tee -a $S
timeout 200 python tools/tc_check.py simt gpurun_out/c1_simt.npz 2>&1 | tail -2 | tee -a $S
MDG_SCHNET_TC=1 timeout 120 python tools/tc_check.py tc gpurun_out/c1_tc.npz 2>&1 | tail -4 | tee -a $S
python tools/tc_check.py compare gpurun_out/c1_simt.npz gpurun_out/c1_tc.npz 2>&1 | tail -8 | tee -a $S
nvidia-smi --query-gpu=name,clocks.sm,clocks_throttle_reasons.active --format=csv | tee -a $S
echo "== 5. ncu: launch list + full captures (numbers under ncu are never bench values)" | tee -a $S
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 300 --csv --log-file gpurun_out/c1_launches.csv \
    python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_force_rows|k_build_fast|k_step_ba" -s 700 -c 4 \
    -o gpurun_out/c1_prof_lj python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/c1_launches_si.csv \
    python tools/schnet_md_bench.py --config si --steps 2 > /dev/null 2>&1
ls -la gpurun_out | tee -a $S
