#!/bin/bash
# First GPU call of the next round (run under gpurun, 1 GPU):  bash tools/round2_first_call.sh
# Everything written after the round-1 GPU budget ran out gets its first hardware run here, each stage under its own
# timeout, all output under gpurun_out/.  Order = value of the information per GPU-minute.
# BEFORE calling gpurun, build the variant libraries in the container (about 5 minutes of nvcc, no GPU needed):
#     python -c "from mdgrad_b200 import build as b; [b.build(variant=v) for v in b.VARIANTS]"
# the .so files travel with the snapshot and the stamps turn section 3's build loop into a no-op on the (charged) GPU box.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== 1. GPU parity suite" | tee gpurun_out/r2_summary.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee -a gpurun_out/r2_summary.txt
echo "== 2. bench, default build" | tee -a gpurun_out/r2_summary.txt
timeout 600 python bench.py --steps 1000 --warmup 200 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
tail -c 1500 gpurun_out/r2_bench_default.json | tee -a gpurun_out/r2_summary.txt
echo "== 3. build variants (device-resident steps/s, force-kernel us)" | tee -a gpurun_out/r2_summary.txt
python -c "
from mdgrad_b200 import build as b
for v in b.VARIANTS:
    try:
        b.build(variant=v)
    except Exception as e:
        print('variant %s does not build: %s' % (v, str(e)[-300:]))" 2>&1 | grep -v "^ptxas info\|bytes stack frame" | tail -6
for v in pf pfmb6 pfmb4 i8 lean i8lean mb6 mb4 u2mb6 u2mb4 fbw2; do   # (sne3 only matters for SchNet: section 5)
  MDG_LIB_VARIANT=$v timeout 200 python bench.py --steps 600 --warmup 60 --no-e2e --no-cpu-baseline \
      > gpurun_out/r2_ab_$v.json 2> gpurun_out/r2_ab_$v.err
  python - "$v" <<'PY' | tee -a gpurun_out/r2_summary.txt
import json, sys
v = sys.argv[1]
try:
    r = json.loads(open("gpurun_out/r2_ab_%s.json" % v).read().strip().splitlines()[-1])
    print("variant %-7s %8.1f steps/s  force %.2f us  finite=%s  K=%s" % (v, r["value"], 1e3 * r["roofline"]["kernel_ms"],
                                                                        r["config"].get("finite"), r["config"].get("rebuild_every")))
except Exception as e:
    print("variant %-7s no result (%s)" % (v, e))
PY
done
echo "== 4. skin scan on the default build" | tee -a gpurun_out/r2_summary.txt
for s in 0.35 0.55 0.65; do
  timeout 200 python bench.py --steps 600 --warmup 60 --skin $s --no-e2e --no-cpu-baseline 2> /dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('skin $s: %.1f steps/s, K=%s' % (r['value'], r['config'].get('rebuild_every')))" | tee -a gpurun_out/r2_summary.txt
done
echo "== 5. SchNet MD (configs[2] / configs[4] shapes)" | tee -a gpurun_out/r2_summary.txt
timeout 300 python tools/schnet_md_bench.py --config water --route engine  2>&1 | tail -1 | tee -a gpurun_out/r2_summary.txt
timeout 300 python tools/schnet_md_bench.py --config water --route oplevel 2>&1 | tail -1 | tee -a gpurun_out/r2_summary.txt
timeout 600 python tools/schnet_md_bench.py --config si --route engine     2>&1 | tail -1 | tee -a gpurun_out/r2_summary.txt
echo "== 5b. SchNet MD: synchronous vs asynchronous vs graph-replay steps" | tee -a gpurun_out/r2_summary.txt
MDG_GNN_SYNC=1 timeout 300 python tools/schnet_md_bench.py --config water --route engine 2>&1 | tail -1 | sed 's/^/sync : /' | tee -a gpurun_out/r2_summary.txt
MDG_GNN_GRAPH=1 timeout 300 python tools/schnet_md_bench.py --config water --route engine 2>&1 | tail -1 | sed 's/^/graph: /' | tee -a gpurun_out/r2_summary.txt
MDG_LIB_VARIANT=sne3 timeout 600 python tools/schnet_md_bench.py --config si --route engine 2>&1 | tail -1 | sed 's/^/sne3 : /' | tee -a gpurun_out/r2_summary.txt
echo "== 6. tcgen05 dense layers: first execution ever, under timeouts" | tee -a gpurun_out/r2_summary.txt
timeout 200 python tools/tc_check.py simt gpurun_out/r2_simt.npz 2>&1 | tail -2 | tee -a gpurun_out/r2_summary.txt
MDG_SCHNET_TC=1 timeout 120 python tools/tc_check.py tc gpurun_out/r2_tc.npz 2>&1 | tail -4 | tee -a gpurun_out/r2_summary.txt
python tools/tc_check.py compare gpurun_out/r2_simt.npz gpurun_out/r2_tc.npz 2>&1 | tail -8 | tee -a gpurun_out/r2_summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks_throttle_reasons.active --format=csv | tee -a gpurun_out/r2_summary.txt
echo "== 7. ncu: launch list + full captures (numbers under ncu are never bench values)" | tee -a gpurun_out/r2_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 300 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_force_rows|k_build_fast|k_step_ba" -s 700 -c 4 \
    -o gpurun_out/r2_prof_lj python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sn_" -c 12 \
    -o gpurun_out/r2_prof_schnet python tools/schnet_md_bench.py --config si --steps 2 > /dev/null 2>&1
ls -la gpurun_out | tee -a gpurun_out/r2_summary.txt
