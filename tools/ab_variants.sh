#!/bin/bash
# A/B of the experimental build variants (mdgrad_b200/build.py VARIANTS) on one B200:
# bench every variant (device-resident steps/s, no e2e / CPU legs), then run the whole GPU suite on the fastest.
# usage (under gpurun): bash tools/ab_variants.sh x1 x2 x3 x4
mkdir -p gpurun_out
for v in "$@"; do
  MDG_LIB_VARIANT=$v timeout 150 python bench.py --steps 600 --warmup 30 --no-e2e --no-cpu-baseline \
      > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err || echo "variant $v: bench failed rc=$?"
done
best=$(python - "$@" <<'PY'
import json, sys
best, bv = "", 0.0
for v in sys.argv[1:]:
    try:
        r = json.loads(open("gpurun_out/ab_%s.json" % v).read().strip().splitlines()[-1])
        ok = r["config"].get("finite", False)
        print("variant %s: %.1f steps/s, force %.2f us, finite=%s" % (v, r["value"], 1e3 * r["roofline"]["kernel_ms"], ok), file=sys.stderr)
        if ok and r["value"] > bv:
            best, bv = v, r["value"]
    except Exception as e:
        print("variant %s: no result (%s)" % (v, e), file=sys.stderr)
print(best)
PY
)
echo "best variant: '$best'"
if [ -n "$best" ]; then
  MDG_LIB_VARIANT=$best timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/ab_pytest_$best.log
  cat gpurun_out/ab_pytest_$best.log
fi
