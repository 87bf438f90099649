#!/bin/bash
# round 2, GPU call 20: new defaults (priorities + 4 look-ahead waves); compression lane on the same GPU once more,
# now with priorities and short-lived Gram CTAs
set -u
T=r02v
mkdir -p gpurun_out
run() { local name=$1; shift
    env "$@" timeout 400 python bench.py --no-cpu --steps 4 --warmup 3 > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err
    echo "$name rc=$?" >> gpurun_out/${T}_status.txt; }
run default DRE_AB=1
run lane DRE_ASYNC_COMPRESS=1
run lane_gw4 DRE_ASYNC_COMPRESS=1 DRE_GRAM_WAVES=4
run lane_gw4_norm DRE_ASYNC_COMPRESS=1 DRE_GRAM_WAVES=4 DRE_ASYNC_NORM=1
run norm DRE_ASYNC_NORM=1
run depth2 DRE_PREFACTOR_DEPTH=2
cat gpurun_out/${T}_status.txt
for f in gpurun_out/${T}_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    kc = d.get("kernel_classes", {})
    print(sys.argv[1], round(d["value"], 4), round(d.get("e2e", {}).get("value", 0), 4), round(d.get("ms_per_step", 0), 1), {k[:8]: round(v.get("ms_total", 0), 1) for k, v in kc.items()}, d.get("compression_lane"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
