#!/bin/bash
# round 2, GPU call 19: look-ahead Gram products in several waves of short CTAs + stream priorities (A/B)
set -u
T=r02u
mkdir -p gpurun_out
run() { local name=$1; shift
    env "$@" timeout 400 python bench.py --no-cpu --steps 4 --warmup 3 > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err
    echo "$name rc=$?" >> gpurun_out/${T}_status.txt; }
run default DRE_AB=1
run prio DRE_PRIO=1
run prio_w4 DRE_PRIO=1 DRE_LOOK_WAVES=4
run prio_w8 DRE_PRIO=1 DRE_LOOK_WAVES=8
run prio_w16 DRE_PRIO=1 DRE_LOOK_WAVES=16
run w8 DRE_LOOK_WAVES=8
for v in "DRE_AB=1" "DRE_PRIO=1 DRE_LOOK_WAVES=8"; do
  env $v DRE_RR_STATS=1 timeout 300 python tools/profile_step.py 79841 42 2>&1 | grep "dre rr totals" | tail -1
done
cat gpurun_out/${T}_status.txt
for f in gpurun_out/${T}_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    kc = d.get("kernel_classes", {})
    print(sys.argv[1], round(d["value"], 4), round(d.get("e2e", {}).get("value", 0), 4), round(d.get("ms_per_step", 0), 1), {k[:8]: round(v.get("ms_total", 0), 1) for k, v in kc.items()})
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
