#!/bin/bash
# round 2, GPU call 12 (session 5): state of the default path after the container was re-created: bench default,
# device-side timeline of 12 ADI iterations, per-step wall times, config 2/3/5 lines
set -u
T=r02m
mkdir -p gpurun_out
run() { local name=$1; shift
    env "$@" timeout 400 python bench.py --no-cpu --steps 5 --warmup 3 > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err
    echo "$name rc=$?" >> gpurun_out/${T}_status.txt; }
run default DRE_AB=1
DRE_TIMELINE=gpurun_out/${T}_timeline.txt timeout 300 python tools/profile_step.py 79841 22 > gpurun_out/${T}_profile_step.log 2>&1
echo "timeline rc=$?" >> gpurun_out/${T}_status.txt
timeout 300 python tools/step_times.py 79841 6 > gpurun_out/${T}_step_times.log 2>&1
for cfg in 2 3; do
  timeout 600 python bench.py --config $cfg --no-cpu --steps 3 --warmup 3 > gpurun_out/${T}_config$cfg.json 2> gpurun_out/${T}_config$cfg.err
  echo "config$cfg rc=$?" >> gpurun_out/${T}_status.txt
done
timeout 900 python bench.py --config 5 --nside 60 > gpurun_out/${T}_config5_n60.json 2> gpurun_out/${T}_config5_n60.err
echo "config5 n60 rc=$?" >> gpurun_out/${T}_status.txt
cat gpurun_out/${T}_status.txt
tail -12 gpurun_out/${T}_step_times.log
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
