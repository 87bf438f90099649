#!/bin/bash
# round 2, GPU call 11: stream priorities (DRE_PRIO) with the compression lane, Gram wave count; bench lines of configs 2/3/5
set -u
T=r02k
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/${T}_tests.log 2>&1; echo "gpu tests rc=$?" > gpurun_out/${T}_status.txt
run() { local name=$1; shift
    env "$@" timeout 400 python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err
    echo "$name rc=$?" >> gpurun_out/${T}_status.txt; }
run default DRE_AB=1
run prio DRE_PRIO=1
run lane DRE_ASYNC_COMPRESS=1
run lane_prio DRE_ASYNC_COMPRESS=1 DRE_PRIO=1
run lane_prio_w3 DRE_ASYNC_COMPRESS=1 DRE_PRIO=1 DRE_GRAM_WAVES=3
run lane_prio_d1 DRE_ASYNC_COMPRESS=1 DRE_PRIO=1 DRE_PREFACTOR_DEPTH=1
DRE_ASYNC_COMPRESS=1 DRE_PRIO=1 DRE_TIMELINE=gpurun_out/${T}_timeline_lane_prio.txt timeout 300 python bench.py --no-cpu --no-clocks --steps 2 --warmup 1 > gpurun_out/${T}_timeline_lane_prio.json 2> gpurun_out/${T}_timeline_lane_prio.err
for cfg in 2 3; do
  timeout 600 python bench.py --config $cfg --no-cpu --steps 3 --warmup 3 > gpurun_out/${T}_config$cfg.json 2> gpurun_out/${T}_config$cfg.err
  echo "config$cfg rc=$?" >> gpurun_out/${T}_status.txt
done
timeout 900 python bench.py --config 5 --nside 60 > gpurun_out/${T}_config5_n60.json 2> gpurun_out/${T}_config5_n60.err
echo "config5 n60 rc=$?" >> gpurun_out/${T}_status.txt
cat gpurun_out/${T}_status.txt
tail -3 gpurun_out/${T}_tests.log
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
