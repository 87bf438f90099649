#!/bin/bash
# round 2, GPU call 15: compress! changes (DMMA core product, whole-chunk projection): parity subset + A/B bench
set -u
T=r02p
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/${T}_tests.log 2>&1; echo "gpu tests rc=$?" > gpurun_out/${T}_status.txt
run() { local name=$1; shift
    env "$@" timeout 400 python bench.py --no-cpu --steps 5 --warmup 3 > gpurun_out/${T}_$name.json 2> gpurun_out/${T}_$name.err
    echo "$name rc=$?" >> gpurun_out/${T}_status.txt; }
run default DRE_AB=1
run legacy DRE_RR_LEGACY=1
DRE_RR_STATS=1 DRE_EIG_DEBUG=1 timeout 300 python tools/profile_step.py 79841 22 > gpurun_out/${T}_rrstats.log 2>&1
DRE_RR_LEGACY=1 DRE_RR_STATS=1 timeout 300 python tools/profile_step.py 79841 22 > gpurun_out/${T}_rrstats_legacy.log 2>&1
timeout 300 python tools/step_times.py 79841 5 > gpurun_out/${T}_step_times.log 2>&1
cat gpurun_out/${T}_status.txt
tail -4 gpurun_out/${T}_tests.log
grep "dre rr totals\|PROFILE_REGION_END" gpurun_out/${T}_rrstats.log gpurun_out/${T}_rrstats_legacy.log
tail -4 gpurun_out/${T}_step_times.log
for f in gpurun_out/${T}_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    kc = d.get("kernel_classes", {})
    print(sys.argv[1], round(d["value"], 4), round(d.get("e2e", {}).get("value", 0), 4), round(d.get("ms_per_step", 0), 1), {k[:8]: round(v.get("ms_total", 0), 1) for k, v in kc.items()}, d["config"]["rank_X_and_residual"][-1], d["e2e_vs_resident_K_relerr"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
