#!/bin/bash
# round 2, GPU call 5: k_diag2 + factor-chain CUDA graph: correctness (kernel + lock-step tiers), A/B against the old
# kernel / no graphs / narrower supernode caps, ncu launch list, host-side profile
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_cabi_c.py -m gpu -q -x > gpurun_out/r02e_tests.log 2>&1; echo "gpu tests rc=$?" > gpurun_out/r02e_status.txt
run() { local name=$1; shift
    env "$@" timeout 400 python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/r02e_$name.json 2> gpurun_out/r02e_$name.err
    echo "$name rc=$?" >> gpurun_out/r02e_status.txt; }
run default DRE_AB=1
run diagv1 DRE_DIAG_V=1
run nographs DRE_GRAPHS=0
run snode128 DRE_MAX_SNODE=128
run snode64 DRE_MAX_SNODE=64
run leaf64 DRE_LEAF_SIZE=64
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv"
DRE_GRAPHS=0 timeout 600 $NCU --log-file gpurun_out/r02e_launches_default.csv python tools/profile_step.py 79841 12 > gpurun_out/r02e_ncu_default.log 2>&1
echo "ncu default rc=$?" >> gpurun_out/r02e_status.txt
timeout 300 python tools/host_profile.py 79841 30 > gpurun_out/r02e_host_profile.log 2>&1
echo "host profile rc=$?" >> gpurun_out/r02e_status.txt
cat gpurun_out/r02e_status.txt
tail -5 gpurun_out/r02e_tests.log
for f in gpurun_out/r02e_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    kc = d.get("kernel_classes", {})
    print(sys.argv[1], round(d["value"], 4), round(d["e2e"]["value"], 4), round(d["ms_per_step"], 1), {k: round(v.get("ms_total", 0), 1) for k, v in kc.items()})
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
tail -40 gpurun_out/r02e_host_profile.log
