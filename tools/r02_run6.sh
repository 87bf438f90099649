#!/bin/bash
# round 2, GPU call 6: k_diag2 with shared-memory LDL^T steps; prefactor depth A/B; cProfile of the host loop
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/r02f_tests.log 2>&1; echo "gpu tests rc=$?" > gpurun_out/r02f_status.txt
run() { local name=$1; shift
    env "$@" timeout 400 python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/r02f_$name.json 2> gpurun_out/r02f_$name.err
    echo "$name rc=$?" >> gpurun_out/r02f_status.txt; }
run default DRE_AB=1
run depth0 DRE_PREFACTOR_DEPTH=0
run depth1 DRE_PREFACTOR_DEPTH=1
run depth0_lane DRE_PREFACTOR_DEPTH=0 DRE_ASYNC_COMPRESS=1
timeout 300 python -m cProfile -s tottime tools/profile_step.py 79841 40 > gpurun_out/r02f_cprofile.log 2>&1
echo "cprofile rc=$?" >> gpurun_out/r02f_status.txt
DRE_PREFACTOR_DEPTH=0 timeout 300 python -m cProfile -s tottime tools/profile_step.py 79841 40 > gpurun_out/r02f_cprofile_depth0.log 2>&1
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv"
DRE_GRAPHS=0 timeout 600 $NCU --log-file gpurun_out/r02f_launches_default.csv python tools/profile_step.py 79841 12 > gpurun_out/r02f_ncu_default.log 2>&1
echo "ncu default rc=$?" >> gpurun_out/r02f_status.txt
cat gpurun_out/r02f_status.txt
tail -5 gpurun_out/r02f_tests.log
for f in gpurun_out/r02f_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    kc = d.get("kernel_classes", {})
    print(sys.argv[1], round(d["value"], 4), round(d["e2e"]["value"], 4), round(d["ms_per_step"], 1), {k: round(v.get("ms_total", 0), 1) for k, v in kc.items()})
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
head -45 gpurun_out/r02f_cprofile.log
