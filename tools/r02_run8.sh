#!/bin/bash
# round 2, GPU call 8: k_diag2 v3 (batched smem loops, deferred inverse rows, 4-CTA clusters near the root)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/r02h_tests.log 2>&1; echo "gpu tests rc=$?" > gpurun_out/r02h_status.txt
DRE_DIAG_CLUSTER=0 timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "factor or shift_solve" > gpurun_out/r02h_tests_nocluster.log 2>&1; echo "gpu tests (no cluster) rc=$?" >> gpurun_out/r02h_status.txt
run() { local name=$1; shift
    env "$@" timeout 400 python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/r02h_$name.json 2> gpurun_out/r02h_$name.err
    echo "$name rc=$?" >> gpurun_out/r02h_status.txt; }
run default DRE_AB=1
run nocluster DRE_DIAG_CLUSTER=0
run lane DRE_ASYNC_COMPRESS=1
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv"
DRE_GRAPHS=0 timeout 600 $NCU --log-file gpurun_out/r02h_launches_default.csv python tools/profile_step.py 79841 12 > gpurun_out/r02h_ncu_default.log 2>&1
echo "ncu default rc=$?" >> gpurun_out/r02h_status.txt
DRE_TIMELINE=gpurun_out/r02h_timeline_bench.txt timeout 300 python bench.py --no-cpu --no-clocks --steps 2 --warmup 1 > gpurun_out/r02h_timeline_bench.json 2> gpurun_out/r02h_timeline_bench.err
cat gpurun_out/r02h_status.txt
tail -5 gpurun_out/r02h_tests.log
for f in gpurun_out/r02h_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    kc = d.get("kernel_classes", {})
    print(sys.argv[1], round(d["value"], 4), round(d["e2e"]["value"], 4), round(d["ms_per_step"], 1), {k: round(v.get("ms_total", 0), 1) for k, v in kc.items()})
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
