#!/bin/bash
# round 2, GPU call 10: k_diag2 v5 (+ streamed global operands, batched panel load)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/r02j_tests.log 2>&1; echo "gpu tests rc=$?" > gpurun_out/r02j_status.txt
run() { local name=$1; shift
    env "$@" timeout 400 python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/r02j_$name.json 2> gpurun_out/r02j_$name.err
    echo "$name rc=$?" >> gpurun_out/r02j_status.txt; }
run default DRE_AB=1
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv"
DRE_GRAPHS=0 timeout 600 $NCU --log-file gpurun_out/r02j_launches_default.csv python tools/profile_step.py 79841 12 > gpurun_out/r02j_ncu_default.log 2>&1
echo "ncu default rc=$?" >> gpurun_out/r02j_status.txt
NCUF="ncu --set full --import-source on --clock-control none --profile-from-start off"
DRE_GRAPHS=0 timeout 600 $NCUF --kernel-name regex:k_diag2 --launch-skip 9 --launch-count 1 -o gpurun_out/r02j_kdiag2_l9 -f python tools/profile_step.py 79841 2 > gpurun_out/r02j_ncu_l9.log 2>&1
echo "ncu l9 rc=$?" >> gpurun_out/r02j_status.txt
cat gpurun_out/r02j_status.txt
tail -5 gpurun_out/r02j_tests.log
for f in gpurun_out/r02j_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    kc = d.get("kernel_classes", {})
    print(sys.argv[1], round(d["value"], 4), round(d["e2e"]["value"], 4), round(d["ms_per_step"], 1), {k: round(v.get("ms_total", 0), 1) for k, v in kc.items()})
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
