#!/bin/bash
# round 2, GPU call 4 (session 3): whole gpu tier, A/B of every opt-in path with the in-tree eigensolver and the compress!
# look-ahead, ncu launch list of the default build
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r02d_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s --durations=25 > gpurun_out/r02d_tests.log 2>&1; echo "gpu tests rc=$?" > gpurun_out/r02d_status.txt
run() { local name=$1; shift
    env "$@" DRE_RR_STATS=1 timeout 400 python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/r02d_$name.json 2> gpurun_out/r02d_$name.err
    echo "$name rc=$?" >> gpurun_out/r02d_status.txt; }
run default DRE_AB=1
run sweep2 DRE_SWEEP2=1
run asyncnorm DRE_ASYNC_NORM=1
run lane DRE_ASYNC_COMPRESS=1
run all DRE_SWEEP2=1 DRE_DIAG_NARROW_MIN=296 DRE_SPMM2=1 DRE_ASYNC_NORM=1
run all_lane DRE_SWEEP2=1 DRE_DIAG_NARROW_MIN=296 DRE_SPMM2=1 DRE_ASYNC_NORM=1 DRE_ASYNC_COMPRESS=1
run all_leaf48 DRE_SWEEP2=1 DRE_DIAG_NARROW_MIN=296 DRE_SPMM2=1 DRE_ASYNC_NORM=1 DRE_LEAF_SIZE=48
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv"
timeout 600 $NCU --log-file gpurun_out/r02d_launches_default.csv python tools/profile_step.py 79841 12 > gpurun_out/r02d_ncu_default.log 2>&1
echo "ncu default rc=$?" >> gpurun_out/r02d_status.txt
cat gpurun_out/r02d_status.txt
tail -8 gpurun_out/r02d_tests.log
for f in gpurun_out/r02d_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    kc = d.get("kernel_classes", {})
    print(sys.argv[1], d["value"], d["e2e"]["value"], d["ms_per_step"], {k: round(v.get("ms_total", 0), 1) for k, v in kc.items()})
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
