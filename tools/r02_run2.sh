#!/bin/bash
# round 2, GPU call 2: whole gpu tier (in-tree eigensolver, no cuSOLVER), bench default / lane
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s --durations=15 > gpurun_out/r02_tests2.log 2>&1; echo "gpu tests rc=$?" > gpurun_out/r02_status2.txt
run() { local name=$1; shift
    env "$@" DRE_RR_STATS=1 python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/r02b_$name.json 2> gpurun_out/r02b_$name.err
    echo "$name rc=$?" >> gpurun_out/r02_status2.txt; }
run default DRE_AB=1
run lane DRE_ASYNC_COMPRESS=1
DRE_TRACE=1 python tools/profile_step.py 79841 12 2>&1 | grep -E "eig_sym|PROFILE" | tail -20 > gpurun_out/r02_eigtrace.log
cat gpurun_out/r02_status2.txt
tail -8 gpurun_out/r02_tests2.log
