#!/bin/bash
# round 2, GPU call 1: full gpu test tier (incl. the baseline-config parity tests), A/B of the compress! threshold and
# the compression lane, ncu launch lists of the default and the row-split sweeps
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s --durations=15 > gpurun_out/r02_tests1.log 2>&1; echo "gpu tests rc=$?" > gpurun_out/r02_status1.txt
run() { local name=$1; shift
    env "$@" DRE_RR_STATS=1 python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/r02a_$name.json 2> gpurun_out/r02a_$name.err
    echo "$name rc=$?" >> gpurun_out/r02_status1.txt; }
run default DRE_AB=1
run olddrop DRE_RR_DROP=3e-15
run lane DRE_ASYNC_COMPRESS=1
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv"
$NCU --log-file gpurun_out/r02_launches_default.csv python tools/profile_step.py 79841 12 > gpurun_out/r02_ncu_default.log 2>&1
echo "ncu default rc=$?" >> gpurun_out/r02_status1.txt
DRE_SWEEP2=1 $NCU --log-file gpurun_out/r02_launches_sweep2.csv python tools/profile_step.py 79841 12 > gpurun_out/r02_ncu_sweep2.log 2>&1
echo "ncu sweep2 rc=$?" >> gpurun_out/r02_status1.txt
cat gpurun_out/r02_status1.txt
tail -5 gpurun_out/r02_tests1.log
