#!/bin/bash
# round 2, GPU call 3: A/B of the compress! look-ahead, with the in-tree eigensolver
set -u
mkdir -p gpurun_out
run() { local name=$1; shift
    env "$@" DRE_RR_STATS=1 python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/r02c_$name.json 2> gpurun_out/r02c_$name.err
    echo "$name rc=$?" >> gpurun_out/r02_status3.txt; }
rm -f gpurun_out/r02_status3.txt
run default DRE_AB=1
run nolook DRE_RR_LOOKAHEAD=0
run lane DRE_ASYNC_COMPRESS=1
cat gpurun_out/r02_status3.txt
