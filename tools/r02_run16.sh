#!/bin/bash
# round 2, GPU call 16: steady-state phase breakdown of compress! (stats window = the profiled iterations), launch list
set -u
T=r02q
mkdir -p gpurun_out
DRE_RR_STATS=1 timeout 300 python tools/profile_step.py 79841 42 > gpurun_out/${T}_rrstats.log 2>&1
DRE_RR_LEGACY=1 DRE_RR_STATS=1 timeout 300 python tools/profile_step.py 79841 42 > gpurun_out/${T}_rrstats_legacy.log 2>&1
DRE_RR_LOOKAHEAD=0 DRE_RR_STATS=1 timeout 300 python tools/profile_step.py 79841 42 > gpurun_out/${T}_rrstats_nolook.log 2>&1
DRE_RR_STATS=1 DRE_EIG_DEBUG=1 timeout 300 python tools/profile_step.py 79841 42 > gpurun_out/${T}_rrstats_eig.log 2>&1
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv"
DRE_GRAPHS=0 timeout 600 $NCU --log-file gpurun_out/${T}_launches_default.csv python tools/profile_step.py 79841 12 > gpurun_out/${T}_ncu_default.log 2>&1
echo "ncu default rc=$?"
grep "dre rr totals\|PROFILE_REGION_END" gpurun_out/${T}_rrstats*.log
grep "dre eig" gpurun_out/${T}_rrstats_eig.log | tail -5
