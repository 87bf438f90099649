#!/bin/bash
# round 2, GPU call 7: device timeline of 40 ADI iterations; ncu --set full of k_diag2 (leaf level and the level-9
# launch with two fat supernodes), source page
set -u
mkdir -p gpurun_out
DRE_TIMELINE=gpurun_out/r02g_timeline.txt timeout 300 python tools/profile_step.py 79841 40 > gpurun_out/r02g_timeline.log 2>&1
echo "timeline rc=$?" > gpurun_out/r02g_status.txt
export DRE_GRAPHS=0
NCU="ncu --set full --import-source on --clock-control none --profile-from-start off"
timeout 600 $NCU --kernel-name regex:k_diag2 --launch-skip 9 --launch-count 1 -o gpurun_out/r02g_kdiag2_l9 -f python tools/profile_step.py 79841 2 > gpurun_out/r02g_ncu_l9.log 2>&1
echo "ncu l9 rc=$?" >> gpurun_out/r02g_status.txt
timeout 600 $NCU --kernel-name regex:k_diag2 --launch-skip 0 --launch-count 1 -o gpurun_out/r02g_kdiag2_l0 -f python tools/profile_step.py 79841 2 > gpurun_out/r02g_ncu_l0.log 2>&1
echo "ncu l0 rc=$?" >> gpurun_out/r02g_status.txt
cat gpurun_out/r02g_status.txt; ls -la gpurun_out/*.ncu-rep; wc -l gpurun_out/r02g_timeline.txt
