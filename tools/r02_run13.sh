#!/bin/bash
# round 2, GPU call 13: 2-GPU pipeline mode (rank 0 ADI chain, rank 1 compress!) -- parity test + bench line
set -u
T=r02n
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/${T}_dist_tests.log 2>&1; echo "dist tests rc=$?" > gpurun_out/${T}_status.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/${T}_bench2.json 2> gpurun_out/${T}_bench2.err
echo "bench2 rc=$?" >> gpurun_out/${T}_status.txt
DRE_DIST_MODE=columns timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/${T}_bench2_columns.json 2> gpurun_out/${T}_bench2_columns.err
echo "bench2 columns rc=$?" >> gpurun_out/${T}_status.txt
cat gpurun_out/${T}_status.txt; tail -3 gpurun_out/${T}_dist_tests.log
tail -c 1500 gpurun_out/${T}_bench2.json; echo; tail -c 600 gpurun_out/${T}_bench2_columns.json; tail -3 gpurun_out/${T}_bench2.err
