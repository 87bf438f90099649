#!/bin/bash
# round 2, GPU call 23 (run with --gpus 4): two compression lanes (DRE_PIPE_LANES=2): 3-GPU lock-step test, bench at
# N = 4 with one and two lanes
set -u
T=r02y
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/${T}_dist_tests.log 2>&1; echo "dist tests rc=$?" > gpurun_out/${T}_status.txt
for L in 2 1; do
  DRE_PIPE_LANES=$L timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2952$L bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu > gpurun_out/${T}_bench4_lanes$L.json 2> gpurun_out/${T}_bench4_lanes$L.err
  echo "bench4 lanes$L rc=$?" >> gpurun_out/${T}_status.txt
done
cat gpurun_out/${T}_status.txt; tail -5 gpurun_out/${T}_dist_tests.log
grep "bench rank" gpurun_out/${T}_bench4*.err
grep -v Warning gpurun_out/${T}_bench4_lanes2.err | tail -5
for L in 2 1; do python - $L <<'PY'
import json, sys
try:
    d=json.loads(open(f'gpurun_out/r02y_bench4_lanes{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print("lanes", sys.argv[1], d['value'], d['ms_per_step'], d['e2e']['value'], d.get('nccl_exchange'), d['config']['rank_X_and_residual'][-1])
except Exception as e:
    print("lanes", sys.argv[1], "unreadable", e)
PY
done
