#!/bin/bash
# round 2, GPU call 21: 2-GPU pipeline mode with receiver thread + fixed-size receive buffers
set -u
T=r02w
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/${T}_bench2.json 2> gpurun_out/${T}_bench2.err
echo "bench2 rc=$?" > gpurun_out/${T}_status.txt
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/${T}_dist_tests.log 2>&1; echo "dist tests rc=$?" >> gpurun_out/${T}_status.txt
cat gpurun_out/${T}_status.txt; tail -3 gpurun_out/${T}_dist_tests.log
grep "bench rank" gpurun_out/${T}_bench2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02w_bench2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('nccl_exchange'))
PY
