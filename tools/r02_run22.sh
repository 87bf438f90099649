#!/bin/bash
# round 2, GPU call 22 (run with --gpus 4): scaling lines N = 2 and N = 4 in pipeline mode, 2-GPU parity test
set -u
T=r02x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/${T}_dist_tests.log 2>&1; echo "dist tests rc=$?" > gpurun_out/${T}_status.txt
for N in 2 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/${T}_bench$N.json 2> gpurun_out/${T}_bench$N.err
  echo "bench$N rc=$?" >> gpurun_out/${T}_status.txt
done
cat gpurun_out/${T}_status.txt; tail -3 gpurun_out/${T}_dist_tests.log
grep "bench rank" gpurun_out/${T}_bench*.err
for N in 2 4; do python - $N <<'PY'
import json, sys
d=json.loads(open(f'gpurun_out/r02x_bench{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], d['value'], d['ms_per_step'], d['e2e']['value'], d.get('nccl_exchange'))
PY
done
