#!/bin/bash
# round 2, GPU call 18: full gpu test tier, bench lines (default with CPU baseline, reference arm), ncu --set full
# captures of the dominant kernels and the launch list with DRAM traffic
set -u
T=r02t
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -m gpu -q -x --durations=15 > gpurun_out/${T}_gpu_tests.log 2>&1; echo "gpu tests rc=$?" > gpurun_out/${T}_status.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err
echo "bench default rc=$?" >> gpurun_out/${T}_status.txt
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
echo "bench reference rc=$?" >> gpurun_out/${T}_status.txt
for cfg in 2 3; do
  timeout 600 python bench.py --config $cfg --no-cpu --steps 3 --warmup 3 > gpurun_out/${T}_config$cfg.json 2> gpurun_out/${T}_config$cfg.err
  echo "config$cfg rc=$?" >> gpurun_out/${T}_status.txt
done
timeout 900 python bench.py --config 5 --nside 60 > gpurun_out/${T}_config5_n60.json 2> gpurun_out/${T}_config5_n60.err
echo "config5 n60 rc=$?" >> gpurun_out/${T}_status.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_status.txt
NCU="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv"
DRE_GRAPHS=0 timeout 900 $NCU --log-file gpurun_out/${T}_launches_traffic.csv python tools/profile_step.py 79841 12 > gpurun_out/${T}_ncu_traffic.log 2>&1
echo "ncu traffic rc=$?" >> gpurun_out/${T}_status.txt
NCUF="ncu --set full --import-source on --clock-control none --profile-from-start off"
cap() { local name=$1 regex=$2 skip=$3
  DRE_GRAPHS=0 timeout 600 $NCUF --kernel-name regex:$regex --launch-skip $skip --launch-count 1 -o gpurun_out/${T}_full_$name -f python tools/profile_step.py 79841 12 > gpurun_out/${T}_ncu_full_$name.log 2>&1
  echo "ncu full $name rc=$?" >> gpurun_out/${T}_status.txt
  ncu -i gpurun_out/${T}_full_$name.ncu-rep --page raw --csv > gpurun_out/${T}_ncu_full_$name.raw.csv 2>/dev/null
}
cap k_gram2 'k_gram2' 20
cap k_fwd 'k_fwd' 12
cap k_bwd 'k_bwd' 23
cap k_diag2 'k_diag2' 12
cap k_tall_gemm2 'k_tall_gemm2' 10
rm -f gpurun_out/${T}_full_k_bwd.ncu-rep gpurun_out/${T}_full_k_tall_gemm2.ncu-rep
cat gpurun_out/${T}_status.txt
tail -25 gpurun_out/${T}_gpu_tests.log
tail -c 600 gpurun_out/${T}_bench_reference.json
tail -c 1500 gpurun_out/${T}_config5_n60.json; tail -3 gpurun_out/${T}_config5_n60.err
for f in gpurun_out/${T}_bench_default.json gpurun_out/${T}_config2.json gpurun_out/${T}_config3.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    kc = d.get("kernel_classes", {})
    print(sys.argv[1], round(d["value"], 4), round(d.get("e2e", {}).get("value", 0), 4), round(d.get("ms_per_step", 0), 1), {k[:8]: round(v.get("ms_total", 0), 1) for k, v in kc.items()}, d.get("roofline"), d.get("cpu_baseline"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
