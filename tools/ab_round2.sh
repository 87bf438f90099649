#!/bin/bash
# A/B of the opt-in code paths in ONE gpurun call:
#   gpurun --timeout 1500 -- bash tools/ab_round2.sh
# Writes gpurun_out/ab_*.json (one bench line each) and gpurun_out/ab_summary.md.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/ab_tests_default.log 2>&1; echo "default gpu tests rc=$?" > gpurun_out/ab_status.txt
DRE_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -m gpu -q > gpurun_out/ab_tests_experimental.log 2>&1
echo "experimental gpu tests rc=$?" >> gpurun_out/ab_status.txt
run() {   # name, env assignments...
    local name=$1; shift
    env "$@" python bench.py --no-cpu --steps 2 --warmup 3 > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
    echo "$name rc=$?" >> gpurun_out/ab_status.txt
}
run default DRE_AB=default
run sweep2 DRE_SWEEP2=1
run narrow DRE_DIAG_NARROW_MIN=296
run spmm2 DRE_SPMM2=1
run eager DRE_RR_EAGER=1 DRE_RR_STATS=1
run async_norm DRE_ASYNC_NORM=1
run async_compress DRE_ASYNC_COMPRESS=1
run all DRE_SWEEP2=1 DRE_DIAG_NARROW_MIN=296 DRE_SPMM2=1
run all_async DRE_SWEEP2=1 DRE_DIAG_NARROW_MIN=296 DRE_SPMM2=1 DRE_ASYNC_NORM=1
run all_async2 DRE_SWEEP2=1 DRE_DIAG_NARROW_MIN=296 DRE_SPMM2=1 DRE_ASYNC_NORM=1 DRE_ASYNC_COMPRESS=1
run all_leaf64 DRE_SWEEP2=1 DRE_DIAG_NARROW_MIN=296 DRE_SPMM2=1 DRE_LEAF_SIZE=64
# nnz(L) 8.41M (leaf 96, 12 levels) -> 5.58M (leaf 48, 13 levels): the leaf level is FP64-bound, levels are cheap with sweep v2
run all_leaf48 DRE_SWEEP2=1 DRE_DIAG_NARROW_MIN=296 DRE_SPMM2=1 DRE_LEAF_SIZE=48
python - <<'PY'
import glob, json, os
rows = []
for f in sorted(glob.glob("gpurun_out/ab_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        rows.append((os.path.basename(f), "unreadable: %s" % e))
        continue
    kc = d.get("kernel_classes", {})
    g = lambda k: kc.get(k, {}).get("ms_avg", float("nan"))
    rows.append((os.path.basename(f)[3:-5], d["value"], d["e2e"]["value"], d["ms_per_step"],
                 g("sptrsm_fwd_bwd_sweeps"), g("supernodal_ldlt_factor"), g("csr_spmm"),
                 kc.get("gram_dmma", {}).get("ms_total", float("nan")),
                 kc.get("tall_gemm_dmma", {}).get("ms_total", float("nan")), d["config"].get("opt_in")))
with open("gpurun_out/ab_summary.md", "w") as out:
    out.write("| run | steps/s | e2e | ms/step | sweeps ms/shift | factor ms/shift | spmm ms | gram ms/pass | tall ms/pass | opt_in |\n")
    out.write("|---|---|---|---|---|---|---|---|---|---|\n")
    for r in rows:
        out.write("| " + " | ".join(("%.4g" % x) if isinstance(x, float) else str(x) for x in r) + " |\n")
print(open("gpurun_out/ab_summary.md").read())
PY
