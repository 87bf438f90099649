"""Summarise an ncu launch list (csv with gpu__time_duration.sum and, optionally, dram__bytes_read.sum /
dram__bytes_write.sum per launch) of tools/profile_step.py into
  profiles/<tag>_summary.md      per-kernel table: launches, total/avg time, share, DRAM bytes
  profiles/ncu_traffic.json      DRAM traffic per launch GROUP of the kernel classes bench.py reports
                                 (one block solve = all k_fwd + k_bwd launches of one shift, one SpMM, ...)
usage: python tools/ncu_summarize.py gpurun_out/launchesN.csv <tag> <adi iterations in the profiled region>"""
import collections
import csv
import json
import os
import sys

src, tag, iters = sys.argv[1], sys.argv[2], int(sys.argv[3])
rows = list(csv.reader(open(src)))
hdr = None
per = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
seen = set()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        ki, mi, vi, ii, ui = r.index("Kernel Name"), r.index("Metric Name"), r.index("Metric Value"), r.index("ID"), r.index("Metric Unit")
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    name = r[ki].split("(")[0].replace("void ", "").replace("dre::", "")
    unit = r[ui]
    scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    per[name][r[mi]] += v * scale
    if (r[ii], name) not in seen:
        seen.add((r[ii], name))
        cnt[name] += 1
tot = sum(m.get("gpu__time_duration.sum", 0.0) for m in per.values())
lines = [f"# ncu launch list summary `{tag}` ({os.path.basename(src)}; {iters} ADI iterations of the n=79841 Ros1 step, "
         "cold-cache serialised replays: compare SHARES, not absolutes)", "",
         "| kernel | launches | total ms | avg us | share | DRAM read MB | DRAM write MB |", "|---|---|---|---|---|---|---|"]
for name, m in sorted(per.items(), key=lambda kv: -kv[1].get("gpu__time_duration.sum", 0.0)):
    t = m.get("gpu__time_duration.sum", 0.0)
    if t < 1e-3 * tot:
        continue
    lines.append(f"| `{name}` | {cnt[name]} | {t / 1e6:.3f} | {t / cnt[name] / 1e3:.1f} | {t / tot:.3f} | "
                 f"{m.get('dram__bytes_read.sum', 0) / 1e6:.1f} | {m.get('dram__bytes_write.sum', 0) / 1e6:.1f} |")
lines.append("")
lines.append(f"total kernel time {tot / 1e6:.2f} ms = {tot / 1e6 / iters:.2f} ms per ADI iteration")
os.makedirs("profiles", exist_ok=True)
open(f"profiles/{tag}_summary.md", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))


def group(prefixes, ngroups):
    rd = sum(m.get("dram__bytes_read.sum", 0.0) for n, m in per.items() if n.startswith(prefixes))
    wr = sum(m.get("dram__bytes_write.sum", 0.0) for n, m in per.items() if n.startswith(prefixes))
    t = sum(m.get("gpu__time_duration.sum", 0.0) for n, m in per.items() if n.startswith(prefixes))
    return {"dram_bytes_per_launch_group": (rd + wr) / ngroups if ngroups else None, "launch_groups": ngroups,
            "ncu_ms_per_launch_group": t / 1e6 / ngroups if ngroups else None}


if any("dram__bytes_read.sum" in m for m in per.values()):
    nspmm = cnt.get("k_spmm", 0)
    out = {"_source": f"{os.path.basename(src)} (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, n=79841, "
                      f"{iters} ADI iterations; per launch group)",
           "sptrsm_fwd_bwd_sweeps": group(("k_fwd", "k_bwd"), iters),
           "supernodal_ldlt_factor": group(("k_diag", "k_l21", "k_schur", "k_extend_add", "k_assemble"), iters),
           "csr_spmm": group(("k_spmm",), nspmm),
           "gram_dmma": group(("k_gram",), sum(v for k, v in cnt.items() if k.startswith("k_gram"))),
           "tall_gemm_dmma": group(("k_tall_gemm",), sum(v for k, v in cnt.items() if k.startswith("k_tall_gemm")))}
    json.dump(out, open("profiles/ncu_traffic.json", "w"), indent=1)
    print(json.dumps(out, indent=1))
