"""Aggregate the ncu launch list of bench.py's timed region (tools/ncu_bench.sh) by kernel.
Usage: summarize_bench_launches.py <tag> out.md"""
import csv
import sys
from collections import defaultdict

tag, out = sys.argv[1], sys.argv[2]
rows = list(csv.DictReader(l for l in open(f"gpurun_out/bench_launches_{tag}.csv") if l.startswith('"')))
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v, u = float(r["Metric Value"].replace(",", "")), r["Metric Unit"]
    us = v / 1e3 if u in ("ns", "nsecond") else v if u in ("us", "usecond") else v * 1e3
    k = r["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    agg[k][0] += 1
    agg[k][1] += us
tot = sum(a[1] for a in agg.values())
n = sum(a[0] for a in agg.values())
lines = ["# ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (timed regions only: `--profile-from-start off`",
         "# with B2Y_CUDA_PROFILER_RANGE=1; `--metrics gpu__time_duration.sum --clock-control none`): 2 resident steps + 2 e2e steps.",
         f"# Per-launch times are cold-cache and serialised -- compare shares.  total {tot:.0f} us over {n} launches", "",
         "| kernel | launches | us | share |", "|---|---|---|---|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"| {k} | {a[0]} | {a[1]:.0f} | {100 * a[1] / tot:.1f}% |")
fam = defaultdict(float)
for k, a in agg.items():
    fam[k.split("<")[0]] += a[1]
lines += ["", "| kernel family | us | share |", "|---|---|---|"]
for k, v in sorted(fam.items(), key=lambda kv: -kv[1]):
    lines.append(f"| {k} | {v:.0f} | {100 * v / tot:.1f}% |")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[-8:]))
