"""Join an ncu launch list (tools/ncu_capture.sh) with the op table of tools/profile_step.py into a per-launch summary:
device time, DRAM traffic vs algorithmic bytes, tensor-pipe activity.  Usage: summarize_launches.py <tag> [out.md]"""
import csv
import json
import sys
from collections import defaultdict

tag = sys.argv[1]
rows = list(csv.DictReader(l for l in open(f"gpurun_out/launches_{tag}.csv") if l.startswith('"')))
tab = json.load(open(f"gpurun_out/optable_{tag}.json"))
per = defaultdict(dict)
for r in rows:
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        v = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3 if u in ("ms", "msecond") else v)
    if r["Metric Name"].startswith("dram__bytes"):
        v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    per[int(r["ID"])][r["Metric Name"]] = v
    per[int(r["ID"])]["kernel"] = r["Kernel Name"].split("(")[0].replace("<unnamed>::", "")
ops = tab["ops"]
if len(per) != len(ops):  # torch-side launches inside the pass (e.g. output zero-fills) are listed by kernel name only
    names = [per[i]["kernel"] for i in range(len(per))]
    fixed, j = [], 0
    for i in range(len(per)):
        if names[i].startswith("void at::") or names[i].startswith("at::"):
            fixed.append({"name": "(torch) " + names[i][:40], "kind": "torch"})
        else:
            fixed.append(ops[j]); j += 1
    ops = fixed
assert len(per) == len(ops), (len(per), len(ops))
tot = sum(p["gpu__time_duration.sum"] for p in per.values())
lines = [f"# ncu launch list `{tag}`: {tab['workload']} batch {tab['batch']}, one eager pass, {len(ops)} launches, "
         f"sum {tot:.0f} us (cold-cache, serialised: compare shares)", "",
         "| # | layer | kernel | us | share | alg MB | dram MB (rd+wr) | traffic/alg | GB/s (alg) | tensor % |", "|---|---|---|---|---|---|---|---|---|---|"]
agg = defaultdict(lambda: [0.0, 0.0, 0.0, 0])
for i, o in enumerate(ops):
    p = per[i]
    t = p["gpu__time_duration.sum"]
    dr = p.get("dram__bytes_read.sum", 0) + p.get("dram__bytes_write.sum", 0)
    alg = o.get("bytes", 0)
    tp = p.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0)
    lines.append(f"| {i} | {o['name']} | {p['kernel'][:24]} | {t:.1f} | {100*t/tot:.1f}% | {alg/1e6:.1f} | {dr/1e6:.1f} | "
                 f"{(dr/alg if alg else 0):.2f} | {(alg/t/1e3 if alg else 0):.0f} | {tp:.1f} |")
    a = agg[p["kernel"]]
    a[0] += t; a[1] += alg; a[2] += dr; a[3] += 1
lines += ["", "| kernel | launches | us | share | alg MB | dram MB | alg GB/s |", "|---|---|---|---|---|---|---|"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    lines.append(f"| {k} | {a[3]} | {a[0]:.0f} | {100*a[0]/tot:.1f}% | {a[1]/1e6:.0f} | {a[2]/1e6:.0f} | {(a[1]/a[0]/1e3 if a[0] else 0):.0f} |")
conv = [(k, a) for k, a in agg.items() if "conv_tc_kernel" in k]
traffic = {"workload": tab["workload"], "batch": tab["batch"], "source": f"ncu launch list {tag} (tools/ncu_capture.sh)",
           "conv_tc_launches_per_step": sum(a[3] for _, a in conv), "conv_tc_dram_bytes_per_step": sum(a[2] for _, a in conv),
           "conv_tc_alg_bytes_per_step": sum(a[1] for _, a in conv), "conv_tc_us_per_step_serialised": sum(a[0] for _, a in conv),
           "all_kernels_us_per_step_serialised": tot}
if len(sys.argv) > 3:
    json.dump(traffic, open(sys.argv[3], "w"), indent=1)
out = "\n".join(lines) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(out)
print(out)
