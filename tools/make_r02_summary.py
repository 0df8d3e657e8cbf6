"""Collect the round-2 evidence run (tools/final_evidence.sh r02 -> gpurun_out/) into profiles/: bench lines, per-op tables,
ncu launch lists (aggregated), the parity audit, and one summary table (profiles/r02_summary.md)."""
import csv
import json
import os
import shutil
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF, TAG = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles"), "r02"
HBM, TF_BURST, TF_SUS = 6545.0, 1628.3, 1383.2


def last_json(path):
    for line in reversed(open(path).read().strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    raise ValueError(path)


def launch_list(path):
    rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
    per = defaultdict(dict)
    for r in rows:
        per[int(r["ID"])]["name"] = r["Kernel Name"].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            v = v / 1e3 if u in ("ns", "nsecond") else v if u in ("us", "usecond") else v * 1e3
        elif r["Metric Name"].startswith("dram__bytes"):
            v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        per[int(r["ID"])][r["Metric Name"]] = v
    return [per[k] for k in sorted(per)]


def f4_section():
    """SURVEY §8 f4 (Pose / Segment heads): tools/f4_probe.py -> profiles/r02_f4_probe.json."""
    p = f"{PROF}/{TAG}_f4_probe.json"
    if not os.path.exists(p):
        return []
    d = json.load(open(p))
    md = [f"## Pose and Segment heads (SURVEY §8 f4; `tools/f4_probe.py`, {d['frames']} synthetic {d['imgsz']}x{d['imgsz']} frames, one B200)", "",
          "| head | drop-in end to end (img/s) | reference CPU predictor (img/s) | ratio | network graph ms | NMS ms | head post-process | device-timed img/s |",
          "|---|---|---|---|---|---|---|---|"]
    for k, h in d["heads"].items():
        post = (f"kept-row gather + keypoint decode + scale_coords: {h['keypoints_kernel_ms'] * 1e3:.0f} us" if k == "pose" else
                f"coefficient gather {h['coeff_gather_ms'] * 1e3:.0f} us + `process_mask_kernel` {h['process_mask_ms']:.2f} ms "
                f"({h['mask_rows']} masks, {h['mask_bytes_written'] / 1e9:.1f} GB written: {h['process_mask_gbps']:.0f} GB/s = "
                f"{h['process_mask_frac_of_hbm']:.2f} of the {HBM:.0f} GB/s copy peak)")
        md.append(f"| {k} (`{h['model']}`, `{h['predictor']}`) | {h['e2e_images_per_s']:,.0f} | {h['reference_cpu_images_per_s']:.1f} "
                  f"({h['reference_cpu_sample']}) | {h['e2e_over_reference']:.0f}x | {h['network_graph_ms']:.3f} | {h['nms_ms']:.3f} | {post} | "
                  f"{h['images_per_s_device']:,.0f} |")
    md += ["", "The synthetic checkpoints fill `max_det`: 300 detections per image, i.e. 19,200 640x640 fp32 masks per 64-frame batch -- the "
           "worst case for the mask kernel and for `Results` construction (real scenes keep tens of rows).  First version of the mask "
           "kernel (scalar 4-byte stores, per-pixel index arithmetic): 13.24 ms = 0.36 of HBM; with one column quad per thread, 16-byte "
           "stores and fixed x4 weights: 5.15 ms = 0.93.  GPU tests: `profiles/r02_pose_gpu_tests.log`, `profiles/r02_seg_gpu_tests.log`.", "",
           "ncu on the two stand-alone f4 kernels (`tools/ncu_f4.py`, `profiles/r02_ncu_f4.csv`; 2400 masks per launch, cold and serialised): "
           "`process_mask_kernel` 746 us, DRAM write 3.868 GB + read 0.018 GB against 3.932 GB algorithmic (the output) -- traffic = 0.99 x "
           "algorithmic, no re-reads (prototypes and coefficients come from L2); `gather_extras_kernel` 6.3 us, 0.34 MB read.", ""]
    return md


def family_table(launches, title):
    agg = defaultdict(lambda: [0, 0.0])
    for L in launches:
        k = L["name"]
        agg[k][0] += 1
        agg[k][1] += L.get("gpu__time_duration.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    lines = [title, f"total {tot:.0f} us over {sum(a[0] for a in agg.values())} launches (cold-cache, serialised: compare shares)", "",
             "| kernel | launches | us | share |", "|---|---|---|---|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| {k} | {a[0]} | {a[1]:.0f} | {100 * a[1] / tot:.1f}% |")
    fam = defaultdict(float)
    for k, a in agg.items():
        fam[k.split("<")[0]] += a[1]
    lines += ["", "| kernel family | us | share |", "|---|---|---|"]
    for k, v in sorted(fam.items(), key=lambda kv: -kv[1]):
        lines.append(f"| {k} | {v:.0f} | {100 * v / tot:.1f}% |")
    return lines, fam, tot


def main():
    os.makedirs(PROF, exist_ok=True)
    b = last_json(f"{OUT}/bench_{TAG}.json")
    ref = last_json(f"{OUT}/bench_ref_{TAG}.json")
    bs = last_json(f"{OUT}/bench_bsyolo_{TAG}.json")
    bx = last_json(f"{OUT}/bench_11x_{TAG}.json")
    for src, dst in ((f"bench_{TAG}.json", f"{TAG}_bench_final.json"), (f"bench_ref_{TAG}.json", f"{TAG}_bench_reference.json"),
                     (f"bench_bsyolo_{TAG}.json", f"{TAG}_bench_bsyolo.json"), (f"bench_11x_{TAG}.json", f"{TAG}_bench_yolo11x.json"),
                     (f"layers_yolo11n_{TAG}.json", f"{TAG}_layers_yolo11n.json"), (f"layers_bsyolo_{TAG}.json", f"{TAG}_layers_bsyolo.json"),
                     (f"layers_11x_{TAG}.json", f"{TAG}_layers_yolo11x.json"), ("parity_audit.jsonl", f"{TAG}_parity_audit.jsonl")):
        if os.path.exists(f"{OUT}/{src}"):
            with open(f"{PROF}/{dst}", "w") as f:
                f.write(open(f"{OUT}/{src}").read().strip().splitlines()[-1] + "\n" if src.startswith("bench") else open(f"{OUT}/{src}").read())
    md = [f"# Round-2 evidence summary (`tools/final_evidence.sh {TAG}` on one fresh B200; multi-GPU lines from `tools/multi_gpu_evidence.sh`)", "",
          f"Peaks (MEASURED_PEAKS.json): HBM {HBM:.0f} GB/s copy, bf16 {TF_BURST:.0f} TF/s burst / {TF_SUS:.0f} TF/s sustained.", ""]
    md += ["## Bench lines", "", "| workload | resident (img/s) | ms/step | end to end (img/s) | network graph ms | conv kernel roofline | network Σfloors ÷ graph | Σfloors ÷ step |",
           "|---|---|---|---|---|---|---|---|"]

    def row(name, d):
        r = d["roofline"]
        return (f"| {name} | {d['value']:,.0f} | {d['ms_per_step']:.3f} | {d['e2e']['value']:,.0f} | {d['stages']['network_graph_ms']:.3f} | "
                f"{r['frac']:.3f} of {r['peak']:.0f} {r['unit']} ({r['bound']}) | {r['network_roofline_frac']:.3f} | {r.get('step_roofline_frac', 0):.3f} |")

    md.append(row("yolo11n B=64 @640 (configs[1], headline of `bench.py`)", b))
    w = b.get("workloads", {})
    if "bsyolo_n" in w and "roofline" in w["bsyolo_n"]:
        md.append(row("BS-YOLO-n B=256 @640 (configs[2], sub-record)", w["bsyolo_n"]))
    md.append(row("BS-YOLO-n B=256 @640 (stand-alone run, per-op table)", bs))
    if "yolo11x" in w and "roofline" in w["yolo11x"]:
        md.append(row("yolo11x B=32 @1280 (configs[3], sub-record)", w["yolo11x"]))
    md.append(row("yolo11x B=32 @1280 (stand-alone run)", bx))
    md += ["", f"Reference arm (`bench.py --impl reference`, unmodified reference on the box's {ref['cpu_baseline']['cores']} host cores, "
               f"{ref['cpu_baseline']['sample']}): **{ref['value']:.1f} img/s**; in-line `cpu_baseline` of the headline run: "
               f"{b['cpu_baseline']['value']:.1f} img/s ({b['cpu_baseline']['sample']}).",
           f"Headline end-to-end ÷ reference arm: {b['e2e']['value'] / ref['value']:.0f}x.", "",
           f"Clocks during the run: {json.dumps(b['clocks'])}", ""]
    if "sahi" in w:
        s = w["sahi"]
        md.append(f"SAHI config 5, one GPU: {s.get('value', 0):,.0f} frames/s = {s.get('tiles_per_s', 0):,.0f} tiles/s resident, "
                  f"{(s.get('e2e') or {}).get('value', 0):,.0f} frames/s end to end ({s['config']['frames_per_step']} frames per step).")
    if "dropin" in w:
        d = w["dropin"]
        md.append(f"Drop-in `YOLO.predict(list of {d.get('frames_per_call')} ndarrays, predictor=B200DetectionPredictor)` incl. `Results`: "
                  f"**{d.get('value', 0):,.0f} img/s** (best of {d.get('best_of')}); host-only staging ceiling of the same frames "
                  f"(pageable -> pinned ring, all cores, no GPU work): {(d.get('host_staging_ceiling') or {}).get('value', 0):,.0f} img/s "
                  f"({(d.get('host_staging_ceiling') or {}).get('gbps', 0):.1f} GB/s); `predict_stream` from pinned batches: {b['e2e']['value']:,.0f} img/s.")
    if "ingest" in w and "value" in w["ingest"]:
        g = w["ingest"]
        md.append(f"Ingest (SURVEY §8 f2) `YOLO.predict('dir of {g['files']} PNG files', batch=64, predictor=B200DetectionPredictor)` incl. decode "
                  f"and `Results`: **{g['value']:,.0f} img/s** with `ingest.PrefetchImageLoader`, {g['with_reference_loader']:,.0f} img/s "
                  f"with the reference's sequential loader (same predictor).")
    if "val" in w and "value" in w["val"]:
        v = w["val"]
        md.append(f"Validator (SURVEY §8 f1) `YOLO.val(data, validator=B200DetectionValidator, batch=64)` on {v['images']} synthetic 640x640 files incl. "
                  f"dataloader and metrics: **{v['value']:,.0f} img/s**; per-image stage times (ms): {json.dumps(v['speed_ms_per_image'])}; "
                  f"the reference's CPU validator: {v['cpu_baseline']['value']:.1f} img/s ({v['cpu_baseline']['sample']}; its NMS hits the "
                  f"reference's own wall-clock abort, ops.py:312-314).")
    md.append("")
    # ---- ncu launch lists of bench.py's timed regions
    for tag, title in ((TAG, "yolo11n"), (f"{TAG}_bsyolo", "BS-YOLO-n")):
        p = f"{OUT}/bench_launches_{tag}.csv"
        if os.path.exists(p):
            L = launch_list(p)
            lines, fam, tot = family_table(L, f"## ncu launch list of `bench.py --workload {title.lower().replace('-', '').replace('bsyolon', 'bsyolo_n')} --steps 2 --warmup 3` "
                                              f"(timed regions only, `--metrics gpu__time_duration.sum --clock-control none`)")
            md += lines + [""]
            shutil.copy(p, f"{PROF}/{tag}_bench_launches.csv")
            if tag == TAG:
                share = fam.get("conv_tc_kernel", 0) / tot
                md.append(f"conv_tc_kernel share of the serialised launch list: {share:.3f}; share of the step claimed by the bench line "
                          f"(`roofline.share_of_step`): {b['roofline']['share_of_step']:.3f}.")
                md.append("")
    # ---- eager pass with DRAM bytes and tensor-pipe activity (yolo11n): refresh traffic.json
    p = f"{OUT}/launches_{TAG}.csv"
    if os.path.exists(p) and os.path.exists(f"{OUT}/optable_{TAG}.json"):
        L = launch_list(p)
        tab = json.load(open(f"{OUT}/optable_{TAG}.json"))
        ops = tab["ops"]
        n = tab["launches_per_pass"]
        L = L[-n:]
        conv = [(o, l) for o, l in zip(ops, L) if l["name"].startswith("conv_tc_kernel")]
        dram = sum(l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0) for _, l in conv)
        alg = sum(o.get("bytes", 0) for o, _ in conv)
        json.dump({"yolo11n": {"workload": "yolo11n", "batch": tab["batch"], "source": f"ncu launch list {TAG} (tools/ncu_capture.sh)",
                               "conv_tc_launches_per_step": len(conv), "conv_tc_dram_bytes_per_step": dram,
                               "conv_tc_alg_bytes_per_step": alg,
                               "conv_tc_us_per_step_serialised": sum(l["gpu__time_duration.sum"] for _, l in conv),
                               "all_kernels_us_per_step_serialised": sum(l["gpu__time_duration.sum"] for l in L)}},
                  open(f"{PROF}/traffic.json", "w"), indent=1)
        md += ["## One eager pass of yolo11n B=64 under ncu (time, DRAM bytes vs algorithmic bytes, tensor pipe)", "",
               f"conv_tc_kernel: {len(conv)} launches, DRAM {dram / 1e6:.0f} MB vs algorithmic {alg / 1e6:.0f} MB per pass "
               f"(ratio {dram / max(alg, 1):.2f}: L2 reuse, no wasted re-reads in aggregate) -> `profiles/traffic.json`.", "",
               "| launch | layer | kernel | us | DRAM MB | alg MB | tensor pipe % |", "|---|---|---|---|---|---|---|"]
        for o, l in sorted(zip(ops, L), key=lambda t: -t[1].get("gpu__time_duration.sum", 0))[:25]:
            md.append(f"| {o['launch']} | {o['name']} | {l['name']} | {l.get('gpu__time_duration.sum', 0):.1f} | "
                      f"{(l.get('dram__bytes_read.sum', 0) + l.get('dram__bytes_write.sum', 0)) / 1e6:.1f} | {o.get('bytes', 0) / 1e6:.1f} | "
                      f"{l.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 0):.1f} |")
        md.append("")
        shutil.copy(p, f"{PROF}/{TAG}_launches_yolo11n.csv")
    p = f"{OUT}/launches_{TAG}_11x.csv"
    if os.path.exists(p) and os.path.exists(f"{OUT}/optable_{TAG}_11x.json"):
        L = launch_list(p)
        tab = json.load(open(f"{OUT}/optable_{TAG}_11x.json"))
        ops, n = tab["ops"], tab["launches_per_pass"]
        L = L[-n:]
        md += [f"## One eager pass of yolo11x @1280, batch {tab['batch']}, under ncu: the tensor-bound convs", "",
               "| launch | layer | kernel | us | TFLOP/s | tensor pipe % | DRAM MB |", "|---|---|---|---|---|---|---|"]
        for o, l in sorted(zip(ops, L), key=lambda t: -t[1].get("gpu__time_duration.sum", 0))[:30]:
            us = l.get("gpu__time_duration.sum", 0)
            md.append(f"| {o['launch']} | {o['name']} | {l['name']} | {us:.1f} | {o.get('flops', 0) / max(us, 1e-9) / 1e6:.0f} | "
                      f"{l.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 0):.1f} | "
                      f"{(l.get('dram__bytes_read.sum', 0) + l.get('dram__bytes_write.sum', 0)) / 1e6:.0f} |")
        md.append("")
        shutil.copy(p, f"{PROF}/{TAG}_launches_yolo11x.csv")
    # ---- parity audit
    p = f"{OUT}/parity_audit.jsonl"
    if os.path.exists(p):
        md += ["## Final-box parity audit (tests/boxmatch.py; `pytest -m gpu` of this run)", "",
               "| test | reference boxes | matched | near conf | class flip | IoU tie | cascade | max_det tail | unexplained |", "|---|---|---|---|---|---|---|---|---|"]
        for line in open(p):
            a = json.loads(line)
            md.append(f"| {a['test']} | {a['total']} | {a['matched']} | {a['near_conf']} | {a['class_flip']} | {a['nms_iou_tie']} | "
                      f"{a['nms_cascade']} | {a['max_det_cut']} | {a['unexplained']} |")
        md.append("")
    md += f4_section()
    open(f"{PROF}/{TAG}_summary.md", "w").write("\n".join(md) + "\n")
    print("\n".join(md[:40]))


if __name__ == "__main__":
    sys.exit(main())
