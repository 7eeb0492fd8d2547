"""Condense the ncu launch lists of a round into small files under profiles/.

    python tools/summarize_launches.py <round-tag> [gpurun_out]

Inputs (written by tools/gpu_final.sh):
  launches_bench.csv.gz    `ncu --metrics gpu__time_duration.sum --clock-control none` of `bench.py --steps 2 --warmup 1`
  launches_forward.csv.gz  same plus dram__bytes_read/write of one un-graphed forward (tools/gpu_ncu_forward.py)
Outputs:
  profiles/<tag>_launches_bench_summary.txt   per-kernel launch count / total time / share (cold-cache, serialised)
  profiles/<tag>_dram_traffic.json            mean DRAM bytes per launch per kernel (bench.py reads it for roofline.traffic)
"""
import csv
import gzip
import json
import os
import re
import sys
from collections import defaultdict

tag = sys.argv[1]
src = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")


def rows(path):
    with gzip.open(path, "rt", errors="replace") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    return list(csv.DictReader(lines))


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name if len(name) < 90 else name[:87] + "..."


def to_float(v):
    return float(v.replace(",", ""))


def unit_scale(unit):
    return {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6,
            "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


bench = os.path.join(src, "launches_bench.csv.gz")
if os.path.exists(bench):
    per = defaultdict(lambda: [0, 0.0])
    for r in rows(bench):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        k = short(r["Kernel Name"])
        per[k][0] += 1
        per[k][1] += to_float(r["Metric Value"]) * unit_scale(r["Metric Unit"])
    total = sum(v[1] for v in per.values()) or 1.0
    ours = ("conv_gemm", "attention", "gn_", "resample", "temporal_attn", "im2col", "time_embed", "emb_layers",
            "p_sample_tail", "set_shifts", "pack_", "add_vec", "q_sample", "wgrad", "attn_bwd", "colsum", "lincomb")
    with open(os.path.join(OUT, f"{tag}_launches_bench_summary.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 1 "
                "--no-cpu-baseline --profile-reps 1\n# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"# {sum(v[0] for v in per.values())} launches, {total / 1e3:.2f} ms summed kernel time\n")
        f.write(f"{'kernel':90s} {'launches':>8s} {'total_us':>12s} {'share':>7s} {'ours':>5s}\n")
        for k, (n, us) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            mine = any(t in k for t in ours)
            f.write(f"{k:90s} {n:8d} {us:12.1f} {100 * us / total:6.2f}% {'yes' if mine else 'no':>5s}\n")
    print("wrote", f"{tag}_launches_bench_summary.txt")

fwd = os.path.join(src, "launches_forward.csv.gz")
if os.path.exists(fwd):
    per = defaultdict(lambda: {"launches": set(), "us": 0.0, "rd": 0.0, "wr": 0.0})
    for r in rows(fwd):
        k = short(r["Kernel Name"])
        k = re.sub(r"<.*$", "", k)
        e = per[k]
        e["launches"].add(r["ID"])
        val = to_float(r["Metric Value"]) * unit_scale(r["Metric Unit"])
        if r["Metric Name"] == "gpu__time_duration.sum":
            e["us"] += val
        elif r["Metric Name"] == "dram__bytes_read.sum":
            e["rd"] += val
        elif r["Metric Name"] == "dram__bytes_write.sum":
            e["wr"] += val
    out = {"batch": 4, "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                                 "--clock-control none python tools/gpu_ncu_forward.py 4 (one un-graphed p_sample step)",
           "kernels": {}}
    for k, e in per.items():
        n = len(e["launches"])
        if n == 0:
            continue
        out["kernels"][k] = {"launches": n, "us_total": round(e["us"], 1),
                             "dram_read_bytes_total": int(e["rd"]), "dram_write_bytes_total": int(e["wr"]),
                             "dram_bytes_per_launch": int((e["rd"] + e["wr"]) / n)}
    # bench.py groups the GroupNorm kernels under one key
    ga, gs = out["kernels"].get("gn_apply_kernel"), out["kernels"].get("gn_stats_kernel")
    if ga:
        tot = ga["dram_read_bytes_total"] + ga["dram_write_bytes_total"] + ((gs["dram_read_bytes_total"] + gs["dram_write_bytes_total"]) if gs else 0)
        out["kernels"]["gn_apply_kernel+gn_stats_kernel"] = {"launches": ga["launches"] + (gs["launches"] if gs else 0),
                                                              "dram_bytes_per_launch": int(tot / (ga["launches"] + (gs["launches"] if gs else 0)))}
    with open(os.path.join(OUT, f"{tag}_dram_traffic.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", f"{tag}_dram_traffic.json")
