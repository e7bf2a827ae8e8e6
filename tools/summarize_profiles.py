#!/usr/bin/env python
"""Turns the raw ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python tools/summarize_profiles.py launches gpurun_out/launches.csv profiles/rNN_launches.md
    python tools/summarize_profiles.py full gpurun_out/prof.ncu-rep profiles/rNN_ncu_full.md [profiles/traffic.json key=kernel_regex ...]
"""
import collections
import csv
import json
import re
import subprocess
import sys


def launches(src, dst):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg, seq = collections.OrderedDict(), []
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        v = v / 1000 if row["Metric Unit"] == "ns" else (v * 1000 if row["Metric Unit"] == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("cb2::<unnamed>::", "").replace("void ", "")[:60]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        seq.append((name, v))
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none): {len(seq)} launches, {tot / 1000:.2f} ms\n\n")
        f.write("Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1] / 1000:.2f} | {a[1] / a[0]:.1f} | {100 * a[1] / tot:.1f} % |\n")
        step = [i for i, (n, _) in enumerate(seq) if "node_tc" in n or "node_update" in n]
        if len(step) > 40:
            i0 = step[24]
            f.write("\nOne diffusion step (16 launches), us:\n\n```\n")
            for n, v in seq[i0:i0 + 16]:
                f.write(f"{n:40s} {v:8.1f}\n")
            f.write("```\n")


METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "smsp__inst_executed.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]


def full(src, dst, extra):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none ({src})\n\n")
        for r in body:
            f.write(f"## {r[col['Kernel Name']]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for m in METRICS:
                if m in col:
                    f.write(f"| {m} | {r[col[m]]} | {units[col[m]]} |\n")
            f.write("\n")
    if extra:
        path = extra[0]
        try:
            traffic = json.load(open(path))
        except (OSError, ValueError):
            traffic = {}
        for kv in extra[1:]:
            key, pat = kv.split("=", 1)
            for r in body:
                if re.search(pat, r[col["Kernel Name"]]):
                    def to_bytes(name):
                        v, u = float(r[col[name]].replace(",", "")), units[col[name]].lower()
                        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
                    traffic[key] = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
                    break
        json.dump(traffic, open(path, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4:])
