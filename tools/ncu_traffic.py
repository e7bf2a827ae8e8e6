#!/usr/bin/env python
"""Joins an ncu launch list of tools/prof_stages.py (csv) with its stage list (json) and updates profiles/traffic.json:
traffic[<workload>_<precision>][<roofline_all kernel name>] = dram__bytes_read.sum + dram__bytes_write.sum of that stage's launch(es).

    python tools/ncu_traffic.py c2 f16 gpurun_out/stages_c2_f16.csv gpurun_out/stages_c2_f16.json
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
workload, precision, csv_path, json_path = sys.argv[1:5]
rows = [r for r in csv.reader(open(csv_path)) if r and not r[0].startswith("==")]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
launch = {}
for r in rows[1:]:
    if len(r) != len(hdr):
        continue
    d = launch.setdefault(int(r[ix["ID"]]), {"kernel": r[ix["Kernel Name"]]})
    val, unit = float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(unit, 1)
    d[r[ix["Metric Name"]]] = val * scale
ids = sorted(launch)
stages = json.load(open(json_path))
out, k = {}, 0
detail = []
for name, n in stages:
    tot_r = tot_w = us = 0.0
    kernels = []
    for _ in range(n):
        d = launch[ids[k]]; k += 1
        tot_r += d.get("dram__bytes_read.sum", 0.0); tot_w += d.get("dram__bytes_write.sum", 0.0); us += d.get("gpu__time_duration.sum", 0.0)
        kernels.append(d["kernel"].split("(")[0][-60:])
    out[name] = tot_r + tot_w
    detail.append({"stage": name, "launches": n, "dram_read": tot_r, "dram_write": tot_w, "us_under_ncu": us, "kernels": kernels})
assert k == len(ids), f"stage list covers {k} launches, ncu saw {len(ids)}"
path = os.path.join(ROOT, "profiles", "traffic.json")
allt = json.load(open(path)) if os.path.exists(path) else {}
allt = {k2: v for k2, v in allt.items() if isinstance(v, dict)}          # drop the round-1 flat entries
allt[f"{workload}_{precision}"] = out
allt.setdefault("_detail", {})[f"{workload}_{precision}"] = detail
json.dump(allt, open(path, "w"), indent=1)
for d in detail:
    print("%-45s x%d  read %8.2f MB  write %8.2f MB  %7.1f us" % (d["stage"], d["launches"], d["dram_read"] / 1e6, d["dram_write"] / 1e6, d["us_under_ncu"]))
