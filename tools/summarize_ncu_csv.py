"""Summarise an ncu --csv log (one row per metric per launch) into per-kernel tables.

    python tools/summarize_ncu_csv.py gpurun_out/forward_sol_r1.csv > profiles/forward_sol_r1.md
"""
import collections
import csv
import re
import sys

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
launch = collections.OrderedDict()
for r in rd:
    key = r["ID"]
    d = launch.setdefault(key, {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    d[r["Metric Name"]] = (v, r["Metric Unit"])


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("csd::", "")
    return name[:48]


def get(d, k, unit_scale=None):
    if k not in d:
        return None
    v, u = d[k]
    if k == "gpu__time_duration.sum" or k == "Duration":
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        return v * scale   # us
    if u in ("Kbyte", "Mbyte", "Gbyte", "byte"):
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    return v


agg = collections.OrderedDict()
total_us = 0.0
for key, d in launch.items():
    dur = get(d, "gpu__time_duration.sum")
    if dur is None:
        dur = get(d, "Duration")
    if dur is None:
        continue
    total_us += dur
    a = agg.setdefault(short(d["name"]), {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "dram_pct_w": 0.0, "cmp_pct_w": 0.0})
    a["n"] += 1
    a["us"] += dur
    a["rd"] += get(d, "dram__bytes_read.sum") or 0.0
    a["wr"] += get(d, "dram__bytes_write.sum") or 0.0
    a["dram_pct_w"] += (get(d, "DRAM Throughput") or 0.0) * dur
    a["cmp_pct_w"] += (get(d, "Compute (SM) Throughput") or 0.0) * dur

print(f"source: {path}")
print(f"launches: {len(launch)}   total device time: {total_us / 1e3:.3f} ms (cold-cache, serialised under ncu: compare shares)")
print()
print("| kernel | launches | time ms | share | DRAM read GB | DRAM write GB | DRAM GB/s | DRAM thr % (time-weighted) | SM thr % |")
print("|---|---|---|---|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    gbs = (a["rd"] + a["wr"]) / (a["us"] * 1e-6) / 1e9 if a["us"] else 0.0
    print(f"| {k} | {a['n']} | {a['us'] / 1e3:.3f} | {100 * a['us'] / total_us:.1f}% | {a['rd'] / 1e9:.3f} | {a['wr'] / 1e9:.3f} | "
          f"{gbs:.0f} | {a['dram_pct_w'] / a['us'] if a['us'] else 0:.1f} | {a['cmp_pct_w'] / a['us'] if a['us'] else 0:.1f} |")
