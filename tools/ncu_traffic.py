"""profiles/ncu_traffic.json from an ncu --csv launch list that carries dram__bytes_read/write.sum:
    python tools/ncu_traffic.py gpurun_out/step_launches_r1.csv > profiles/ncu_traffic.json"""
import collections
import csv
import json
import re
import sys

lines = [l for l in open(sys.argv[1], newline="") if not l.startswith("==")]
agg = collections.defaultdict(lambda: {"ids": set(), "bytes": 0.0, "us": 0.0})
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
for r in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("csd::", "")
    name = re.sub(r"<.*", "", name)
    a = agg[name]
    a["ids"].add(r["ID"])
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    if r["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        a["bytes"] += v * scale.get(r["Metric Unit"], 1)
    elif r["Metric Name"] == "gpu__time_duration.sum":
        a["us"] += v * tscale.get(r["Metric Unit"], 1.0)
out = {k: {"launches": len(a["ids"]), "dram_bytes_per_launch": a["bytes"] / max(1, len(a["ids"])),
           "us_per_launch_under_ncu": a["us"] / max(1, len(a["ids"]))} for k, a in agg.items()}
out["_source"] = sys.argv[1]
print(json.dumps(out, indent=1))
