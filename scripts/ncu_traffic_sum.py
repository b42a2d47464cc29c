"""Sums dram bytes / durations of the LAST step out of an ncu --csv log produced for scripts/traffic_headline.py (3 identical steps).
Usage: python scripts/ncu_traffic_sum.py log.csv [kernel-name-substring ...]"""
import csv
import io
import json
import sys

rows = []
with open(sys.argv[1]) as f:
    text = f.read()
start = text.find('"ID"')
for r in csv.DictReader(io.StringIO(text[start:])):
    rows.append(r)
names = sys.argv[2:] or ["q1hex_gather", "cell_geom", "q1hex_general", "vector_kernel", "q2_elasticity", "generic", "q1hex_rhs", "Memset", "memset", "bog_gather", "cng_gather", "cng_factors", "nh_q1"]
kern = {}
for r in rows:
    kn = r["Kernel Name"]
    if not any(s in kn for s in names):
        continue
    kid = int(r["ID"])
    d = kern.setdefault(kid, {"name": kn})
    val = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3}.get(unit, 1)
    d[r["Metric Name"]] = val * scale
ids = sorted(kern)
n = len(ids) // 3
last = ids[-n:] if n else ids
tot = {"kernels_per_step": len(last), "dram_read": 0.0, "dram_write": 0.0, "time_s": 0.0, "by_kernel": {}}
for i in last:
    d = kern[i]
    short = d["name"].split("(")[0].split("::")[-1][:60]
    b = tot["by_kernel"].setdefault(short, {"launches": 0, "dram_read": 0.0, "dram_write": 0.0, "time_s": 0.0})
    b["launches"] += 1
    for k_src, k_dst in (("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"), ("gpu__time_duration.sum", "time_s")):
        tot[k_dst] += d.get(k_src, 0.0)
        b[k_dst] += d.get(k_src, 0.0)
tot["dram_bytes_per_step"] = tot["dram_read"] + tot["dram_write"]
print(json.dumps(tot))
