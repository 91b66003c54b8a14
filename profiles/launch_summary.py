"""Summarise an ncu launch list (--csv with gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum):
   python profiles/launch_summary.py gpurun_out/launches.csv [out.json]
per kernel: launches, total device time, DRAM bytes; shares of the step."""
import csv
import json
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
ix = {n: i for i, n in enumerate(h)}
per = defaultdict(dict)
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    u = r[ix["Metric Unit"]]
    m = r[ix["Metric Name"]]
    if m == "gpu__time_duration.sum":
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)
    elif m.startswith("dram__bytes"):
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1.0)
    per[(r[ix["ID"]], r[ix["Kernel Name"]])][m] = v
agg = defaultdict(lambda: defaultdict(float))
for (_, name), d in per.items():
    k = name.split("(")[0].replace("void ", "").replace("e2i::", "")
    agg[k]["launches"] += 1
    agg[k]["ms"] += d.get("gpu__time_duration.sum", 0.0)
    agg[k]["dram_bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
tot = sum(a["ms"] for a in agg.values())
out = {}
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    out[k] = {"launches": int(a["launches"]), "ms": round(a["ms"], 3), "share": round(a["ms"] / tot, 4),
              "dram_bytes": int(a["dram_bytes"]), "dram_tb_s": round(a["dram_bytes"] / (a["ms"] / 1e3) / 1e12, 3) if a["ms"] else None}
    print(f"{k:48s} {out[k]['launches']:6d} launches {out[k]['ms']:10.2f} ms {100 * out[k]['share']:5.1f} %  {out[k]['dram_bytes'] / 1e9:9.2f} GB  {out[k]['dram_tb_s']} TB/s")
if len(sys.argv) > 2:
    json.dump({"total_ms": tot, "kernels": out}, open(sys.argv[2], "w"), indent=1)
