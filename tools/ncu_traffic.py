"""Write profiles/traffic.json: DRAM bytes (read + write) per launch of each stage kernel, averaged over the launches in
an `ncu --set full` report of the bench command.
    python tools/ncu_traffic.py gpurun_out/prof_X.ncu-rep"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
names = {"cell_stats": "cell_stats", "region_grow": "region_grow", "labeling": "labeling", "edge_mask": "edge_mask",
         "refine": "refine"}
acc = {}
for d in data:
    kern = d[ix["Kernel Name"]]
    key = next((v for k, v in names.items() if k in kern), None)
    if key is None:
        continue
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(d[ix[m]].replace(",", "")) * scale[units[ix[m]]]
    acc.setdefault(key, []).append(tot)
out = {k: sum(v) / len(v) for k, v in acc.items()}
out["_source"] = os.path.basename(sys.argv[1]) + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, mean per launch)"
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
