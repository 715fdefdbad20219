"""Summarise an ncu report (one row per profiled launch) as a markdown table for profiles/.
    python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep > profiles/X_summary.md
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
cols = [
    ("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("smsp__inst_executed.sum", "warp inst"),
]
cols = [(c, n) for c, n in cols if c in ix]
print("| " + " | ".join(n for _, n in cols) + " |")
print("|" + "---|" * len(cols))
for d in data:
    cells = []
    for c, _ in cols:
        v, u = d[ix[c]], units[ix[c]]
        if c == "Kernel Name":
            v = v.replace("void ", "").replace("dpx::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
            v = v.split("(")[0]
        else:
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.3f}".rstrip("0").rstrip(".") if f < 1e6 else f"{f:.4g}"
            except ValueError:
                pass
            if u and u not in ("%",):
                v += " " + u
        cells.append(v)
    print("| " + " | ".join(cells) + " |")
