"""Aggregate an `ncu --page source --csv` dump: executed-instruction mix per opcode, stall samples.
    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME --launch-count 1 > src.csv
    python tools/ncu_source_mix.py src.csv [top_lines]
"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
data = []
for r in rows[1:]:
    if r == hdr:
        break  # next launch of the dump: keep the first only
    data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
ops, samples, stall = collections.Counter(), collections.Counter(), collections.Counter()
tot = totsamp = 0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
lines = []
for d in data:
    src = d[ix["Source"]].strip()
    m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)", src)
    op = m.group(2) if m else src[:10]
    n, s = int(d[ix["Instructions Executed"]]), int(d[ix["# Samples"]])
    ops[op] += n
    samples[op] += s
    tot += n
    totsamp += s
    for c in stall_cols:
        stall[c] += int(d[ix[c]])
    lines.append((s, n, src))
print("total warp-instructions", tot, "samples", totsamp)
for op, n in ops.most_common(30):
    print(f"{op:10s} {n:11d} {100 * n / tot:5.1f}%   samples {100 * samples[op] / max(totsamp, 1):5.1f}%")
print("stalls:", [(k, f"{100 * v / max(totsamp, 1):.1f}%") for k, v in stall.most_common(8)])
top = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for s, n, src in sorted(lines, reverse=True)[:top]:
    print(f"{s:6d} {n:9d}  {src}")
