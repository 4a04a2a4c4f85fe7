#!/usr/bin/env python
"""Print the per-launch durations of the LAST minibatch step (from the last gather_kernel on) of an
`ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[start]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
gi = h.index("Grid Size") if "Grid Size" in h else None
out = []
for r in rows[start + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("void ", "")[-44:]
    out.append((name, float(r[vi].replace(",", "")) / 1000.0, r[gi] if gi else ""))
anchor = sys.argv[2] if len(sys.argv) > 2 else "gather_kernel"
idx = [i for i, o in enumerate(out) if anchor in o[0]]
seg = out[idx[-1]:] if idx else out
tot = 0
for n, t, g in seg:
    print(f"{t:9.1f} us  {g:>14}  {n}")
    tot += t
print(f"total {tot:.1f} us over {len(seg)} launches")
