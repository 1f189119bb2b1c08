#!/usr/bin/env python3
"""Print kernel name + device time (us) from an ncu `--metrics gpu__time_duration.sum --csv` log."""
import csv
import sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot = 0.0
for r in rows[1:]:
    v = float(r[iv].replace(",", "")) / 1e3
    tot += v
    print(f"{v:9.2f} us  {r[ik][:90]}")
print(f"{tot:9.2f} us  total of {len(rows) - 1} launches")
