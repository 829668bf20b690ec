"""Print (kernel, grid, block, ns) rows from an `ncu --metrics gpu__time_duration.sum --csv` log."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
h = None
for r in rows:
    if "Kernel Name" in r:
        h = r
        continue
    if h and len(r) == len(h):
        d = dict(zip(h, r))
        print(re.sub(r"\(.*", "", d["Kernel Name"])[-64:], d["Grid Size"], d["Block Size"], d["Metric Value"])
