"""Compact per-launch table from an `ncu --csv --metrics ...` log:  python tools/ncu_launch_table.py file.csv"""
import collections
import csv
import sys

rows = [l for l in open(sys.argv[1]) if not l.startswith("==")]
d = collections.OrderedDict()
for r in csv.DictReader(rows):
    d.setdefault((int(r["ID"]), r["Kernel Name"][:44]), {})[r["Metric Name"]] = r["Metric Value"]
for (i, k), v in d.items():
    print(f"{i:4d} {k:44s} " + " ".join(f"{m.split('__')[-1][:18]}={x}" for m, x in v.items()))
