"""Summarises an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of bench.py: picks ONE
ELBO+grad evaluation (the launches between two consecutive kbuild_kernel launches) and prints the
time share per kernel.  usage: python profiles/summarize_launches.py <launches.csv> [eval_index]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
ix = {h: i for i, h in enumerate(rows[0])}
data = rows[1:]
names = [r[ix["Kernel Name"]] for r in data]
unit = data[0][ix["Metric Unit"]]
scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1.0)
starts = [i for i, nm in enumerate(names) if "kbuild_kernel" in nm]
a, b = starts[which], starts[which + 1]
agg = collections.OrderedDict()
for r in data[a:b]:
    nm = r[ix["Kernel Name"]]
    nm = re.sub(r"gpinv::", "", nm)
    nm = re.sub(r"void ", "", nm)
    nm = re.sub(r"\(.*", "", nm)
    nm = nm[:96]
    t = float(r[ix["Metric Value"]].replace(",", "")) * scale
    c, s = agg.get(nm, (0, 0.0))
    agg[nm] = (c + 1, s + t)
total = sum(s for _, s in agg.values())
print("one ELBO+grad evaluation (launches %d..%d): %d launches, %.1f us summed (serialised, cold-cache)" % (a, b, b - a, total))
for nm, (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-96s n=%3d %9.1f us %5.1f%%" % (nm, c, s, 100 * s / total))
