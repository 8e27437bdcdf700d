"""Per-kernel totals of the LAST step in an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of a run that made
`nsteps` identical steps.    python tools/launch_summary.py launches.csv [nsteps] [top]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]
iK, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
data = rows[1:]
n = len(data) // nsteps
tot, cnt = collections.Counter(), collections.Counter()
for r in data[-n:]:
    v = float(r[iV].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iU], 1.0)
    name = re.sub(r"\(.*", "", r[iK])
    name = re.sub(r"void |lisec::|\(anonymous namespace\)::|<unnamed>::|at::native::", "", name)[:72]
    tot[name] += v
    cnt[name] += 1
s = sum(tot.values())
print("%d launches in the file, last step: %d launches, %.1f us (serialised, cold caches)" % (len(data), n, s))
for k, v in tot.most_common(top):
    print("%9.1f us %5.1f%% x%-3d %s" % (v, 100 * v / s, cnt[k], k))
