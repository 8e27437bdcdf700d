"""Opcode histogram and hottest SASS instructions of one kernel from `ncu -i X.ncu-rep --page source --csv`."""
import collections
import csv
import sys

path, which = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(open(path)))
ks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        ks.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
k = ks[which]
h = k["hdr"]
iS, iE, iN = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
tot = sum(int(r[iE]) for r in k["rows"])
totS = sum(int(r[iN]) for r in k["rows"])
print(k["name"][:70], len(k["rows"]), "sass instr; executed", tot, "samples", totS)
byop, byopS = collections.Counter(), collections.Counter()
for r in k["rows"]:
    t = r[iS].split()
    op = t[1] if t[0].startswith("@") else t[0]
    op = op.split(".")[0]
    byop[op] += int(r[iE])
    byopS[op] += int(r[iN])
for op, c in byop.most_common(30):
    print("%-10s exec %5.1f%%  samples %5.1f%%" % (op, 100 * c / tot, 100 * byopS[op] / totS))
print("--- hottest by samples")
for i in sorted(range(len(k["rows"])), key=lambda i: -int(k["rows"][i][iN]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
    r = k["rows"][i]
    print("%5d  samples %5.2f%%  exec %5.2f%%  %s" % (i, 100 * int(r[iN]) / totS, 100 * int(r[iE]) / tot, r[iS][:90]))
