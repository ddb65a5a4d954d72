"""Summarises an `ncu --metrics gpu__time_duration.sum[,launch__grid_size] --csv` launch list.
usage: python tools/ncu_launches.py launches.csv [--update]   (--update: print the last complete DDPG update, gather to gather)"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[hdr]
ki, mi, vi, gi = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("Grid Size")
seq = []
for r in rows[hdr + 1:]:
    if len(r) > vi and r[mi] == "gpu__time_duration.sum":
        seq.append((r[ki].split("(")[0][:70], float(r[vi].replace(",", "")) / 1000.0, r[gi]))
if "--update" in sys.argv:
    idx = [i for i, x in enumerate(seq) if "gather" in x[0]]
    one = seq[idx[-2]:idx[-1]]
    for x in one:
        print("%-72s %8.1f us  grid %s" % x)
    print("one update: %d kernels, %.1f us summed" % (len(one), sum(x[1] for x in one)))
else:
    agg = collections.OrderedDict()
    for k, t, g in seq:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += t
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | mean us | total ms | share |\n|---|---|---|---|---|")
    for k, (n, t) in agg.items():
        print("| %s | %d | %.2f | %.3f | %.4f |" % (k, n, t / n, t / 1e3, t / tot))
