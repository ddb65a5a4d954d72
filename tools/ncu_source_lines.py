"""Aggregates an `ncu -i rep --page source --csv --print-source cuda,sass` export per CUDA source line:
share of warp-stall samples (where the warps' time goes) and of executed instructions.
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:<k> > src.csv; python tools/ncu_source_lines.py src.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.OrderedDict()
for r in rows[3:]:
    if len(r) < 8 or not r[0] or r[0] == "Line No":
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    smp = int(r[4]) if r[4].isdigit() else 0
    ins = int(r[7]) if r[7].isdigit() else 0
    a = agg.setdefault(ln, [r[1][:100], 0, 0])
    a[1] += smp
    a[2] += ins
ts = sum(a[1] for a in agg.values()) or 1
ti = sum(a[2] for a in agg.values()) or 1
print("total stall samples %d, warp instructions %d" % (ts, ti))
for ln, a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print("%4d %5.1f%% samples %5.1f%% instr  %s" % (ln, 100 * a[1] / ts, 100 * a[2] / ti, a[0]))
