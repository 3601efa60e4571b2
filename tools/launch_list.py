"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` log into a per-kernel launch list: tools/launch_list.py raw.csv "header line" """
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(r[4], [0.0, 0])
    a[0] += float(r[14]) / 1e3; a[1] += 1
tot = sum(a[0] for a in agg.values())
print(sys.argv[2] if len(sys.argv) > 2 else "")
for k, (us, n) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print(f"{us:14.1f} us  {100 * us / tot:6.2f}%  x{n:<3d} {k}")
