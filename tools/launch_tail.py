"""Summarise the LAST 1/k of an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel (a script that repeats the same work
k times).  Usage: launch_tail.py list.csv [k=3]"""
import csv, collections, sys
k = int(sys.argv[2]) if len(sys.argv) > 2 else 3
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = [(x['Kernel Name'], float(x['Metric Value'].replace(',', ''))) for x in csv.DictReader(lines)]
last = rows[-(len(rows) // k):]
tot = sum(v for _, v in last)
agg = collections.OrderedDict()
for n, v in last:
    a = agg.setdefault(n.split('(')[0][-70:], [0, 0.0]); a[0] += 1; a[1] += v
print(f"# {sys.argv[1]}: last of {k} repeats, {len(last)} launches, {tot / 1e3:.1f} us of kernel time (ncu serialises launches and runs them cold: compare SHARES)")
for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:30]:
    print(f"{v / 1e3:10.1f} us {100 * v / tot:5.1f}%  x{c:3d}  {n}")
