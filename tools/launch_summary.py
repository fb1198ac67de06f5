"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: one engine step, by kernel.
Usage: launch_summary.py list.csv [step index] [name of the kernel that starts a step, default get_rays_fwd]"""
import csv, collections, sys
path = sys.argv[1]
step = int(sys.argv[2]) if len(sys.argv) > 2 else 3
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = [(x['Kernel Name'], float(x['Metric Value'].replace(',', ''))) for x in csv.DictReader(lines)]
marker = sys.argv[3] if len(sys.argv) > 3 else 'get_rays_fwd'
idx = [i for i, (n, _) in enumerate(rows) if marker in n]
s, e = idx[step], idx[step + 1] if step + 1 < len(idx) else len(rows)
agg, tot = collections.OrderedDict(), 0.0
for n, t in rows[s:e]:
    k = n.split('(')[0][-70:]
    a = agg.setdefault(k, [0.0, 0]); a[0] += t; a[1] += 1; tot += t
print(f"# {path}: step {step}, {e - s} launches, {tot / 1e3:.1f} us of kernel time (ncu serialises launches and runs them cold: compare SHARES)")
for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:22]:
    print(f"{t / 1e3:10.1f} us {100 * t / tot:5.1f}%  x{c:3d}  {k}")
