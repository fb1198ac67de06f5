"""Top stall locations (by sampled warps) from an `ncu --page source --csv --print-source cuda,sass` dump."""
import csv, io, re, sys
txt = open(sys.argv[1]).read()
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for part in re.split(r'(?m)^"File Path",', txt)[1:]:
    lines = part.split('\n')
    rows = list(csv.reader(io.StringIO('\n'.join(lines[2:]))))
    hdr = rows[0]
    si = hdr.index("# Samples")
    stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    data = []
    cur_src = ""
    for r in rows[1:]:
        if len(r) <= si: continue
        if r[0].strip().isdigit() and r[1].strip() not in ("", "-"): cur_src = r[1].strip()
        try: n = int(r[si])
        except ValueError: continue
        data.append((n, r, cur_src))
    tot = sum(n for n, _, _ in data)
    print("=====", lines[1][:110], "samples", tot)
    for n, r, src in sorted(data, key=lambda x: -x[0])[:topn]:
        st = sorted(((hdr[i], float(r[i])) for i in stall if r[i] not in ('0', '')), key=lambda kv: -kv[1])[:2]
        print(f"{n:7d} {100 * n / max(tot, 1):5.1f}%  {r[3][:58]:58s} {st}  | {src[:70]}")
