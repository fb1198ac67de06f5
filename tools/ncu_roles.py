"""Warp-sampling histogram of an ncu report's source page: samples and top stall reasons per block of SASS lines, tagged with the
tell-tale opcodes of the block (UTCHMMA = MMA issuer, LDTM/STTM = epilogue, UBLKCP = producer).  Usage: python tools/ncu_roles.py rep [lines]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; step = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
num = lambda x: float(x) if x.replace('.', '', 1).isdigit() else 0.0
tot = sum(num(r[ix['# Samples']]) for r in data)
print(rows[0][1], 'total samples', int(tot), 'SASS lines', len(data))
for a in range(0, len(data), step):
    blk = data[a:a + step]
    n = sum(num(r[ix['# Samples']]) for r in blk)
    if n < tot * 0.002: continue
    agg = {s: sum(num(r[ix[s]]) for r in blk) for s in stalls}
    top = sorted(agg.items(), key=lambda x: -x[1])[:4]
    ops = sorted({k for r in blk for k in ('UTCHMMA', 'LDTM', 'STTM', 'UBLKCP', 'STG', 'STS', 'LDS', 'BAR.SYNC', 'SYNCS.ARRIVE', 'TRYWAIT', 'UTCBAR', 'F2FP', 'MUFU', 'RED') if k in r[ix['Source']]})
    print(f"{a:5d} {int(n):7d} {n / tot:6.1%}  " + ' '.join(f"{k[6:]}={int(v)}" for k, v in top) + '  ' + ','.join(ops))
