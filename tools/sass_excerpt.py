"""SASS evidence per MMA kernel of libnefes_b200.so: counts of the Blackwell tensor / TMEM / bulk-copy mnemonics and the first
few of each.  Usage: python tools/sass_excerpt.py > profiles/r2_sass_mma_kernels.txt"""
import os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "nefes_b200", "lib", "libnefes_b200.so")], capture_output=True, text=True).stdout
cur, body = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); body[cur] = []
    elif cur and re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", line) and not re.match(r"\s*/\* 0x", line):
        body[cur].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line).strip())
KEYS = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "ELECT", "HMMA", "FFMA")
print("cuobjdump -sass nefes_b200/lib/libnefes_b200.so, sm_100a.  UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st,")
print("UBLKCP = cp.async.bulk (1-D TMA), SYNCS = mbarrier ops, ELECT = elect.sync.  No HMMA (mma.sync) in any tensor kernel.\n")
for fn, ls in body.items():
    cnt = {k: sum(1 for l in ls if re.search(r"\b" + k, l)) for k in KEYS}
    if cnt["UTCHMMA"] == 0:
        continue
    demangled = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
    print(f"== {demangled}   ({len(ls)} SASS instructions)")
    print("   " + "  ".join(f"{k} {v}" for k, v in cnt.items()))
    for k in ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP"):
        for l in [l for l in ls if re.search(r"\b" + k, l)][:2]:
            print("      " + l)
    print()
