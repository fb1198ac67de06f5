"""Stage the UNMODIFIED reference path under oracle/_ref/ so that it travels to the GPU box.  TEST INFRASTRUCTURE.

The reference is Python: there is nothing to compile.  This recipe copies, byte for byte, the few modules of
/root/reference/script that the hot path imports (models/{rendering,nerfh_nff,ray_utils,losses,poses}.py,
utils/{utils,lie_group_helper}.py and the two package __init__ files) into oracle/_ref/script/.  oracle/_ref/ is
git-ignored (it never enters the history) but not gpurun-ignored, like a built .so.  `bench.py --impl reference`
imports it through oracle/ref_loader.py and times the reference's own render() there; when oracle/_ref is absent it
falls back to the oracle port and says so (`cpu_baseline.kind`).

    python oracle/build_ref.py          # run in the build container (needs /root/reference); __graft_entry__.build() calls it
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
DST = os.path.join(HERE, "_ref")
FILES = ["script/models/__init__.py", "script/models/rendering.py", "script/models/nerfh_nff.py", "script/models/ray_utils.py",
         "script/models/losses.py", "script/models/poses.py", "script/utils/utils.py", "script/utils/lie_group_helper.py"]


def build(verbose=True) -> bool:
    if not os.path.isdir(REF):
        if verbose:
            print(f"[build_ref] {REF} not present (GPU box): using the staged copy as is" if os.path.isdir(DST) else
                  f"[build_ref] neither {REF} nor {DST}: the reference arm will fall back to the oracle port")
        return os.path.isdir(DST)
    manifest = []
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        if not os.path.exists(src):
            if rel.endswith("__init__.py"):
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                open(dst, "w").close()
                continue
            raise SystemExit(f"[build_ref] missing {src}")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest.append(f"{hashlib.sha256(open(src, 'rb').read()).hexdigest()}  {rel}")
    for pkg in ("script/utils/__init__.py",):
        p = os.path.join(DST, pkg)
        if not os.path.exists(p):
            open(p, "w").close()
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(manifest) + "\n")
    if verbose:
        print(f"[build_ref] staged {len(manifest)} reference files under {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
