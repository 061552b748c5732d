#!/usr/bin/env python
"""Stage the UNMODIFIED reference (victorca25/traiNNer, /root/reference/codes) under baseline/_ref/codes.

baseline/_ref/ is git-ignored (the reference is not product source) but NOT gpurun-ignored, so the
staged tree travels to the GPU box, where /root/reference does not exist.  The reference is plain
Python with no setup.py / pyproject.toml, so `pip install --target baseline/_ref /root/reference`
has nothing to build (recorded in DESIGN.md); a byte-for-byte copy of the `codes/` tree is the
install.  Nothing under trainner_b200/ imports it: it is used by
  * bench.py --impl reference / reference-cudnn (the reference arms), and
  * tests/test_reference_parity_gpu.py (full-size parity against the reference's own SRModel).

    python tools/stage_reference.py [--src /root/reference] [--force]
"""
import argparse
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")
KEEP_EXT = (".py", ".yml", ".yaml", ".json", ".md", ".txt")


def stage(src="/root/reference", force=False, quiet=False):
    codes = os.path.join(src, "codes")
    if not os.path.isdir(codes):
        if not quiet:
            print("stage_reference: %s not found (GPU box?) -- keeping whatever is staged" % codes)
        return os.path.isdir(os.path.join(DST, "codes"))
    dst_codes = os.path.join(DST, "codes")
    if force and os.path.isdir(DST):
        shutil.rmtree(DST)
    n = 0
    for d, _dirs, files in os.walk(codes):
        rel = os.path.relpath(d, codes)
        for f in files:
            if not f.endswith(KEEP_EXT):
                continue
            s = os.path.join(d, f)
            t = os.path.join(dst_codes, rel, f)
            if os.path.exists(t) and filecmp.cmp(s, t, shallow=False):
                continue
            os.makedirs(os.path.dirname(t), exist_ok=True)
            shutil.copyfile(s, t)
            n += 1
    for f in ("LICENSE", "requirements.txt"):
        s = os.path.join(src, f)
        if os.path.exists(s):
            shutil.copyfile(s, os.path.join(DST, f))
    if not quiet:
        print("stage_reference: %d file(s) copied to %s" % (n, dst_codes))
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    sys.exit(0 if stage(a.src, a.force) else 1)
