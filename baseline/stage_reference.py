#!/usr/bin/env python
"""Stage the UNMODIFIED reference under ``baseline/_ref/`` (git-ignored, travels to the GPU box with the
snapshot) so that ``bench.py --impl reference`` and the ``gpu_eager_baseline`` leg can import the reference's own
modules there.  The reference is a flat directory of pure-Python files with no setup.py / pyproject, so
``pip install /root/reference`` does not apply: the "install" is a byte-for-byte copy of its ``*.py`` / ``*.sh`` files.

    python baseline/stage_reference.py [/root/reference]

Nothing under ``baseline/_ref`` is ever committed or imported by the product package.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")


def stage(src="/root/reference"):
    if not os.path.isdir(src):
        return False
    os.makedirs(DST, exist_ok=True)
    n = 0
    for name in sorted(os.listdir(src)):
        if not name.endswith((".py", ".sh")):
            continue
        a, b = os.path.join(src, name), os.path.join(DST, name)
        if not (os.path.exists(b) and filecmp.cmp(a, b, shallow=False)):
            shutil.copyfile(a, b)
        n += 1
    with open(os.path.join(DST, "STAGED_FROM"), "w") as f:
        f.write("%s (%d files, byte-for-byte)\n" % (src, n))
    return True


if __name__ == "__main__":
    ok = stage(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("staged" if ok else "reference tree not found; nothing staged", DST)
