#!/usr/bin/env python
"""
Install the UNMODIFIED reference (ratt-ru/codex-africanus) into baseline/_ref so that bench.py can
time the reference's own numba kernels on the GPU box's host cores.

The sanctioned recipe
    python -m pip install --no-index --no-build-isolation --no-deps --target baseline/_ref /root/reference
fails in this image: the reference's build backend (hatchling) is not installed and there is no
network.  The package is pure Python, so what that command would produce is a copy of its
`africanus/` tree; this script makes exactly that copy.  baseline/_ref is git-ignored (no reference
source enters the history) but not gpurun-ignored, so it travels to the GPU box with the snapshot.
Nothing on the product path or in the tests imports it; bench.py falls back to the C port of the
oracle when it is absent.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")


def main():
    pkg = os.path.join(SRC, "africanus")
    if not os.path.isdir(pkg):
        print("install_reference: %s not found (GPU box?): nothing to do" % pkg)
        return 0
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    shutil.copytree(pkg, os.path.join(DST, "africanus"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "tests"))
    print("installed", pkg, "->", DST)
    return 0


if __name__ == "__main__":
    sys.exit(main())
