"""Build libafricanus_b200.so in-tree with nvcc for sm_100a (no GPU needed)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libafricanus_b200.so")


def build(force=False, jobs=None, verbose=False):
    """Compile every CUDA source for sm_100a (-gencode arch=compute_100a,code=sm_100a
    -lineinfo, see csrc/Makefile) and link the C-ABI shared library."""
    jobs = jobs or os.cpu_count() or 4
    cmd = ["make", "-C", CSRC, "-j%d" % jobs]
    if force:
        cmd.append("-B")
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("build finished but %s is missing" % LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(verbose=True))
