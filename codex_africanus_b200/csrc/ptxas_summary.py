#!/usr/bin/env python
"""Summarise nvcc -Xptxas -v logs: one line per kernel (registers, stack, spills)."""
import re
import subprocess
import sys


def demangle(name):
    try:
        return subprocess.check_output(["c++filt", name], text=True).strip()
    except Exception:
        return name


def main(paths):
    pat = re.compile(
        r"Compiling entry function '([^']+)' for 'sm_100a'\n"
        r".*?Function properties for \1\n"
        r"\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
        r".*?Used (\d+) registers([^\n]*)", re.S)
    for path in paths:
        txt = open(path).read()
        for m in pat.finditer(txt):
            name, stack, sst, sld, regs, rest = m.groups()
            dm = demangle(name)
            dm = re.sub(r"afr::\(anonymous namespace\)::", "", dm)
            dm = re.sub(r"\((afr::)?\(?anonymous.*$", "", dm)
            dm = re.sub(r"^void ", "", dm)
            print("%4s regs  stack %4s  spill %4s/%-4s  %s" % (regs, stack, sst, sld, dm))


if __name__ == "__main__":
    main(sys.argv[1:])
