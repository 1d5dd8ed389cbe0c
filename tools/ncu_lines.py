#!/usr/bin/env python
"""Stall samples of an ncu report aggregated per CUDA source line (needs -lineinfo and
--import-source on).  Usage: ncu_lines.py rep [top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda",
                      "--launch-skip", "0", "--launch-count", "1"], capture_output=True, text=True).stdout
cur_file = None
rows = []
hdr = None
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        hdr = None
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None:
        continue
    d = dict(zip(hdr, r))
    try:
        n = int(d.get("# Samples", "0") or 0)
    except ValueError:
        continue
    if n == 0:
        continue
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v not in ("", "0")}
    rows.append((n, cur_file, r[0], r[1].strip()[:110], stalls))
agg = {}
for n, f, ln, src, st in rows:
    a = agg.setdefault((f, ln), [0, src, {}])
    a[0] += n
    for k, v in st.items():
        a[2][k] = a[2].get(k, 0) + v
tot = sum(x[0] for x in rows)
print("total samples", tot)
for (f, ln), (n, src, st) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:4])
    print("%6d %5.1f%%  %s:%s  %s\n        %s" % (n, 100.0 * n / tot, f, ln, st, src))
