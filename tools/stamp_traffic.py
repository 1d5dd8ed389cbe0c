#!/usr/bin/env python
"""profiles/traffic.json from an ncu CSV (--metrics dram__bytes_read.sum,dram__bytes_write.sum) of the
headline kernel, stamped with the sha256 of the kernel's source file: bench.py reports
`roofline.traffic` only while the stamp matches the source that built the library it measures.
usage: stamp_traffic.py ncu.csv capture-name"""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
tot, kernel = {}, None
for r in rows[1:]:
    if "phasor_stream" not in r[col["Kernel Name"]]:
        continue
    kernel = r[col["Kernel Name"]]
    val = float(r[col["Metric Value"]].replace(",", ""))
    unit = r[col["Metric Unit"]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
    tot[r[col["Metric Name"]]] = val * scale
src = os.path.join(ROOT, "codex_africanus_b200", "csrc", "afr_dft.cu")
out = {"phasor_stream_im_to_vis_cfg2": {
    "dram_bytes_per_launch": tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"],
    "dram_bytes_read": tot["dram__bytes_read.sum"], "dram_bytes_write": tot["dram__bytes_write.sum"],
    "kernel": kernel, "capture": sys.argv[2],
    "source_sha256": hashlib.sha256(open(src, "rb").read()).hexdigest(),
    "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none on the first "
           "im_to_vis launch of `bench.py --steps 1 --warmup 0` (full configs[1] shape)"}}
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
