#!/usr/bin/env python
"""DRAM traffic per kernel launch from an `ncu --set full` capture of one RK3 step's first two
stages (x, y, z sweeps of stage 1, then of stage 2) -> profiles/*_traffic.json, the file
bench.py's roofline.traffic reads:
    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep CELLS > profiles/r01_v4_traffic.json"""
import csv
import io
import json
import subprocess
import sys

rep, cells = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
kern = {}
seen = {}
for d in data:
    name = d[ik]
    if "k_xstream" in name:
        label = "k_xstream"
    elif "k_march3" in name:
        dirs = name.split("k_march3<")[1].split(",")
        label = "k_march3<%s>" % ("y" if dirs[2].strip().endswith("1") else "z")
    else:
        continue
    seen[label] = seen.get(label, 0) + 1
    b = float(d[ir]) * scale[units[ir]] + float(d[iw]) * scale[units[iw]]
    kern[f"{label}/stage{seen[label]}"] = {"dram_bytes_per_cell": b / cells, "dram_bytes_launch": b}
print(json.dumps({"source": f"{rep} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, {cells} cells, E=8)",
                  "cells": cells, "kernels": kern}, indent=1))
