#!/usr/bin/env python
"""Per-instruction stall samples of the kernels in an ncu source-page CSV:
    ncu -i rep --page source --csv --print-source sass > f.csv
    python tools/ncu_stalls.py f.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) > 10:
        cur["data"].append(r)
for sec in sections:
    hdr, data = sec["hdr"], sec["data"]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    val = lambda r, h: int(r[ix[h]] or 0) if ix[h] < len(r) else 0
    tot = {s: sum(val(r, s) for r in data) for s in stalls}
    S = sum(val(r, "# Samples") for r in data)
    print("=====", sec["name"][:90], "samples", S)
    print("  ".join(f"{s[6:]}:{100*v/S:.1f}%" for s, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v))
    order = sorted(range(len(data)), key=lambda i: -val(data[i], "# Samples"))[:top]
    for i in sorted(order):
        r = data[i]
        st = {s[6:]: val(r, s) for s in stalls if val(r, s)}
        print(f"{i:5d} {val(r, '# Samples'):6d}  {r[ix['Source']].strip():58s} {st}")
