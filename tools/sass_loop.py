#!/usr/bin/env python
"""Opcode mix of the largest loop (backward branch with the widest span) of one kernel:
    python tools/sass_loop.py <mangled-name-substring> [lib]"""
import collections
import re
import subprocess
import sys

lib = sys.argv[2] if len(sys.argv) > 2 else "microfc_b200/libmfc_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
for b in out.split("Function : ")[1:]:
    name = b.split("\n", 1)[0].strip()
    if sys.argv[1] not in name:
        continue
    ins = []
    for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)([^;]*);", b):
        ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
    best = (0, 0, 0)
    for addr, op, rest in ins:
        if op.startswith("BRA"):
            t = re.search(r"0x([0-9a-f]+)", rest)
            if t and int(t.group(1), 16) < addr and addr - int(t.group(1), 16) > best[0]:
                best = (addr - int(t.group(1), 16), int(t.group(1), 16), addr)
    lo, hi = best[1], best[2]
    ops = collections.Counter()
    for addr, op, rest in ins:
        if lo <= addr <= hi:
            base = op.split(".")[0]
            if base == "IMAD" and ("MOV" in op or "WIDE" in op):
                base = op
            ops[base] += 1
    tot = sum(ops.values())
    fp = ops["DFMA"] + ops["DMUL"] + ops["DADD"]
    print(name, f"loop 0x{lo:x}..0x{hi:x}: {tot} instructions, FP64 {fp} ({100*fp/tot:.0f}%), non-FP64 {tot-fp}")
    print("  " + "  ".join(f"{k}:{v}" for k, v in ops.most_common(30)))
