#!/usr/bin/env python
"""Static SASS opcode histogram of one kernel of libmfc_b200.so (no GPU needed):
    python tools/sass_mix.py <mangled-name-substring> [lib]"""
import collections
import re
import subprocess
import sys

lib = sys.argv[2] if len(sys.argv) > 2 else "microfc_b200/libmfc_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = out.split("Function : ")
for b in blocks[1:]:
    name = b.split("\n", 1)[0].strip()
    if sys.argv[1] not in name:
        continue
    ops = collections.Counter()
    for m in re.finditer(r"/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+(\.[A-Z0-9_]+)*)", b):
        op = m.group(2)
        base = op.split(".")[0]
        if base == "IMAD" and ("MOV" in op or "WIDE" in op):
            base = op
        ops[base] += 1
    tot = sum(ops.values())
    fp = ops["DFMA"] + ops["DMUL"] + ops["DADD"]
    print(name, "total", tot, "FP64", fp, f"{100*fp/tot:.0f}%")
    print("  " + "  ".join(f"{k}:{v}" for k, v in ops.most_common(26)))
