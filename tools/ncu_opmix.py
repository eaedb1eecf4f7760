#!/usr/bin/env python
"""Dynamic SASS opcode mix + stall samples of one kernel in an .ncu-rep (needs --import-source on):
    python tools/ncu_opmix.py prof.ncu-rep <kernel-regex> [launch-skip]"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    print(rows[0][1][:120])
    hdr = rows[1]
    iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    ops, samp, stalls, tot = collections.Counter(), collections.Counter(), collections.Counter(), 0
    for r in rows[2:]:
        if len(r) <= iE or not r[iE].isdigit():
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)((\.[A-Z0-9_]+)*)", r[iS])
        op = m.group(2)
        if op == "IMAD" and ("MOV" in m.group(3) or "WIDE" in m.group(3)):
            op += m.group(3)
        n = int(r[iE])
        ops[op] += n
        tot += n
        samp[op] += int(r[iN])
        for i in stall_cols:
            stalls[hdr[i]] += int(r[i] or 0)
    print("total warp instructions", tot)
    fp64 = sum(ops[o] for o in ("DFMA", "DMUL", "DADD"))
    print(f"FP64 (DFMA+DMUL+DADD) {fp64} = {100*fp64/tot:.1f}%")
    for op, n in ops.most_common(28):
        print(f"{op:16s} {n:12d} {100*n/tot:5.1f}%  samples {samp[op]}")
    print(stalls.most_common(10))


if __name__ == "__main__":
    main()
