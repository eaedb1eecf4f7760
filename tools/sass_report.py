#!/usr/bin/env python
"""Static SASS evidence of the hot kernels of libmfc_b200.so (no GPU needed):
    python tools/sass_report.py > profiles/r02_sass_opcodes.txt
Per kernel: registers are in the build log (-Xptxas -v); here the opcode counts that prove the
design claims -- TMA loads / stores (UTMALDG / UTMASTG), mbarriers (SYNCS), no local-memory spills
(LDL / STL), the FP64 mix (DFMA / DMUL / DADD), shared-memory and global traffic instructions --
for the whole kernel and for its main loop (the widest backward branch)."""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "microfc_b200/libmfc_b200.so"
COLS = ["UTMALDG", "UTMASTG", "SYNCS", "LDL", "STL", "DFMA", "DMUL", "DADD", "MUFU", "LDS", "STS", "LDG", "STG", "SHFL", "BAR"]
# (label, substring of the mangled name)
KERNELS = [
    ("k_xstream  3-D 2-fluid uniform (512^3 workload)", "fast9k_xstreamILi2ELi3ELi0ELb0ELi0ELi5"),
    ("k_march3<y> 3-D 2-fluid uniform", "fast8k_march3ILi2ELi3ELi1ELi0ELb0ELi0ELi5"),
    ("k_march3<z> 3-D 2-fluid uniform (+ fused RK)", "fast8k_march3ILi2ELi3ELi2ELi0ELb0ELi0ELi5"),
    ("k_xstream  2-D 2-fluid uniform, bc -4 (shock-bubble)", "fast9k_xstreamILi2ELi2ELi0ELb1ELi0ELi5"),
    ("k_march3<y> 2-D 2-fluid uniform, bc -4 (+ fused RK)", "fast8k_march3ILi2ELi2ELi1ELi0ELb1ELi0ELi5"),
    ("k_xstream  2-D viscous in-sweep (configs[3])", "fast9k_xstreamILi2ELi2ELi0ELb1ELi2ELi5"),
    ("k_march3<y> 2-D viscous in-sweep (+ fused RK)", "fast8k_march3ILi2ELi2ELi1ELi0ELb1ELi2ELi5"),
    ("k_xstream  3-D stretched grid (coefficient tables)", "fast9k_xstreamILi2ELi3ELi1ELb0ELi0ELi5"),
    ("k_march3<y> 3-D stretched grid", "fast8k_march3ILi2ELi3ELi1ELi1ELb0ELi0ELi5"),
    ("k_march3<z> 3-D stretched grid", "fast8k_march3ILi2ELi3ELi2ELi1ELb0ELi0ELi5"),
    ("k_xstream  1-D 1-fluid (Sod, + fused RK)", "fast9k_xstreamILi1ELi1ELi0ELb0ELi0ELi5"),
    ("k_xstream  strict build 3-D", "strict9k_xstreamILi2ELi3ELi1ELb0ELi0ELi5"),
    ("k_march3<z> strict build 3-D", "strict8k_march3ILi2ELi3ELi2ELi1ELb0ELi0ELi5"),
]


def count(ins):
    c = collections.Counter()
    for _, op in ins:
        c[op.split(".")[0]] += 1
    return c


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    blocks = {b.split("\n", 1)[0].strip(): b for b in out.split("Function : ")[1:]}
    print(f"# {LIB}: SASS opcode counts (cuobjdump -sass), whole kernel / main loop")
    print("# " + " ".join(f"{c:>8s}" for c in ["instr", "FP64%"] + COLS))
    for label, key in KERNELS:
        names = [n for n in blocks if key in n]
        if not names:
            print(f"{label}: NOT FOUND ({key})")
            continue
        b = blocks[names[0]]
        ins = [(int(m.group(1), 16), m.group(3)) for m in
               re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)([^;]*);", b)]
        # main loop = the backward branch with the widest span
        best = (0, 0, 0)
        for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d+\s+)?BRA[^;]*?0x([0-9a-f]+)", b):
            a, t = int(m.group(1), 16), int(m.group(3), 16)
            if t < a and a - t > best[0]:
                best = (a - t, t, a)
        print(f"{label}\n   {names[0]}")
        for scope, sel in (("kernel", ins), ("loop", [i for i in ins if best[1] <= i[0] <= best[2]])):
            c = count(sel)
            tot = len(sel)
            fp = c["DFMA"] + c["DMUL"] + c["DADD"]
            print(f"   {scope:6s}" + " ".join(f"{v:8d}" for v in [tot]) + f" {100.0*fp/max(tot,1):7.1f}% " + " ".join(f"{c[k]:8d}" for k in COLS))


if __name__ == "__main__":
    main()
