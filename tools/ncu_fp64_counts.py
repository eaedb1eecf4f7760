#!/usr/bin/env python
"""Executed FP64 thread-instructions per cell and the issue-slot budget of the sweeps, from an
`ncu --set full` capture of one RK3 step's first two stages:
    python tools/ncu_fp64_counts.py gpurun_out/prof.ncu-rep CELLS > profiles/rNN_fp64_instruction_counts.json
Per launch: DFMA / DMUL / DADD thread-instructions per cell (smsp__sass_thread_inst_executed_op_d*_pred_on),
all warp-instructions per cell (smsp__inst_executed.sum), and the ISSUE model of DESIGN.md 4: an FP64 warp
instruction holds its scheduler's issue port for two cycles, so a launch needs at least
(2 x FP64 + other) warp-instructions / (4 schedulers x 148 SMs) cycles."""
import csv
import io
import json
import subprocess
import sys

rep, cells = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def num(d, key):
    return float(d[ix[key]].replace(",", ""))


res, seen = {}, {}
for d in data:
    name = d[ix["Kernel Name"]]
    if "k_xstream" in name:
        label = "k_xstream"
    elif "k_march3" in name:
        label = "k_march3<%s>" % ("y" if name.split("k_march3<")[1].split(",")[2].strip().endswith("1") else "z")
    else:
        continue
    seen[label] = seen.get(label, 0) + 1
    cyc = num(d, "smsp__cycles_elapsed.max")
    ops = {}
    for op in ("dfma", "dmul", "dadd"):
        ops[op] = num(d, f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed") * cyc
    fp64_thread = sum(ops.values())
    inst_warp = num(d, "smsp__inst_executed.sum")
    fp64_warp = fp64_thread / 32.0                      # full warps in the sweeps (predicated-off lanes are rare)
    other_warp = inst_warp - fp64_warp
    t_us = num(d, "gpu__time_duration.sum") * {"us": 1.0, "ms": 1e3, "ns": 1e-3}[units[ix["gpu__time_duration.sum"]]]
    sm_hz = num(d, "smsp__cycles_elapsed.max") / (t_us * 1e-6)
    slots = (2.0 * fp64_warp + other_warp) / (4 * 148)
    res[f"{label}/stage{seen[label]}"] = {
        "time_us": t_us,
        "fp64_thread_inst_per_cell": round(fp64_thread / cells, 1),
        "dfma": round(ops["dfma"] / cells, 1), "dmul": round(ops["dmul"] / cells, 1), "dadd": round(ops["dadd"] / cells, 1),
        "flops_per_cell": round((2 * ops["dfma"] + ops["dmul"] + ops["dadd"]) / cells, 1),
        "warp_inst_per_32_cells": round(inst_warp / cells * 32, 1),
        "fp64_warp_inst_per_32_cells": round(fp64_warp / cells * 32, 1),
        "other_warp_inst_per_32_cells": round(other_warp / cells * 32, 1),
        "issue_model_us": round(slots / sm_hz * 1e6, 1),
        "issue_model_over_measured": round(slots / sm_hz * 1e6 / t_us, 3),
        "fp64_pipe_active_pct": num(d, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": num(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    }
print(json.dumps({"source": f"{rep} (ncu --set full, {cells} cells, E = 8)", "kernels": res}, indent=1))
