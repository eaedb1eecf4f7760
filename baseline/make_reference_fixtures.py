#!/usr/bin/env python
"""Turn the output of a REAL run of the Fortran reference into golden vectors for this repo.

    python baseline/make_reference_fixtures.py <case.py> [--case-dir DIR] [--name NAME] [--max-cells N]

Reads what the unmodified `pre_process` + `simulation` executables wrote for the case
(`restart_data/lustre_<t>.dat`, `lustre_{x,y}_cb.dat`; parallel_io = T, formats in
microfc_b200/data_io.py citing m_data_output.fpp:477-540) at t_step_start and at the last saved step,
and stores them as tests/golden/ref_<name>.npz together with the case dictionary.
tests/test_reference_fixtures.py then pins the CPU oracle (and, on a GPU box, the CUDA path) against
the reference ITSELF; without such files that test skips and parity stays "unpinned" (DESIGN.md 5).

This image has no Fortran toolchain, so the script cannot be exercised against real output here; its
reading side is the same code the CLI uses to start from files written by the Fortran pre_process
(tests/test_data_io.py, tests/test_cli_gpu.py), and tests/test_reference_fixtures.py checks the whole
loop on files written by this repo's own pre_process + oracle standing in for the reference."""
import argparse
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from microfc_b200 import data_io  # noqa: E402
from microfc_b200.case import parse_case  # noqa: E402


def make(case_py: str, case_dir: str, name: str, max_cells: int, out_dir: str) -> str:
    d = json.loads(subprocess.run([sys.executable, case_py], capture_output=True, text=True, check=True).stdout)
    cfg = parse_case(d)
    cells = int(np.prod(cfg.shape_glb))
    if cells > max_cells:
        raise SystemExit(f"{cells} cells: too large for a committed fixture (--max-cells {max_cells}); run the case at a smaller m, n")
    if not cfg.parallel_io:
        raise SystemExit("fixtures are read from the parallel_io = T restart files")
    saved = [t for t in range(cfg.t_step_start, cfg.t_step_stop + 1, max(1, cfg.t_step_save))
             if os.path.exists(os.path.join(case_dir, "restart_data", f"lustre_{t}.dat"))]
    if len(saved) < 2:
        raise SystemExit(f"need restart files of at least two steps under {case_dir}/restart_data, found {saved}")
    cb = data_io.read_grid_parallel(case_dir, cfg)
    q0 = data_io.read_restart_parallel(case_dir, saved[0], cfg)
    q1 = data_io.read_restart_parallel(case_dir, saved[-1], cfg)
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, f"ref_{name}.npz")
    np.savez_compressed(path, case=json.dumps(d), t0=saved[0], t1=saved[-1], q0=q0, q1=q1,
                        **{f"cb{i}": c for i, c in enumerate(cb)})
    return path


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("case")
    ap.add_argument("--case-dir", default=None)
    ap.add_argument("--name", default=None)
    ap.add_argument("--max-cells", type=int, default=200000)
    ap.add_argument("--out-dir", default=os.path.join(ROOT, "tests", "golden"))
    a = ap.parse_args()
    cd = a.case_dir or os.path.dirname(os.path.abspath(a.case))
    nm = a.name or os.path.basename(cd.rstrip("/"))
    print(make(a.case, cd, nm, a.max_cells, a.out_dir))
