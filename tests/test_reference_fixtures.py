"""Golden vectors produced by the REAL Fortran reference (tests/golden/ref_*.npz, written by
baseline/make_reference_fixtures.py from the restart files of an unmodified pre_process + simulation
run).  None can exist in this image (no Fortran toolchain: DESIGN.md 5, "parity unpinned"), so
these tests skip; the moment a maintainer with gfortran + MPI + fypp runs

    baseline/run_reference_fortran.sh /path/to/MicroFC examples/2D_advection/case.py 4 --fixtures

they pin the CPU oracle -- and on a GPU box the CUDA path -- against the reference itself at the
north-star tolerance.  The last test keeps the fixture pipeline itself alive: this repo's own
pre_process + oracle stand in for the reference executables and the loop must close."""
import dataclasses
import glob
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from microfc_b200 import data_io, pre_process
from microfc_b200.case import parse_case

from common import gpu_run, norm_linf, oracle_run, setup_case

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "ref_*.npz")))
TOL = 1e-10


def _load(path):
    z = np.load(path)
    cfg = parse_case(json.loads(str(z["case"])))
    cfg = dataclasses.replace(cfg, t_step_start=int(z["t0"]), t_step_stop=int(z["t1"]))
    cb = [z[f"cb{i}"] for i in range(cfg.num_dims)]
    return cfg, cb, z["q0"], z["q1"]


@pytest.mark.skipif(not FIXTURES, reason="no reference-produced fixtures (no Fortran toolchain in this image)")
@pytest.mark.parametrize("path", FIXTURES or [None])
def test_oracle_matches_the_fortran_reference(path):
    cfg, cb, q0, q1 = _load(path)
    q, _ = oracle_run(cfg, cb, q0)
    assert (norm_linf(q, q1, cfg) <= TOL).all(), norm_linf(q, q1, cfg)


@pytest.mark.gpu
@pytest.mark.skipif(not FIXTURES, reason="no reference-produced fixtures (no Fortran toolchain in this image)")
@pytest.mark.parametrize("path", FIXTURES or [None])
@pytest.mark.parametrize("strict", [True, False])
def test_cuda_matches_the_fortran_reference(path, strict):
    cfg, cb, q0, q1 = _load(path)
    q, _ = gpu_run(cfg, cb, q0, strict=strict)
    assert (norm_linf(q, q1, cfg) <= TOL).all(), norm_linf(q, q1, cfg)


def test_fixture_pipeline_closes_on_files_in_the_reference_formats(tmp_path):
    """pre_process files + a final restart file in the reference's formats -> make_reference_fixtures
    -> the loader above -> the oracle reproduces the final state (here bitwise: the stand-in for the
    reference IS the oracle)."""
    case_py = tmp_path / "case.py"
    from microfc_b200 import cases
    d = dict(cases.advection_2d(N=39, Nt=12), t_step_save=12)
    case_py.write_text("import json\nprint(json.dumps(%r))\n" % d)
    cfg, cb, q0 = setup_case(d)
    data_io.write_grid_parallel(str(tmp_path), cb)
    data_io.write_restart_parallel(str(tmp_path), 0, q0, cfg)
    q1, _ = oracle_run(cfg, cb, q0)
    data_io.write_restart_parallel(str(tmp_path), 12, q1, cfg)
    out = tmp_path / "golden"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "make_reference_fixtures.py"), str(case_py),
                        "--name", "selftest", "--out-dir", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    cfg2, cb2, a0, a1 = _load(str(out / "ref_selftest.npz"))
    assert np.array_equal(a0, q0) and np.array_equal(a1, q1) and cfg2.t_step_stop == 12
    q, _ = oracle_run(cfg2, cb2, a0)
    assert np.array_equal(q, q1)
