"""End to end through the command-line driver on a GPU: an unchanged-format case script ->
pre_process files -> CUDA simulation -> restart files / run_time.inf in the reference's formats,
compared with the CPU oracle run on the same case."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from microfc_b200 import cases, data_io

from common import oracle_run, setup_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("parallel_io", ["T", "F"])
def test_run_case_script_matches_oracle_bitwise_in_strict_mode(tmp_path, parallel_io):
    dct = cases.shockbubble_2d(Ny=30)
    dct.update({"t_step_stop": 12, "t_step_save": 6, "parallel_io": parallel_io})
    (tmp_path / "case.py").write_text("import json\nprint(json.dumps(" + json.dumps(dct) + "))\n")
    r = subprocess.run([sys.executable, "-m", "microfc_b200", "run", str(tmp_path / "case.py"), "--strict"],
                       cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    cfg, cb, q0 = setup_case(dct)
    q_ref, rows = oracle_run(cfg, cb, q0)
    if parallel_io == "T":
        q = data_io.read_restart_parallel(str(tmp_path), 12, cfg)
        assert os.path.exists(tmp_path / "restart_data" / "lustre_6.dat")
    else:
        _, q = data_io.read_serial(str(tmp_path), 0, 12, cfg, cfg.shape_glb)
        assert os.path.exists(tmp_path / "D" / "cons.1.00.000006.dat")
    assert np.array_equal(q, q_ref)
    lines = open(tmp_path / "run_time.inf").read().splitlines()
    assert len([l for l in lines if l.startswith(" " * 13)]) >= 13 + 4       # 4 header lines + one row per step
    assert float(lines[-1].split()[-1]) == pytest.approx(rows[-1][2][0], abs=1e-6)
