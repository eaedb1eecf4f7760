"""The reference's file formats, microfc_b200/data_io.py: byte-level layout
against hand-computed expectations, and round trips."""
import dataclasses
import os
import struct

import numpy as np

from microfc_b200 import cases, data_io, pre_process
from microfc_b200.domain import rank_layout

from common import setup_case


def test_parallel_restart_layout_and_round_trip(tmp_path):
    cfg, cb, q0 = setup_case(cases.shockbubble_2d(Ny=30), n_steps=1)
    d = str(tmp_path)
    data_io.write_grid_parallel(d, cb)
    data_io.write_restart_parallel(d, 0, q0, cfg)
    path = os.path.join(d, "restart_data", "lustre_0.dat")
    Nz, Ny, Nx = cfg.shape_glb
    assert os.path.getsize(path) == cfg.sys_size * Nx * Ny * 8            # no header
    raw = np.fromfile(path)
    # variable i at offset (m_glb+1)*(n_glb+1)*8*(i-1), Fortran order (x fastest)
    for i in range(cfg.sys_size):
        blk = raw[i * Nx * Ny:(i + 1) * Nx * Ny].reshape(Ny, Nx)
        assert np.array_equal(blk, q0[i, 0])
    assert np.fromfile(os.path.join(d, "restart_data", "lustre_x_cb.dat")).size == cfg.m + 2
    back = data_io.read_restart_parallel(d, 0, cfg)
    assert np.array_equal(back, q0)
    assert all(np.array_equal(a, b) for a, b in zip(data_io.read_grid_parallel(d, cfg), cb))


def test_parallel_restart_replaces_a_stale_larger_file(tmp_path):
    """The reference deletes an existing restart file before writing (MPI_FILE_DELETE,
    m_data_output.fpp:506-509): a file left by an earlier, larger run must not keep its length."""
    big = setup_case(cases.shockbubble_2d(Ny=40), n_steps=1)
    cfg, cb, q0 = setup_case(cases.shockbubble_2d(Ny=30), n_steps=1)
    d = str(tmp_path)
    data_io.write_restart_parallel(d, 0, big[2], big[0])
    data_io.write_restart_parallel(d, 0, q0, cfg)
    Nz, Ny, Nx = cfg.shape_glb
    assert os.path.getsize(os.path.join(d, "restart_data", "lustre_0.dat")) == cfg.sys_size * Nx * Ny * 8
    assert np.array_equal(data_io.read_restart_parallel(d, 0, cfg), q0)


def test_parallel_restart_written_by_ranks_equals_single_writer(tmp_path):
    cfg, cb, q0 = setup_case(cases.shockbubble_2d_cells(120, 64), n_steps=1)
    a, b = str(tmp_path / "one"), str(tmp_path / "four")
    data_io.write_restart_parallel(a, 7, q0, cfg)
    for r in range(4):
        lay = rank_layout(r, 4, cfg)
        sl = lay.interior_slices()
        data_io.write_restart_parallel(b, 7, q0[(slice(None),) + sl], cfg, sl)
        got = data_io.read_restart_parallel(a, 7, cfg, sl)
        assert np.array_equal(got, q0[(slice(None),) + sl])
    assert open(os.path.join(a, "restart_data", "lustre_7.dat"), "rb").read() == \
        open(os.path.join(b, "restart_data", "lustre_7.dat"), "rb").read()


def test_serial_files_are_single_fortran_records(tmp_path):
    cfg, cb, q0 = setup_case(cases.advection_2d(N=31), n_steps=1)
    d = str(tmp_path)
    data_io.write_serial(d, 3, 5, cb, q0)
    p = os.path.join(d, "p_all", "p3", "5")
    raw = open(os.path.join(p, "x_cb.dat"), "rb").read()
    n = (cfg.m + 2) * 8
    assert struct.unpack("<i", raw[:4])[0] == n and struct.unpack("<i", raw[-4:])[0] == n and len(raw) == n + 8
    raw = open(os.path.join(p, "q_cons_vf2.dat"), "rb").read()
    n = (cfg.m + 1) * (cfg.n + 1) * 8
    assert struct.unpack("<i", raw[:4])[0] == n and len(raw) == n + 8
    assert np.array_equal(np.frombuffer(raw[4:-4]).reshape(cfg.n + 1, cfg.m + 1), q0[1, 0])   # sf(0:m,0:n), x fastest
    cb2, q2 = data_io.read_serial(d, 3, 5, cfg, (1, cfg.n + 1, cfg.m + 1))
    assert np.array_equal(q2, q0) and all(np.array_equal(x, y) for x, y in zip(cb2, cb))


def test_ascii_dump_format(tmp_path):
    cfg, cb, q0 = setup_case(cases.sod_1d(), n_steps=1)
    d = str(tmp_path)
    data_io.write_ascii(d, 0, 12, cb, q0, cfg)
    lines = open(os.path.join(d, "D", "cons.1.00.000012.dat")).read().splitlines()
    assert len(lines) == cfg.m + 1 and all(len(l) == 80 for l in lines)                     # (2F40.14)
    assert lines[0] == f"{cb[0][1]:40.14f}{q0[0, 0, 0, 0]:40.14f}"
    # prim.3 = pressure = (E - 0.5 mom^2/rho - pi_inf)/gamma -> 1 | 0.1 for the Sod tube
    pres = np.array([float(l[40:]) for l in open(os.path.join(d, "D", "prim.3.00.000012.dat")).read().splitlines()])
    assert abs(pres[0] - 1.0) < 1e-12 and abs(pres[-1] - 0.1) < 1e-12
    cfg2, cb2, q2 = setup_case(cases.advection_2d(N=24), n_steps=1)
    data_io.write_ascii(d, 1, 0, cb2, q2, cfg2)
    txt = open(os.path.join(d, "D", "cons.4.01.000000.dat")).read().split("\n")
    # for j: (n+1) rows of (3F40.14), then a blank line
    assert len(txt[0]) == 120 and txt[cfg2.n + 1] == "" and len(txt) == (cfg2.m + 1) * (cfg2.n + 2) + 1
    assert txt[1] == f"{cb2[0][1]:40.14f}{cb2[1][2]:40.14f}{q2[3, 0, 1, 0]:40.14f}"


def test_run_time_inf_rows(tmp_path):
    d = str(tmp_path)
    r = data_io.RunTimeInfo(d, viscous=False)
    r.row(12, 1e-4, [0.123456789])
    r.close()
    lines = open(os.path.join(d, "run_time.inf")).read().splitlines()
    assert lines[0].startswith("Description: Stability information")
    assert lines[-2] == "=========== Time-steps ============== Time ============== ICFL Max ============="
    # '(13X,I8,14X,F10.6,13X,F9.6)'
    assert lines[-1] == " " * 13 + "      12" + " " * 14 + "  0.001200" + " " * 13 + " 0.123457"
    r = data_io.RunTimeInfo(d, viscous=True)           # appended, header not repeated
    r.row(3, 0.5, [0.1, 0.2, 12345.0])
    r.close()
    lines = open(os.path.join(d, "run_time.inf")).read().splitlines()
    assert sum(l.startswith("Description") for l in lines) == 1
    assert lines[-1] == " " * 6 + "       3" + " " * 6 + "  1.500000" + " " * 6 + " 0.100000" + " " * 6 + " 0.200000" + " " * 6 + "**********"


def test_cli_pre_process_writes_what_simulation_reads(tmp_path):
    """python -m microfc_b200 pre_process on an unchanged-format case script."""
    import json
    import subprocess
    import sys
    case = tmp_path / "case.py"
    dct = cases.shockbubble_2d(Ny=30)
    case.write_text("import json\nprint(json.dumps(" + json.dumps(dct) + "))\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([sys.executable, "-m", "microfc_b200", "pre_process", str(case)], check=True, cwd=root)
    cfg = cases.config(dct)
    cb = pre_process.generate_grid(cfg)
    q0 = pre_process.generate_initial_condition(cfg, cb)
    if cfg.parallel_io:
        assert np.array_equal(data_io.read_restart_parallel(str(tmp_path), 0, cfg), q0)
    else:
        _, q = data_io.read_serial(str(tmp_path), 0, 0, cfg, cfg.shape_glb)
        assert np.array_equal(q, q0)
