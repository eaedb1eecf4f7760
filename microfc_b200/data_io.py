"""The reference's on-disk formats, read and written natively, so that a case
directory round-trips with the Fortran ``pre_process`` / ``simulation`` / ``post_process``
executables: this driver can start from files ``pre_process`` wrote and leaves files
``post_process`` reads.  Citations are relative to the reference tree.

Parallel I/O (``parallel_io = T``), ``src/pre_process/m_data_output.f90:204-269``,
``src/simulation/m_data_output.fpp:477-540``, ``src/simulation/m_start_up.fpp:398-511``:
  ``restart_data/lustre_<t>.dat``   for variable i = 1..sys_size, at byte offset
                                    ``(m_glb+1)*max(1, n_glb+1)*8*(i-1)``, the GLOBAL array
                                    ``(0:m_glb, 0:n_glb)`` in Fortran order, raw native doubles,
                                    no header (every rank writes its sub-array through an MPI
                                    file view; here with seek + write)
  ``restart_data/lustre_x_cb.dat``  ``x_cb(-1:m_glb)``: m_glb + 2 raw doubles (``y_cb`` likewise)

Serial I/O (``parallel_io = F``), ``src/simulation/m_data_output.fpp:316-471``,
``src/simulation/m_start_up.fpp:298-396``, ``src/pre_process/m_data_output.f90:66-199``:
  ``p_all/p<rank>/<t>/x_cb.dat``        one gfortran sequential-unformatted record (4-byte
                                        length, payload, 4-byte length): ``x_cb(-1:m)``
  ``p_all/p<rank>/<t>/q_cons_vf<i>.dat`` one record: ``q_cons_vf(i)%sf(0:m, 0:n)``
  ``D/cons.<i>.<rank:02>.<t:06>.dat``   ASCII, ``(2F40.14)`` rows ``x_cb(j), q`` in 1-D;
                                        ``(3F40.14)`` rows ``x_cb(j), y_cb(k), q`` with a blank
                                        line after every j in 2-D
  ``D/prim.<i>...``                     1-D only: alpha_rho/alpha as stored, u = mom/rho,
                                        p = (E - 0.5 mom**2/rho - pi_inf)/gamma

Run-time files: ``run_time.inf`` (``m_data_output.fpp:90-146,283-294``), ``time_data.dat``
(``p_main.fpp:261-270``).

The extension to 3-D (``p > 0``) keeps the same rules with one more array dimension
(``z_cb.dat``, global array ``(0:m_glb, 0:n_glb, 0:p_glb)``); the reference has no 3-D.
Arrays in this module are ``(E, Nz, Ny, Nx)`` C-ordered with x fastest, which is byte for
byte the Fortran order ``(x, y[, z])`` of one variable after another.
"""
from __future__ import annotations

import datetime
import os
import struct
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .case import CaseConfig

MPIIOFS = "lustre_"          # m_global_parameters.fpp:413


# ---- parallel (MPI-IO layout) --------------------------------------------------------------
def _restart_path(case_dir: str, name: str) -> str:
    return os.path.join(case_dir, "restart_data", MPIIOFS + name)


def write_grid_parallel(case_dir: str, cb_glb: Sequence[np.ndarray]) -> None:
    """``cb_glb[d]`` = global cell boundaries ``s_cb(-1:N_glb)`` (N_glb + 2 values)."""
    os.makedirs(os.path.join(case_dir, "restart_data"), exist_ok=True)
    for d, name in enumerate(("x_cb.dat", "y_cb.dat", "z_cb.dat")[:len(cb_glb)]):
        np.ascontiguousarray(cb_glb[d], dtype=np.float64).tofile(_restart_path(case_dir, name))


def read_grid_parallel(case_dir: str, cfg: CaseConfig) -> List[np.ndarray]:
    out = []
    for d, name in enumerate(("x_cb.dat", "y_cb.dat", "z_cb.dat")[:cfg.num_dims]):
        path = _restart_path(case_dir, name)
        if not os.path.exists(path):
            raise FileNotFoundError(f"File {path} is missing. Exiting...")      # m_start_up.fpp:427-430
        a = np.fromfile(path, dtype=np.float64)
        n = cfg.shape_glb[::-1][d] + 1
        if a.size != n:
            raise ValueError(f"{path}: expected {n} doubles (N_glb + 2), found {a.size}")
        out.append(a)
    return out


def write_restart_parallel(case_dir: str, t_step: int, q_block: np.ndarray, cfg: CaseConfig,
                           interior: Optional[Tuple[slice, slice, slice]] = None) -> None:
    """Write this rank's block ``q_block`` (E, nz, ny, nx) of the global arrays into
    ``restart_data/lustre_<t>.dat``.  ``interior`` = the block's (z, y, x) slices in the global
    array (None: the block IS the global array).  Ranks may call this concurrently: every rank
    only touches its own byte ranges, exactly like the MPI file view of the reference."""
    os.makedirs(os.path.join(case_dir, "restart_data"), exist_ok=True)
    path = _restart_path(case_dir, f"{t_step}.dat")
    Nz, Ny, Nx = cfg.shape_glb
    E = q_block.shape[0]
    var_bytes = Nx * Ny * Nz * 8
    if interior is None:
        interior = (slice(0, Nz), slice(0, Ny), slice(0, Nx))
    zs, ys, xs = interior
    q_block = np.ascontiguousarray(q_block, dtype=np.float64)
    assert q_block.shape[1:] == (zs.stop - zs.start, ys.stop - ys.start, xs.stop - xs.start)
    fd = os.open(path, os.O_WRONLY | os.O_CREAT, 0o644)
    try:
        # the reference deletes an existing file first (MPI_FILE_DELETE, m_data_output.fpp:506-509);
        # every rank computes the same size, so setting it exactly is race-free and drops the
        # trailing bytes of a stale, larger file
        if os.fstat(fd).st_size != E * var_bytes:
            os.ftruncate(fd, E * var_bytes)
        for v in range(E):
            for iz, z in enumerate(range(zs.start, zs.stop)):
                if xs.start == 0 and xs.stop == Nx:            # whole rows: one write per plane
                    off = v * var_bytes + ((z * Ny + ys.start) * Nx) * 8
                    os.pwrite(fd, q_block[v, iz].tobytes(), off)
                    continue
                for iy, y in enumerate(range(ys.start, ys.stop)):
                    off = v * var_bytes + ((z * Ny + y) * Nx + xs.start) * 8
                    os.pwrite(fd, q_block[v, iz, iy].tobytes(), off)
    finally:
        os.close(fd)


def read_restart_parallel(case_dir: str, t_step: int, cfg: CaseConfig,
                          interior: Optional[Tuple[slice, slice, slice]] = None) -> np.ndarray:
    """This rank's block (E, nz, ny, nx) of ``restart_data/lustre_<t>.dat``."""
    path = _restart_path(case_dir, f"{t_step}.dat")
    if not os.path.exists(path):
        raise FileNotFoundError(f"File {path} is missing. Exiting...")          # m_start_up.fpp:503-506
    Nz, Ny, Nx = cfg.shape_glb
    E = cfg.sys_size
    if os.path.getsize(path) != E * Nx * Ny * Nz * 8:
        raise ValueError(f"{path}: size does not match sys_size*(m_glb+1)*(n_glb+1)*8")
    mm = np.memmap(path, dtype=np.float64, mode="r", shape=(E, Nz, Ny, Nx))
    if interior is None:
        return np.array(mm)
    zs, ys, xs = interior
    return np.array(mm[:, zs, ys, xs])


# ---- serial (gfortran sequential unformatted) ------------------------------------------------
def _write_record(path: str, a: np.ndarray) -> None:
    payload = np.ascontiguousarray(a, dtype=np.float64).tobytes()
    if len(payload) >= 2 ** 31:
        raise ValueError("records of 2 GiB and more use gfortran's sub-record scheme; use parallel_io")
    mark = struct.pack("<i", len(payload))
    with open(path, "wb") as f:
        f.write(mark + payload + mark)


def _read_record(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        raw = f.read()
    (n,) = struct.unpack("<i", raw[:4])
    if n < 0 or len(raw) != n + 8 or struct.unpack("<i", raw[-4:])[0] != n:
        raise ValueError(f"{path}: not a single gfortran sequential-unformatted record")
    return np.frombuffer(raw[4:4 + n], dtype=np.float64).copy()


def _step_dir(case_dir: str, rank: int, t_step: int) -> str:
    return os.path.join(case_dir, "p_all", f"p{rank}", str(t_step))


def write_serial(case_dir: str, rank: int, t_step: int, cb_loc: Sequence[np.ndarray], q_loc: np.ndarray) -> None:
    """``cb_loc[d]`` = this rank's ``s_cb(-1:N)``; ``q_loc`` = (E, nz, ny, nx) interior cells."""
    d = _step_dir(case_dir, rank, t_step)
    os.makedirs(d, exist_ok=True)
    for i, name in enumerate(("x_cb.dat", "y_cb.dat", "z_cb.dat")[:len(cb_loc)]):
        _write_record(os.path.join(d, name), cb_loc[i])
    for v in range(q_loc.shape[0]):
        _write_record(os.path.join(d, f"q_cons_vf{v + 1}.dat"), q_loc[v])


def read_serial(case_dir: str, rank: int, t_step: int, cfg: CaseConfig, local_shape: Tuple[int, int, int]):
    """-> (cb_loc, q_loc) of ``p_all/p<rank>/<t>``; ``local_shape`` = (nz, ny, nx)."""
    d = _step_dir(case_dir, rank, t_step)
    nz, ny, nx = local_shape
    dims = (nx, ny, nz)
    cb = []
    for i, name in enumerate(("x_cb.dat", "y_cb.dat", "z_cb.dat")[:cfg.num_dims]):
        a = _read_record(os.path.join(d, name))
        if a.size != dims[i] + 1:
            raise ValueError(f"{name}: expected {dims[i] + 1} values")
        cb.append(a)
    q = np.empty((cfg.sys_size, nz, ny, nx))
    for v in range(cfg.sys_size):
        a = _read_record(os.path.join(d, f"q_cons_vf{v + 1}.dat"))
        if a.size != nx * ny * nz:
            raise ValueError(f"q_cons_vf{v + 1}.dat: expected {nx * ny * nz} values")
        q[v] = a.reshape(nz, ny, nx)
    return cb, q


def _f40(*vals: float) -> str:
    return "".join(f"{v:40.14f}" for v in vals) + "\n"


def write_ascii(case_dir: str, rank: int, t_step: int, cb_loc: Sequence[np.ndarray], q_loc: np.ndarray,
                cfg: CaseConfig) -> None:
    """``D/cons.*`` (and ``D/prim.*`` in 1-D) of s_write_serial_data_files,
    m_data_output.fpp:377-462 (``precision = 2``: F40.14)."""
    D = os.path.join(case_dir, "D")
    os.makedirs(D, exist_ok=True)
    E = q_loc.shape[0]
    x = cb_loc[0][1:]                       # x_cb(0:m)
    tag = f"{rank:02d}.{t_step:06d}.dat"
    if cfg.num_dims == 1:
        nf = cfg.num_fluids
        rho = q_loc[:nf, 0, 0].sum(axis=0)
        gamma = sum(q_loc[nf + 2 + i, 0, 0] * cfg.gamma[i] for i in range(nf))
        pi_inf = sum(q_loc[nf + 2 + i, 0, 0] * cfg.pi_inf[i] for i in range(nf))
        for v in range(E):
            if v < nf or v > nf + 1:
                vals = q_loc[v, 0, 0]
            elif v == nf:
                vals = q_loc[nf, 0, 0] / rho
            else:
                vals = (q_loc[nf + 1, 0, 0] - 0.5 * q_loc[nf, 0, 0] ** 2.0 / rho - pi_inf) / gamma
            with open(os.path.join(D, f"prim.{v + 1}.{tag}"), "w") as f:
                f.writelines(_f40(x[j], vals[j]) for j in range(x.size))
            with open(os.path.join(D, f"cons.{v + 1}.{tag}"), "w") as f:
                f.writelines(_f40(x[j], q_loc[v, 0, 0, j]) for j in range(x.size))
    elif cfg.num_dims == 2:
        y = cb_loc[1][1:]
        for v in range(E):
            with open(os.path.join(D, f"cons.{v + 1}.{tag}"), "w") as f:
                for j in range(x.size):
                    f.writelines(_f40(x[j], y[k], q_loc[v, 0, k, j]) for k in range(y.size))
                    f.write("\n")
    # 3-D: the reference has none; the extension writes no ASCII dump


# ---- run-time information ----------------------------------------------------------------------
class RunTimeInfo:
    """``run_time.inf``: header of s_open_run_time_information_file (m_data_output.fpp:90-146)
    and one row per step of s_write_run_time_information (:283-294), rank 0 only."""

    def __init__(self, case_dir: str, viscous: bool):
        self.path = os.path.join(case_dir, "run_time.inf")
        self.viscous = viscous
        exist = os.path.exists(self.path)
        self.f = open(self.path, "a")
        if not exist:
            self.f.write("Description: Stability information at each time-step of the simulation. This\n")
            for line in ("data is composed of the inviscid Courant–Friedrichs–Lewy (ICFL)",
                         "number and the cell Reynolds (Rc) number. Please note that only",
                         "those stability conditions pertinent to the physics included in",
                         "the current computation are displayed."):
                self.f.write(" " * 13 + line + "\n")
            self.f.write("Date: " + datetime.date.today().strftime("%m/%d/%y") + "\n")
        self.f.write("\n\n")
        if viscous:
            self.f.write("==== Time-steps ====== Time ======= ICFL Max ==== VCFL Max ====== Rc Min =======\n")
        else:
            self.f.write("=========== Time-steps ============== Time ============== ICFL Max =============\n")

    @staticmethod
    def _f(v: float, w: int, d: int) -> str:
        s = f"{v:{w}.{d}f}"
        return s if len(s) <= w else "*" * w          # Fortran overflow rule

    def row(self, t_step: int, dt: float, stab: Sequence[float]) -> None:
        t = t_step * dt
        if self.viscous:
            self.f.write(" " * 6 + f"{t_step:8d}" + " " * 6 + self._f(t, 10, 6) + " " * 6 + self._f(stab[0], 9, 6)
                         + " " * 6 + self._f(stab[1], 9, 6) + " " * 6 + self._f(stab[2], 10, 6) + "\n")
        else:
            self.f.write(" " * 13 + f"{t_step:8d}" + " " * 14 + self._f(t, 10, 6) + " " * 13 + self._f(stab[0], 9, 6) + "\n")

    def close(self) -> None:
        self.f.close()


def append_time_data(case_dir: str, num_procs: int, time_final: float, name: str = "time_data.dat") -> None:
    """``write (1, *) num_procs, time_final`` (p_main.fpp:261-270); list-directed output is
    compiler-defined, this is gfortran's integer(4) width and a 17-digit real."""
    with open(os.path.join(case_dir, name), "a") as f:
        f.write(f"{num_procs:12d}{time_final:26.16E}\n")
