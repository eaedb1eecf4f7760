"""ctypes wrapper of the CPU oracle (oracle/liborc_*.so).  TEST INFRASTRUCTURE: imported only
by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from microfc_b200.abi import Params, c_double_p, field_pointers
from microfc_b200.case import CaseConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
_libs = {}


def build():
    subprocess.run(["make", "-C", ODIR, "-j2"], check=True, capture_output=True)


def load(kind: str = "strict") -> C.CDLL:
    if kind in _libs:
        return _libs[kind]
    path = os.path.join(ODIR, f"liborc_{kind}.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    pp = C.POINTER(c_double_p)
    L.orc_create.restype = C.c_void_p
    L.orc_create.argtypes = [C.POINTER(Params), c_double_p, c_double_p, c_double_p, C.c_int, C.c_char_p, C.c_int]
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_set_q.argtypes = [C.c_void_p, pp]
    L.orc_get_q.argtypes = [C.c_void_p, pp]
    L.orc_get_prim.argtypes = [C.c_void_p, pp]
    L.orc_compute_rhs.argtypes = [C.c_void_p, C.c_int, pp]
    L.orc_step.argtypes = [C.c_void_p, C.c_int, C.c_double, c_double_p]
    L.orc_run_steps.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
    L.orc_run_steps.restype = C.c_double
    L.orc_decompose.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]
    L.orc_rank_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    L.orc_rank_metrics.argtypes = [C.c_void_p, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p]
    L.orc_weno_coefficients.argtypes = [C.c_void_p, C.c_int, C.c_int] + [c_double_p] * 5
    L.orc_rank_scratch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p]
    _libs[kind] = L
    return L


def global_params(cfg: CaseConfig) -> Params:
    """The GLOBAL case as the params struct (what rank 0 reads from simulation.inp)."""
    p = Params()
    p.abi_version = 1
    p.m, p.n, p.p = cfg.m, cfg.n, cfg.p
    p.m_glb, p.n_glb, p.p_glb = cfg.m, cfg.n, cfg.p
    p.num_dims, p.num_fluids, p.sys_size, p.buff_size = cfg.num_dims, cfg.num_fluids, cfg.sys_size, cfg.buff_size
    p.weno_order, p.weno_eps = cfg.weno_order, cfg.weno_eps
    p.time_stepper, p.weno_Re_flux, p.run_time_info = cfg.time_stepper, int(cfg.weno_Re_flux), int(cfg.run_time_info)
    p.t_step_start, p.t_step_stop = cfg.t_step_start, cfg.t_step_stop
    for d in range(3):
        p.bc[2 * d], p.bc[2 * d + 1] = cfg.bc[d][0], cfg.bc[d][1]
    p.proc_rank, p.num_procs = 0, 1
    for i in range(cfg.num_fluids):
        p.gammas[i], p.pi_infs[i] = cfg.gamma[i], cfg.pi_inf[i]
        p.Re[i][0], p.Re[i][1] = cfg.Re[i][0], cfg.Re[i][1]
    return p


class Oracle:
    """One emulated run of the reference `simulation` on num_procs ranks, global arrays in/out.
    Arrays have shape (sys_size, Nz, Ny, Nx)."""

    def __init__(self, cfg: CaseConfig, cb_glb, num_procs: int = 1, kind: str = "strict"):
        self.L = load(kind)
        self.cfg = cfg
        self.num_procs = num_procs
        self.gp = global_params(cfg)
        self._cb = [np.ascontiguousarray(c, dtype=np.float64) for c in cb_glb]
        ptr = [c.ctypes.data_as(c_double_p) for c in self._cb] + [None] * (3 - len(self._cb))
        err = C.create_string_buffer(256)
        self.h = self.L.orc_create(C.byref(self.gp), ptr[0], ptr[1], ptr[2], num_procs, err, 256)
        if not self.h:
            raise ValueError(err.value.decode())
        self.shape = (cfg.sys_size,) + cfg.shape_glb

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    def set_q(self, q: np.ndarray):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(self.shape)
        self.L.orc_set_q(self.h, field_pointers(list(q)))

    def _get(self, fn):
        out = np.empty(self.shape)
        fn(self.h, field_pointers(list(out)))
        return out

    def get_q(self):
        return self._get(self.L.orc_get_q)

    def get_prim(self):
        return self._get(self.L.orc_get_prim)

    def compute_rhs(self, t_step: int = 0):
        out = np.empty(self.shape)
        self.L.orc_compute_rhs(self.h, t_step, field_pointers(list(out)))
        return out

    def step(self, t_step: int, dt: float):
        stab = (C.c_double * 3)(float("nan"), float("nan"), float("nan"))
        self.L.orc_step(self.h, t_step, dt, stab)
        return list(stab)

    def run_steps(self, t_step0: int, n: int, dt: float) -> float:
        return self.L.orc_run_steps(self.h, t_step0, n, dt)

    def rank_info(self, rank: int):
        out = (C.c_int * 17)()
        self.L.orc_rank_info(self.h, rank, out)
        o = list(out)
        return dict(N=o[0:3], start_idx=o[3:6], bc=[o[6:8], o[8:10], o[10:12]], coords=o[12:15], buff_size=o[15], sys_size=o[16])

    def rank_metrics(self, rank: int, d: int):
        info = self.rank_info(rank)
        N, b = info["N"][d], info["buff_size"]
        cb, cc, ds = np.empty(N + 2 + 2 * b), np.empty(N + 1 + 2 * b), np.empty(N + 1 + 2 * b)
        self.L.orc_rank_metrics(self.h, rank, d, cb.ctypes.data_as(c_double_p), cc.ctypes.data_as(c_double_p), ds.ctypes.data_as(c_double_p))
        return cb, cc, ds

    def weno_coefficients(self, rank: int, d: int):
        info = self.rank_info(rank)
        nc = info["N"][d] + 1 + 2 * info["buff_size"] - 2 * self.cfg.weno_polyn
        pL, pR = np.empty((nc, 3, 2)), np.empty((nc, 3, 2))
        dL, dR, bt = np.empty((nc, 3)), np.empty((nc, 3)), np.empty((nc, 3, 3))
        n = self.L.orc_weno_coefficients(self.h, rank, d, *[a.ctypes.data_as(c_double_p) for a in (pL, pR, dL, dR, bt)])
        assert n == nc
        return dict(poly_L=pL, poly_R=pR, d_L=dL, d_R=dR, beta=bt)

    def rank_scratch(self, rank: int, which: int, d: int, v: int):
        info = self.rank_info(rank)
        b = info["buff_size"]
        ext = [info["N"][e] + 1 + 2 * b if e < self.cfg.num_dims else 1 for e in range(3)]
        out = np.empty((ext[2], ext[1], ext[0]))
        self.L.orc_rank_scratch(self.h, rank, which, d, v, out.ctypes.data_as(c_double_p))
        return out


def run_p_main(stepper, cfg: CaseConfig):
    """The reference's time loop (src/simulation/p_main.fpp:196-318) around any object with a
    .step(t_step, dt) method -- shared by the oracle and (through Simulation) the CUDA path so
    both see the same dt sequence, including the end-of-run dt tweak (:287) and the
    no-update last iteration (m_time_steppers.fpp:296).  Returns the per-step stability rows."""
    t_step = cfg.t_step_start
    dt = cfg.dt
    mytime = 0.0 if t_step == 0 else t_step * dt
    t_stop = cfg.t_step_stop
    finaltime = t_stop * dt
    rows = []
    while True:
        mytime = mytime + dt
        rows.append((t_step, dt, stepper.step(t_step, dt)))
        if t_step == t_stop:
            break
        if (mytime + dt) >= finaltime:
            dt = finaltime - mytime
        t_step += 1
    return rows
