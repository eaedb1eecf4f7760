"""Host driver: plays the reference's ``simulation`` executable around libmfc_b200.so.

The Fortran host (``src/simulation/p_main.fpp``) stays unchanged in a real deployment and
calls the C ABI through ``fortran/m_b200_bindings.f90``; no Fortran compiler exists in this
image, so this module reproduces the host side in Python, call for call:

=====================================  =====================================================
p_main.fpp                             here
=====================================  =====================================================
:118-120 bcast inputs, decompose       :func:`microfc_b200.domain.rank_layout`
:125-151 module initialisers           ``mfc_b200_init`` (+ ``mfc_b200_comm_init``)
:167 s_read_data_files                 arrays handed to :meth:`Simulation.upload`
:170 s_populate_grid_variables_buffers :func:`microfc_b200.domain.ghosted_metrics`
:188-193 ``!$acc update device``       ``mfc_b200_upload``
:205-318 time loop                     :meth:`Simulation.run` (dt end-tweak :287, last
                                       iteration without update m_time_steppers.fpp:296)
:229-235 s_3rd_order_tvd_rk            ``mfc_b200_step``
:296 ``!$acc update host``             ``mfc_b200_download``
:329-341 finalisers                    ``mfc_b200_finalize``
=====================================  =====================================================

The CUDA library is the only compute path: if it is not built or no GPU is present, the
constructor raises (there is no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np

from . import abi
from .case import CaseConfig
from .domain import RankLayout, ghosted_metrics, rank_layout


def patch_array(cfg: CaseConfig):
    """patch_icpp(1:num_patches) of the case as the ``mfc_b200_patch_t`` array of the C ABI."""
    n = cfg.num_patches
    if not 1 <= n <= abi.MAX_PATCHES:
        raise ValueError("num_patches must be 1..%d" % abi.MAX_PATCHES)
    arr = (abi.Patch * n)()
    for i, pt in enumerate(cfg.patches):
        c = arr[i]
        c.geometry, c.smoothen, c.smooth_patch_id = pt.geometry, int(pt.smoothen), pt.smooth_patch_id
        for k in range(abi.MAX_PATCHES + 1):
            c.alter_patch[k] = int(bool(pt.alter_patch.get(k, False))) if k <= n else 0
        c.x_centroid, c.y_centroid, c.z_centroid = pt.x_centroid, pt.y_centroid, pt.z_centroid
        c.length_x, c.length_y, c.length_z = pt.length_x, pt.length_y, pt.length_z
        c.radius, c.epsilon, c.smooth_coeff, c.pres = pt.radius, pt.epsilon, pt.smooth_coeff, pt.pres
        for k in range(3):
            c.radii[k], c.normal[k], c.vel[k] = pt.radii[k], pt.normal[k], pt.vel[k]
        for k in range(abi.MAX_FLUIDS):
            c.alpha_rho[k], c.alpha[k] = pt.alpha_rho[k], pt.alpha[k]
    return arr


def rank_cell_centres(cfg: CaseConfig, lay: RankLayout, cb_glb: List[np.ndarray]):
    """pre_process' cell centres (x_cb(i-1) + x_cb(i))/2 (m_start_up.fpp:717,743) of one rank's
    interior cells, per active direction, and the GLOBAL minimum cell width (s_mpi_reduce_min,
    :720) -- what ``mfc_b200_generate_initial_condition`` takes besides the patches."""
    zs, ys, xs = lay.interior_slices()
    sl = (xs, ys, zs)
    cc, dmin = [], []
    for d in range(cfg.num_dims):
        cb = cb_glb[d]
        cc.append(np.ascontiguousarray(((cb[1:] + cb[:-1]) / 2.0)[sl[d]]))
        dmin.append(float(np.min(cb[1:] - cb[:-1])))
    return cc, min(dmin)


class Simulation:
    def __init__(self, cfg: CaseConfig, cb_glb: List[np.ndarray], rank: int = 0, num_procs: int = 1,
                 strict: bool = False, device: int = -1, unique_id: Optional[bytes] = None,
                 broadcast_id=None):
        """``broadcast_id``: callable(bytes|None) -> bytes that broadcasts rank 0's NCCL unique
        id (the Fortran host would MPI_BCAST it; bench.py uses torch.distributed)."""
        self.cfg = cfg
        self.rank, self.num_procs = rank, num_procs
        self.layouts = [rank_layout(r, num_procs, cfg) for r in range(num_procs)]
        self.layout: RankLayout = self.layouts[rank]
        self.metrics = ghosted_metrics(self.layout, cfg, cb_glb, self.layouts)
        self.b = cfg.buff_size
        self.E = cfg.sys_size
        lay = self.layout
        nd = cfg.num_dims
        self.ghost_shape = tuple(lay.N[d] + 1 + 2 * self.b if d < nd else 1 for d in (2, 1, 0))
        self.local_shape = lay.shape

        p = abi.Params()
        p.abi_version = abi.ABI_VERSION
        p.m, p.n, p.p = lay.N
        p.m_glb, p.n_glb, p.p_glb = cfg.m, cfg.n, cfg.p
        p.num_dims, p.num_fluids, p.sys_size, p.buff_size = nd, cfg.num_fluids, self.E, self.b
        p.weno_order, p.weno_eps = cfg.weno_order, cfg.weno_eps
        p.time_stepper, p.weno_Re_flux, p.run_time_info = cfg.time_stepper, int(cfg.weno_Re_flux), int(cfg.run_time_info)
        p.t_step_start, p.t_step_stop = cfg.t_step_start, cfg.t_step_stop
        for d in range(3):
            p.bc[2 * d], p.bc[2 * d + 1] = lay.bc[d]
            p.proc_coords[d] = lay.coords[d]
            p.num_procs_dir[d] = lay.np_dir[d]
        p.proc_rank, p.num_procs = rank, num_procs
        for i in range(cfg.num_fluids):
            p.gammas[i], p.pi_infs[i] = cfg.gamma[i], cfg.pi_inf[i]
            p.Re[i][0], p.Re[i][1] = cfg.Re[i][0], cfg.Re[i][1]
        self._keep = []
        for d in range(nd):
            for name, arrs in (("cb", self.metrics.cb), ("cc", self.metrics.cc), ("ds", self.metrics.ds)):
                a = np.ascontiguousarray(arrs[d], dtype=np.float64)
                self._keep.append(a)
                getattr(p, name)[d] = a.ctypes.data_as(abi.c_double_p)
        p.strict_math = int(strict)
        p.device = device
        self.params = p
        self.L = abi.lib()
        abi.check(self.L.mfc_b200_init(C.byref(p)))
        self._open = True
        if num_procs > 1:
            if unique_id is None:
                if broadcast_id is None:
                    raise ValueError("num_procs > 1 needs unique_id or broadcast_id")
                mine = None
                if rank == 0:
                    buf = C.create_string_buffer(128)
                    abi.check(self.L.mfc_b200_get_unique_id(buf))
                    mine = buf.raw
                unique_id = broadcast_id(mine)
            abi.check(self.L.mfc_b200_comm_init(unique_id, rank, num_procs))

    # ---- state transfer --------------------------------------------------------------------
    def ghosted(self, q_local: np.ndarray) -> np.ndarray:
        """Local interior (E, Nz, Ny, Nx) -> the host's ghosted fields sf(-b:m+b, ...)."""
        out = np.zeros((self.E,) + self.ghost_shape)
        out[self._interior()] = q_local
        return out

    def _interior(self):
        b, nd, lay = self.b, self.cfg.num_dims, self.layout
        sl = [slice(None)]
        for d in (2, 1, 0):
            sl.append(slice(b, b + lay.N[d] + 1) if d < nd else slice(0, 1))
        return tuple(sl)

    def scatter(self, q_glb: np.ndarray) -> np.ndarray:
        return np.ascontiguousarray(q_glb[(slice(None),) + self.layout.interior_slices()])

    def upload(self, q_local: np.ndarray) -> None:
        """p_main.fpp:188-193; q_local: this rank's interior cells (E, Nz, Ny, Nx)."""
        self.upload_ghosted(self.ghosted(q_local))

    def generate_initial_condition(self, cb_glb: List[np.ndarray]) -> None:
        """The reference's pre_process on the device: patches -> conservative
        state of this rank, written straight into HBM by ``mfc_b200_generate_initial_condition``
        (replaces pre_process' restart files + ``mfc_b200_upload``).  The host only supplies
        what pre_process itself derives from the grid files: its cell centres
        ``(x_cb(i-1) + x_cb(i))/2`` (m_start_up.fpp:717,743) for this rank's cells and the global
        minimum cell width (s_mpi_reduce_min, :720)."""
        arr = patch_array(self.cfg)
        cc, ds_min = rank_cell_centres(self.cfg, self.layout, cb_glb)
        zs, ys, xs = self.layout.interior_slices()
        cbr = [np.ascontiguousarray(cb_glb[d][1:][(xs, ys, zs)[d]]) for d in range(self.cfg.num_dims)]   # x_cb(0:m), ...
        ptrs, bptrs = (abi.c_double_p * 3)(), (abi.c_double_p * 3)()
        for d, c in enumerate(cc):
            ptrs[d] = c.ctypes.data_as(abi.c_double_p)
            bptrs[d] = cbr[d].ctypes.data_as(abi.c_double_p)
        abi.check(self.L.mfc_b200_generate_initial_condition2(len(arr), arr, ptrs, bptrs, ds_min))

    def upload_ghosted(self, q_ghosted: np.ndarray) -> None:
        assert q_ghosted.shape == (self.E,) + self.ghost_shape and q_ghosted.dtype == np.float64
        abi.check(self.L.mfc_b200_upload(abi.field_pointers(list(q_ghosted))))

    def download(self, out_ghosted: Optional[np.ndarray] = None) -> np.ndarray:
        """p_main.fpp:296; returns this rank's interior cells."""
        buf = out_ghosted if out_ghosted is not None else np.empty((self.E,) + self.ghost_shape)
        self.download_ghosted(buf)
        return np.ascontiguousarray(buf[self._interior()])

    def download_ghosted(self, out_ghosted: np.ndarray) -> None:
        """p_main.fpp:296 exactly: the device state lands in the host's own ghosted fields
        sf(-b:m+b, ...) (no host-side repacking)."""
        assert out_ghosted.shape == (self.E,) + self.ghost_shape and out_ghosted.dtype == np.float64
        abi.check(self.L.mfc_b200_download(abi.field_pointers(list(out_ghosted))))

    def download_prim(self) -> np.ndarray:
        buf = np.empty((self.E,) + self.ghost_shape)
        abi.check(self.L.mfc_b200_download_prim(abi.field_pointers(list(buf))))
        return np.ascontiguousarray(buf[self._interior()])

    # ---- the hot path ----------------------------------------------------------------------
    def step(self, t_step: int, dt: float):
        """One call of s_{1st,2nd,3rd}_order_tvd_rk; returns [ICFL, VCFL, Rc] (nan if unset)."""
        stab = (C.c_double * 3)(float("nan"), float("nan"), float("nan"))
        secs = C.c_double(0.0)
        abi.check(self.L.mfc_b200_step(t_step, dt, stab, C.byref(secs)))
        self.last_step_seconds = secs.value
        return list(stab)

    def step_async(self, t_step: int, dt: float, n_steps: int) -> None:
        abi.check(self.L.mfc_b200_step_async(t_step, dt, n_steps))

    def sync(self) -> None:
        abi.check(self.L.mfc_b200_sync())

    def compute_rhs(self, q_local: np.ndarray) -> np.ndarray:
        """s_compute_rhs on an arbitrary state, m_rhs.fpp:405."""
        qg = self.ghosted(q_local)
        rhs = np.empty((self.E,) + self.local_shape)
        abi.check(self.L.mfc_b200_compute_rhs(abi.field_pointers(list(qg)), abi.field_pointers(list(rhs))))
        return rhs

    def run(self, callback=None):
        """The time loop of p_main.fpp:196-318.  Returns the rows of run_time.inf
        (t_step, dt used, [ICFL, VCFL, Rc])."""
        cfg = self.cfg
        t_step = cfg.t_step_start
        dt = cfg.dt
        mytime = 0.0 if t_step == 0 else t_step * dt                    # :197-201
        finaltime = cfg.t_step_stop * dt                                # :202
        rows = []
        while True:
            mytime = mytime + dt                                        # :214
            stab = self.step(t_step, dt)                                # :229-235
            rows.append((t_step, dt, stab))
            # m_data_output.fpp:296-305: rank 0 detects it and calls s_mpi_abort, which takes every rank
            # down.  The all-reduced criteria are identical on every rank, so each one raises by itself
            # (a rank that carried on would hang in the next halo exchange).
            if cfg.run_time_info:
                if stab[0] != stab[0]:
                    raise FloatingPointError("ICFL is NaN. Exiting ...")
                if stab[0] > 1.0:
                    raise FloatingPointError("ICFL is greater than 1.0. Exiting ...")
            if t_step == cfg.t_step_stop:                               # :239
                break
            if (mytime + dt) >= finaltime:                              # :287
                dt = finaltime - mytime
            t_step += 1                                                 # :288
            if callback is not None:
                callback(self, t_step)
        return rows

    # ---- introspection ---------------------------------------------------------------------
    def weno_coefficients(self, d: int):
        nc = self.layout.N[d] + 1 + 2 * self.b - 2 * self.cfg.weno_polyn
        pL, pR = np.empty((nc, 3, 2)), np.empty((nc, 3, 2))
        dL, dR, bt = np.empty((nc, 3)), np.empty((nc, 3)), np.empty((nc, 3, 3))
        abi.check(self.L.mfc_b200_get_weno_coefficients(d, *[a.ctypes.data_as(abi.c_double_p) for a in (pL, pR, dL, dR, bt)]))
        return dict(poly_L=pL, poly_R=pR, d_L=dL, d_R=dR, beta=bt)

    def kernel_launches(self) -> int:
        return int(self.L.mfc_b200_kernel_launches())

    def snapshot(self) -> None:
        abi.check(self.L.mfc_b200_state_snapshot())

    def restore(self) -> None:
        abi.check(self.L.mfc_b200_state_restore())

    def timer_start(self) -> None:
        abi.check(self.L.mfc_b200_timer_start())

    def timer_stop(self) -> float:
        s = C.c_double(0.0)
        abi.check(self.L.mfc_b200_timer_stop(C.byref(s)))
        return s.value

    def profile(self, on: bool) -> None:
        abi.check(self.L.mfc_b200_profile_enable(int(on)))

    def profile_report(self):
        out = {}
        for kc in range(16):
            name = self.L.mfc_b200_kernel_name(kc)
            if not name:
                break
            s, n = C.c_double(0), C.c_int64(0)
            abi.check(self.L.mfc_b200_profile_get(kc, C.byref(s), C.byref(n)))
            out[name.decode()] = (s.value, n.value)
        return out

    def close(self) -> None:
        if getattr(self, "_open", False):
            self.L.mfc_b200_finalize()
            self._open = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
