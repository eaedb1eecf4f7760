"""Host-side restatement of the reference ``pre_process`` executable, as far as the hot path
needs inputs: grid generation and patch-based initial conditions.

There is no pre_process binary to run here (no Fortran toolchain), and the simulation only
sees its output files, so this module produces the same arrays in memory:

* global cell boundaries -- ``s_generate_parallel_grid`` (src/pre_process/m_grid.f90:144-241;
  every example sets parallel_io = T);
* cell centres / minimum widths as pre_process sees them after reading the grid back --
  ``s_read_parallel_grid_data_files`` (src/pre_process/m_start_up.fpp:685-757);
* patches in order with the ``alter_patch`` permission mask and boundary smoothing --
  ``s_generate_initial_condition`` (m_initial_condition.fpp:42-113), geometries
  1 line segment / 2 circle / 3 rectangle / 4 sweep line / 5 ellipse / 18 varcircle
  (m_create_patches.fpp:47-88,96-146,262-313,324-370,200-251,148-192),
  ``s_assign_patch_species_primitive_variables`` (m_assign_patches.fpp:54-165);
* primitive -> conservative (src/common/m_variables_conversion.fpp:385-443).

Geometries 8 (sphere), 9 (cuboid), 10 (cylinder along z) are the 3-D EXTENSION.

numpy evaluates every expression elementwise in IEEE double in the order written, so the
operand order of the Fortran statements is kept.  (tanh/log/cosh come from numpy rather than
glibc; they only shape the inputs, which the CUDA path and the oracle share.)
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .case import CaseConfig, DFLT_REAL


def generate_grid(cfg: CaseConfig) -> List[np.ndarray]:
    """Global cell-boundary arrays s_cb_glb(-1:N_glb) per active direction (index 0 <-> -1)."""
    out = []
    N = [cfg.m, cfg.n, cfg.p]
    for d in range(cfg.num_dims):
        beg, end = cfg.domain[d]
        ds = (end - beg) / float(N[d] + 1)                       # m_grid.f90:166,181
        cb = np.empty(N[d] + 2)
        i = np.arange(0, N[d] + 1, dtype=np.float64)
        cb[:-1] = beg + ds * i                                   # :167-169
        cb[-1] = end                                             # :170
        if cfg.stretch[d]:                                       # :171-187
            length = abs(cb[-1] - cb[0])
            cb = cb / length
            s_a, s_b, a = cfg.s_a[d] / length, cfg.s_b[d] / length, cfg.a_s[d]
            for _ in range(cfg.loops[d]):
                cb = cb / a * (a + np.log(np.cosh(a * (cb - s_a))) + np.log(np.cosh(a * (cb - s_b)))
                               - 2.0 * np.log(np.cosh(a * (s_b - s_a) / 2.0)))
            cb = cb * length
        out.append(cb)
    return out


def _pre_process_centres(cb_glb: List[np.ndarray]) -> Tuple[List[np.ndarray], List[float]]:
    cc, dmin = [], []
    for cb in cb_glb:
        cc.append((cb[1:] + cb[:-1]) / 2.0)                      # m_start_up.fpp:717,743
        dmin.append(float(np.min(cb[1:] - cb[:-1])))             # :719,745 (+ s_mpi_reduce_min)
    return cc, dmin


def generate_initial_condition(cfg: CaseConfig, cb_glb: List[np.ndarray], box=None) -> np.ndarray:
    """Conservative variables, shape (sys_size, Nz, Ny, Nx), C-order so that x is fastest
    exactly like the Fortran arrays sf(0:m, 0:n[, 0:p]).  ``box`` = (slice_z, slice_y, slice_x)
    restricts the output to one rank's cells (pre_process runs decomposed too; the smoothing
    length uses the GLOBAL minimum cell width, s_mpi_reduce_min m_start_up.fpp:720)."""
    nd, nf, E = cfg.num_dims, cfg.num_fluids, cfg.sys_size
    cc, dmin = _pre_process_centres(cb_glb)
    if box is not None:
        cc = [cc[d][box[2 - d]] for d in range(nd)]
    Nx = len(cc[0]); Ny = len(cc[1]) if nd > 1 else 1; Nz = len(cc[2]) if nd > 2 else 1
    X = cc[0].reshape(1, 1, Nx)
    Y = cc[1].reshape(1, Ny, 1) if nd > 1 else np.zeros((1, 1, 1))
    Z = cc[2].reshape(Nz, 1, 1) if nd > 2 else np.zeros((1, 1, 1))
    # right cell boundaries x_cb(0:m) (the analytical patches evaluate their bump at x_cb(i), not x_cc(i))
    cbr = [cb_glb[d][1:] for d in range(nd)]
    if box is not None:
        cbr = [cbr[d][box[2 - d]] for d in range(nd)]
    XB = cbr[0].reshape(1, 1, Nx)
    YB = cbr[1].reshape(1, Ny, 1) if nd > 1 else np.zeros((1, 1, 1))
    dmin_all = min(dmin)                                         # min(dx, dy) in the smoothing formulas
    shape = (Nz, Ny, Nx)
    contxe, momxb, E_idx, advxb = nf, nf, nf + nd, nf + nd + 1    # 0-based starts / exclusive ends

    q_prim = np.zeros((E,) + shape)
    patch_id_fp = np.zeros(shape, dtype=np.int64)                # m_assign_patches.fpp:200

    for pid, pt in enumerate(cfg.patches, start=1):
        geo = pt.geometry
        eta = np.ones(shape)
        smoothable = False
        pres_factor = None                                       # analytical patches: pressure bump
        if geo == 6:                                             # s_isentropic_vortex, m_create_patches.fpp:379-419: a hard circle
            inside = np.broadcast_to((X - pt.x_centroid) ** 2 + (Y - pt.y_centroid) ** 2 <= pt.radius ** 2, shape)
        elif geo == 15:                                          # s_1D_analytical, :424-473
            xb, xe = pt.x_centroid - 0.5 * pt.length_x, pt.x_centroid + 0.5 * pt.length_x
            inside = np.broadcast_to((xb <= X) & (xe >= X), shape)
            pres_factor = np.broadcast_to(1.0 + 0.2 * np.exp(-1.0 * ((XB - pt.x_centroid) ** 2) / (2.0 * 0.005)), shape)   # :466-467
        elif geo == 7:                                           # s_2D_analytical, :479-534
            xb, xe = pt.x_centroid - 0.5 * pt.length_x, pt.x_centroid + 0.5 * pt.length_x
            yb, ye = pt.y_centroid - 0.5 * pt.length_y, pt.y_centroid + 0.5 * pt.length_y
            inside = np.broadcast_to((xb <= X) & (xe >= X) & (yb <= Y) & (ye >= Y), shape)
            pres_factor = np.broadcast_to(1.0 + 0.2 * np.exp(-1.0 * ((XB - pt.x_centroid) ** 2 + (YB - pt.y_centroid) ** 2) / (2.0 * 0.005)), shape)   # :527-528
        elif geo == 1:                                             # s_line_segment, m_create_patches.fpp:47-88
            xb, xe = pt.x_centroid - 0.5 * pt.length_x, pt.x_centroid + 0.5 * pt.length_x
            inside = np.broadcast_to((xb <= X) & (xe >= X), shape)
        elif geo == 2 or geo == 10:                              # s_circle, :96-146 (10: z-invariant cylinder)
            smoothable = True
            r2 = (X - pt.x_centroid) ** 2 + (Y - pt.y_centroid) ** 2
            if pt.smoothen:
                eta = np.broadcast_to(np.tanh(pt.smooth_coeff / dmin_all * (np.sqrt(r2) - pt.radius)) * (-0.5) + 0.5, shape)
            inside = np.broadcast_to(r2 <= pt.radius ** 2, shape)
        elif geo == 18:                                          # s_varcircle, :148-192
            myr = np.sqrt((X - pt.x_centroid) ** 2 + (Y - pt.y_centroid) ** 2)
            inside = np.broadcast_to((myr <= pt.radius + pt.epsilon / 2.0) & (myr >= pt.radius - pt.epsilon / 2.0), shape)
        elif geo == 5:                                           # s_ellipse, :200-251
            smoothable = True
            a, b = pt.radii[0], pt.radii[1]
            r2 = ((X - pt.x_centroid) / a) ** 2 + ((Y - pt.y_centroid) / b) ** 2
            if pt.smoothen:
                eta = np.broadcast_to(np.tanh(pt.smooth_coeff / dmin_all * (np.sqrt(r2) - 1.0)) * (-0.5) + 0.5, shape)
            inside = np.broadcast_to(r2 <= 1.0, shape)
        elif geo == 3:                                           # s_rectangle, :262-313
            xb, xe = pt.x_centroid - 0.5 * pt.length_x, pt.x_centroid + 0.5 * pt.length_x
            yb, ye = pt.y_centroid - 0.5 * pt.length_y, pt.y_centroid + 0.5 * pt.length_y
            inside = np.broadcast_to((xb <= X) & (xe >= X) & (yb <= Y) & (ye >= Y), shape)
        elif geo == 4:                                           # s_sweep_line, :324-370
            smoothable = True
            a, b = pt.normal[0], pt.normal[1]
            c = -a * pt.x_centroid - b * pt.y_centroid
            lin = a * X + b * Y + c
            if pt.smoothen:
                eta = np.broadcast_to(5e-1 + 5e-1 * np.tanh(pt.smooth_coeff / dmin_all * lin / np.sqrt(a ** 2 + b ** 2)), shape)
            inside = np.broadcast_to(lin >= 0.0, shape)
        elif geo == 8:                                           # EXTENSION: sphere
            smoothable = True
            r2 = (X - pt.x_centroid) ** 2 + (Y - pt.y_centroid) ** 2 + (Z - pt.z_centroid) ** 2
            if pt.smoothen:
                eta = np.broadcast_to(np.tanh(pt.smooth_coeff / dmin_all * (np.sqrt(r2) - pt.radius)) * (-0.5) + 0.5, shape)
            inside = np.broadcast_to(r2 <= pt.radius ** 2, shape)
        elif geo == 9:                                           # EXTENSION: cuboid
            xb, xe = pt.x_centroid - 0.5 * pt.length_x, pt.x_centroid + 0.5 * pt.length_x
            yb, ye = pt.y_centroid - 0.5 * pt.length_y, pt.y_centroid + 0.5 * pt.length_y
            zb, ze = pt.z_centroid - 0.5 * pt.length_z, pt.z_centroid + 0.5 * pt.length_z
            inside = np.broadcast_to((xb <= X) & (xe >= X) & (yb <= Y) & (ye >= Y) & (zb <= Z) & (ze >= Z), shape)
        else:
            raise NotImplementedError(f"patch geometry {geo} is not restated (MFC_B200_EUNSUPPORTED)")

        # alter_patch(patch_id_fp(i,j)) -- the permission of this patch to overwrite what is there
        perm = np.zeros(cfg.num_patches + 1, dtype=bool)
        for k, v in pt.alter_patch.items():
            if 0 <= k <= cfg.num_patches:
                perm[k] = v
        mask = inside & perm[patch_id_fp]
        if smoothable:
            mask = mask | (patch_id_fp == pt.smooth_patch_id)
        if not mask.any():
            continue

        # s_assign_patch_species_primitive_variables, m_assign_patches.fpp:54-165
        orig = q_prim.copy() if pt.smoothen else q_prim          # eta == 1 otherwise: orig is multiplied by 0
        one_m_eta = 1.0 - eta
        def blend(val, o):
            return eta * val + one_m_eta * o                      # :132-141,:149-158
        for i in range(nf):
            q_prim[i] = np.where(mask, blend(pt.alpha_rho[i], orig[i]), q_prim[i])
            q_prim[advxb + i] = np.where(mask, blend(pt.alpha[i], orig[advxb + i]), q_prim[advxb + i])
        for i in range(nd):
            q_prim[momxb + i] = np.where(mask, blend(pt.vel[i], orig[momxb + i]), q_prim[momxb + i])
        q_prim[E_idx] = np.where(mask, blend(pt.pres, orig[E_idx]), q_prim[E_idx])
        if pres_factor is not None:
            q_prim[E_idx] = np.where(mask, q_prim[E_idx] * pres_factor, q_prim[E_idx])
        patch_id_fp = np.where(mask & (one_m_eta < 1e-16), pid, patch_id_fp)    # :163

    return prim_to_cons(cfg, q_prim)


def prim_to_cons(cfg: CaseConfig, q_prim: np.ndarray) -> np.ndarray:
    """s_convert_primitive_to_conservative_variables, m_variables_conversion.fpp:385-443."""
    nd, nf = cfg.num_dims, cfg.num_fluids
    momxb, E_idx, advxb = nf, nf + nd, nf + nd + 1
    q_cons = np.empty_like(q_prim)
    rho = np.zeros(q_prim.shape[1:]); gamma = np.zeros_like(rho); pi_inf = np.zeros_like(rho)
    for i in range(nf):                                          # :147-161
        rho = rho + q_prim[i]
        gamma = gamma + q_prim[advxb + i] * cfg.gamma[i]
        pi_inf = pi_inf + q_prim[advxb + i] * cfg.pi_inf[i]
    for i in range(nf):
        q_cons[i] = q_prim[i]                                    # :417-419
    dyn_pres = np.zeros_like(rho)
    for i in range(momxb, E_idx):                                # :426-430
        q_cons[i] = rho * q_prim[i]
        dyn_pres = dyn_pres + q_cons[i] * q_prim[i] / 2.0
    q_cons[E_idx] = gamma * q_prim[E_idx] + dyn_pres + pi_inf    # :434-435
    for i in range(advxb, advxb + nf):
        q_cons[i] = q_prim[i]                                    # :439-441
    return q_cons
