"""The reference's example cases, restated as flat MFC case dictionaries with the resolution
as a parameter (BASELINE.json configs run them at other sizes than shipped).  At the shipped
resolution each function reproduces the JSON its ``examples/<name>/case.py`` prints
(checked by tests/test_host_logic.py when /root/reference is present).

``shockbubble_3d`` is an EXTENSION with no reference counterpart (the reference is 1-D/2-D).
"""
from __future__ import annotations

import math
from typing import Dict

from .case import CaseConfig, parse_case


def sod_1d(Nx: int = 399, Nt: int = 1000) -> Dict:
    """examples/1D_sodshocktube/case.py"""
    Tend = 0.1
    return {
        'run_time_info': 'T', 'x_domain%beg': 0.0, 'x_domain%end': 1.0, 'm': Nx, 'n': 0,
        'dt': Tend / (1. * Nt), 't_step_start': 0, 't_step_stop': int(Nt),
        't_step_save': int(math.ceil(Nt / 10.)),
        'num_patches': 2, 'num_fluids': 1, 'time_stepper': 3, 'weno_order': 5, 'weno_eps': 1.E-16,
        'bc_x%beg': -3, 'bc_x%end': -3, 'parallel_io': 'T',
        'patch_icpp(1)%geometry': 1, 'patch_icpp(1)%x_centroid': 0.25, 'patch_icpp(1)%length_x': 0.5,
        'patch_icpp(1)%vel(1)': 0.0, 'patch_icpp(1)%pres': 1.0,
        'patch_icpp(1)%alpha_rho(1)': 1.E+00, 'patch_icpp(1)%alpha(1)': 1.,
        'patch_icpp(2)%geometry': 1, 'patch_icpp(2)%x_centroid': 0.75, 'patch_icpp(2)%length_x': 0.5,
        'patch_icpp(2)%vel(1)': 0.0, 'patch_icpp(2)%pres': 0.1,
        'patch_icpp(2)%alpha_rho(1)': 0.125E+00, 'patch_icpp(2)%alpha(1)': 1.,
        'fluid_pp(1)%gamma': 1.E+00 / (1.4 - 1.E+00), 'fluid_pp(1)%pi_inf': 0.0,
    }


def kapila_1d(Nx: int = 999, Nt: int = 6025) -> Dict:
    """examples/1D_kapilashocktube/case.py (water/air, stiffened gas)"""
    return {
        'run_time_info': 'T', 'x_domain%beg': 0.0, 'x_domain%end': 1.0, 'm': Nx, 'n': 0,
        'dt': 4.E-08 * (1000.0 / (Nx + 1)), 't_step_start': 0, 't_step_stop': Nt, 't_step_save': 100,
        'num_patches': 2, 'num_fluids': 2, 'time_stepper': 3, 'weno_order': 5, 'weno_eps': 1.E-16,
        'bc_x%beg': -3, 'bc_x%end': -3, 'parallel_io': 'T',
        'patch_icpp(1)%geometry': 1, 'patch_icpp(1)%x_centroid': 0.5, 'patch_icpp(1)%length_x': 1.0,
        'patch_icpp(1)%vel(1)': 0.0, 'patch_icpp(1)%pres': 1.E+09,
        'patch_icpp(1)%alpha_rho(1)': 1000.0, 'patch_icpp(1)%alpha_rho(2)': 0.,
        'patch_icpp(1)%alpha(1)': 1.0, 'patch_icpp(1)%alpha(2)': 0.0,
        'patch_icpp(2)%geometry': 1, 'patch_icpp(2)%x_centroid': 0.85, 'patch_icpp(2)%length_x': 0.3,
        'patch_icpp(2)%alter_patch(1)': 'T', 'patch_icpp(2)%vel(1)': 0.0, 'patch_icpp(2)%pres': 1.E+05,
        'patch_icpp(2)%alpha_rho(1)': 0., 'patch_icpp(2)%alpha_rho(2)': 50.0,
        'patch_icpp(2)%alpha(1)': 0, 'patch_icpp(2)%alpha(2)': 1.,
        'fluid_pp(1)%gamma': 1.0 / (4.4 - 1.0), 'fluid_pp(1)%pi_inf': 4.4 * 6.E+08 / (4.4 - 1.0),
        'fluid_pp(2)%gamma': 1.0 / (1.4 - 1.0), 'fluid_pp(2)%pi_inf': 0.0,
    }


def vacuum_1d(Nx: int = 199, Nt: int = 15000) -> Dict:
    """examples/1D_vacuum/case.py (WENO3)"""
    d = {
        'run_time_info': 'T', 'x_domain%beg': 0.0, 'x_domain%end': 1.0, 'm': Nx, 'n': 0,
        'dt': 5.E-08 * (200.0 / (Nx + 1)), 't_step_start': 0, 't_step_stop': Nt, 't_step_save': 1000,
        'num_patches': 2, 'num_fluids': 2, 'time_stepper': 3, 'weno_order': 3, 'weno_eps': 1.E-16,
        'bc_x%beg': -3, 'bc_x%end': -3, 'parallel_io': 'T',
        'fluid_pp(1)%gamma': 1.0 / (4.4 - 1.0), 'fluid_pp(1)%pi_inf': 4.4 * 6.E+08 / (4.4 - 1.0),
        'fluid_pp(2)%gamma': 1.0 / (1.4 - 1.0), 'fluid_pp(2)%pi_inf': 0.0,
    }
    for i, (xc, u) in enumerate(((0.25, -100.0), (0.75, 100.0)), start=1):
        d.update({
            f'patch_icpp({i})%geometry': 1, f'patch_icpp({i})%x_centroid': xc, f'patch_icpp({i})%length_x': 0.5,
            f'patch_icpp({i})%vel(1)': u, f'patch_icpp({i})%pres': 1.E+05,
            f'patch_icpp({i})%alpha_rho(1)': 1000.0 * 0.99, f'patch_icpp({i})%alpha_rho(2)': 10.0 * 0.01,
            f'patch_icpp({i})%alpha(1)': 0.99, f'patch_icpp({i})%alpha(2)': 0.01,
        })
    d['patch_icpp(2)%alter_patch(1)'] = 'T'
    return d


def advection_2d(N: int = 99, Nt: int = 1000) -> Dict:
    """examples/2D_advection/case.py; BASELINE config 2 runs it at N = 1023 with dt scaled by
    100/(N+1)."""
    return {
        'run_time_info': 'T', 'x_domain%beg': 0.0, 'x_domain%end': 1.0,
        'y_domain%beg': 0.0, 'y_domain%end': 1.0, 'm': N, 'n': N,
        'dt': 5.E-07 * (100.0 / (N + 1)), 't_step_start': 0, 't_step_stop': Nt, 't_step_save': max(1, Nt // 10),
        'num_patches': 2, 'num_fluids': 2, 'time_stepper': 3, 'weno_order': 5, 'weno_eps': 1.E-16,
        'bc_x%beg': -3, 'bc_x%end': -3, 'bc_y%beg': -3, 'bc_y%end': -3, 'parallel_io': 'T',
        'patch_icpp(1)%geometry': 3, 'patch_icpp(1)%x_centroid': 0.5, 'patch_icpp(1)%y_centroid': 0.5,
        'patch_icpp(1)%length_x': 1.0, 'patch_icpp(1)%length_y': 1.0,
        'patch_icpp(1)%vel(1)': 100.0, 'patch_icpp(1)%vel(2)': 100.0, 'patch_icpp(1)%pres': 1.E+05,
        'patch_icpp(1)%alpha_rho(1)': 1000.0, 'patch_icpp(1)%alpha_rho(2)': 1.,
        'patch_icpp(1)%alpha(1)': 1.E-12, 'patch_icpp(1)%alpha(2)': 1. - 1.E-12,
        'patch_icpp(2)%geometry': 2, 'patch_icpp(2)%smoothen': 'T', 'patch_icpp(2)%smooth_patch_id': 1,
        'patch_icpp(2)%smooth_coeff': 0.5, 'patch_icpp(2)%x_centroid': 0.1, 'patch_icpp(2)%y_centroid': 0.1,
        'patch_icpp(2)%radius': 0.1, 'patch_icpp(2)%alter_patch(1)': 'T',
        'patch_icpp(2)%vel(1)': 100.0, 'patch_icpp(2)%vel(2)': 100.0, 'patch_icpp(2)%pres': 1.E+05,
        'patch_icpp(2)%alpha_rho(1)': 1., 'patch_icpp(2)%alpha_rho(2)': 1.0,
        'patch_icpp(2)%alpha(1)': 0, 'patch_icpp(2)%alpha(2)': 1.,
        'fluid_pp(1)%gamma': 1.0 / (2.35 - 1.0), 'fluid_pp(1)%pi_inf': 2.35 * 1.E+09 / (2.35 - 1.0),
        'fluid_pp(2)%gamma': 1.0 / (1.4 - 1.0), 'fluid_pp(2)%pi_inf': 0.0,
    }


def three_fluids(d: Dict, gamma3: float = 1.0 / (1.3 - 1.0), pi_inf3: float = 0.0) -> Dict:
    """A THREE-fluid variant of a two-fluid case (no reference example has more than two fluids; the
    reference allows up to num_fluids_max = 10): a third ideal gas joins with volume fraction 0.2 taken
    from fluid 1 in the background patch and 0.3 in every later patch, partial densities to match."""
    out = dict(d)
    out['num_fluids'] = 3
    out['fluid_pp(3)%gamma'], out['fluid_pp(3)%pi_inf'] = gamma3, pi_inf3
    for i in range(1, d['num_patches'] + 1):
        share = 0.2 if i == 1 else 0.3
        a1 = d.get(f'patch_icpp({i})%alpha(1)', 0.0)
        a2 = d.get(f'patch_icpp({i})%alpha(2)', 0.0)
        donor = 1 if a1 >= a2 else 2
        big = max(a1, a2)
        out[f'patch_icpp({i})%alpha(3)'] = share * big
        out[f'patch_icpp({i})%alpha({donor})'] = big - share * big
        rho_d = d.get(f'patch_icpp({i})%alpha_rho({donor})', 0.0)
        out[f'patch_icpp({i})%alpha_rho(3)'] = 0.5 * share * rho_d
        out[f'patch_icpp({i})%alpha_rho({donor})'] = rho_d - share * rho_d
    return out


def _shockbubble_common(dx: float, leng: float, vel: float):
    ps = 248758.567
    c_l = math.sqrt(1.4 * ps / 1.)
    dt = 0.1 * dx / c_l
    Nt = int((5 * leng / vel) / dt)
    return ps, dt, Nt


def shockbubble_2d(Ny: float = 100., Nx: float | None = None, Nt: int | None = None) -> Dict:
    """examples/2D_shockbubble/case.py (air / helium as shipped; BASELINE.json says
    "air/water" -- the shipped fluids are kept).  The shipped case passes m = Nx, n = Ny,
    i.e. (Nx+1) x (Ny+1) cells; BASELINE config 3 scales it to 4096 x 4096 cells."""
    leng, vel = 1., 230.
    if Nx is None:
        Nx = Ny * 3
    dx = leng / Nx
    ps, dt, Nt_full = _shockbubble_common(dx, leng, vel)
    if Nt is None:
        Nt = Nt_full
    d = {
        'run_time_info': 'T', 'x_domain%beg': -leng / 2., 'x_domain%end': leng / 2 + 2 * leng,
        'y_domain%beg': -leng / 2., 'y_domain%end': leng / 2., 'm': int(Nx), 'n': int(Ny),
        'dt': dt, 't_step_start': 0, 't_step_stop': Nt, 't_step_save': max(1, int(Nt / 20.)),
        'num_patches': 3, 'num_fluids': 2, 'time_stepper': 3, 'weno_order': 5, 'weno_eps': 1.E-16,
        'weno_Re_flux': 'F', 'bc_x%beg': -6, 'bc_x%end': -6, 'bc_y%beg': -6, 'bc_y%end': -6,
        'parallel_io': 'T',
        'patch_icpp(1)%geometry': 3, 'patch_icpp(1)%x_centroid': 0., 'patch_icpp(1)%y_centroid': 0.,
        'patch_icpp(1)%length_x': 10 * leng, 'patch_icpp(1)%length_y': leng,
        'patch_icpp(1)%vel(1)': vel, 'patch_icpp(1)%vel(2)': 0.0, 'patch_icpp(1)%pres': 101325.,
        'patch_icpp(1)%alpha_rho(1)': 1.29, 'patch_icpp(1)%alpha_rho(2)': 0.0,
        'patch_icpp(1)%alpha(1)': 1.0, 'patch_icpp(1)%alpha(2)': 0.0,
        'patch_icpp(2)%geometry': 3, 'patch_icpp(2)%alter_patch(1)': 'T',
        'patch_icpp(2)%x_centroid': -3 * leng / 8., 'patch_icpp(2)%y_centroid': 0.,
        'patch_icpp(2)%length_x': leng / 4., 'patch_icpp(2)%length_y': leng,
        'patch_icpp(2)%vel(1)': vel, 'patch_icpp(2)%vel(2)': 0.0, 'patch_icpp(2)%pres': ps,
        'patch_icpp(2)%alpha_rho(1)': 2.4, 'patch_icpp(2)%alpha_rho(2)': 0.0,
        'patch_icpp(2)%alpha(1)': 1.0, 'patch_icpp(2)%alpha(2)': 0.0,
        'patch_icpp(3)%geometry': 2, 'patch_icpp(3)%x_centroid': 0.0, 'patch_icpp(3)%y_centroid': 0.0,
        'patch_icpp(3)%radius': leng / 5., 'patch_icpp(3)%alter_patch(1)': 'T',
        'patch_icpp(3)%vel(1)': 0., 'patch_icpp(3)%vel(2)': 0.0, 'patch_icpp(3)%pres': 101325.,
        'patch_icpp(3)%alpha_rho(1)': 0.0, 'patch_icpp(3)%alpha_rho(2)': 0.167,
        'patch_icpp(3)%alpha(1)': 0.0, 'patch_icpp(3)%alpha(2)': 1.0,
        'fluid_pp(1)%gamma': 1.0 / (1.4 - 1.0), 'fluid_pp(1)%pi_inf': 0.,
        'fluid_pp(2)%gamma': 1.0 / (1.6666 - 1.0), 'fluid_pp(2)%pi_inf': 0.0,
    }
    return d


def shockbubble_2d_cells(ncx: int, ncy: int, Nt: int = 100) -> Dict:
    """2D_shockbubble on an ncx x ncy CELL grid over the shipped domain [-0.5,2.5]x[-0.5,0.5]
    (BASELINE config 3: 4096 x 4096); dt = 0.1 min(dx,dy)/c_l as in the shipped script."""
    d = shockbubble_2d(Ny=ncy - 1, Nx=ncx - 1, Nt=Nt)
    dxm = min(3.0 / ncx, 1.0 / ncy)
    d['dt'] = 0.1 * dxm / math.sqrt(1.4 * 248758.567 / 1.)
    return d


def shockdroplet_2d(Nx: float = 5000., Ny: float = 740., Nt: int = 10, viscous: bool = False, Re: float = 100.0) -> Dict:
    """examples/2D_shockdroplet/case.py (water droplet in air, bc_y%beg = -2 symmetry);
    viscous=True sets fluid_pp(i)%Re(1:2) for both fluids the way examples/2D_viscous/case.py:91-94
    does (BASELINE config 4).  The VALUE used there (1e-4, i.e. a viscosity of 1e4) is only stable
    at that case's dt = 5e-10; at this case's dt = 0.1 dx/c the scheme -- the CPU oracle included --
    blows up within two steps, so the default here is Re = 100 (viscous CFL ~ 5e-4)."""
    ps = 664016.5
    rho_post_a, rho_w = 3.757918216, 1000
    gam_a, gam_w, pi_w = 1.4, 6.12, 3.43E8
    vel = 575.4980523
    c_l = math.sqrt(1.4 * ps / 1)
    D = 0.022
    dx = 0.25 / Nx
    dt = 0.1 * dx / c_l
    d = {
        'run_time_info': 'F', 'x_domain%beg': 0, 'x_domain%end': 0.25,
        'y_domain%beg': 0, 'y_domain%end': 0.037, 'm': int(Nx), 'n': int(Ny),
        'dt': dt, 't_step_start': 0, 't_step_stop': Nt, 't_step_save': 1,
        'num_patches': 3, 'num_fluids': 2, 'time_stepper': 3, 'weno_order': 5, 'weno_eps': 1.E-16,
        'weno_Re_flux': 'F', 'bc_x%beg': -6, 'bc_x%end': -6, 'bc_y%beg': -2, 'bc_y%end': -6,
        'parallel_io': 'T',
        'patch_icpp(1)%geometry': 3, 'patch_icpp(1)%x_centroid': 0.25 / 2, 'patch_icpp(1)%y_centroid': 0.037 / 2,
        'patch_icpp(1)%length_x': 0.25, 'patch_icpp(1)%length_y': 0.037,
        'patch_icpp(1)%vel(1)': 0., 'patch_icpp(1)%vel(2)': 0.0, 'patch_icpp(1)%pres': 101325.,
        'patch_icpp(1)%alpha_rho(1)': 0., 'patch_icpp(1)%alpha_rho(2)': 1.17,
        'patch_icpp(1)%alpha(1)': 0.0, 'patch_icpp(1)%alpha(2)': 1.0,
        'patch_icpp(2)%geometry': 3, 'patch_icpp(2)%alter_patch(1)': 'T',
        'patch_icpp(2)%x_centroid': 0., 'patch_icpp(2)%y_centroid': 0.037 / 2,
        'patch_icpp(2)%length_x': 0.25 - D, 'patch_icpp(2)%length_y': 0.037,
        'patch_icpp(2)%vel(1)': vel, 'patch_icpp(2)%vel(2)': 0.0, 'patch_icpp(2)%pres': ps,
        'patch_icpp(2)%alpha_rho(1)': 0.0, 'patch_icpp(2)%alpha_rho(2)': rho_post_a,
        'patch_icpp(2)%alpha(1)': 0.0, 'patch_icpp(2)%alpha(2)': 1.0,
        'patch_icpp(3)%geometry': 2, 'patch_icpp(3)%x_centroid': 0.25 / 2, 'patch_icpp(3)%y_centroid': 0,
        'patch_icpp(3)%radius': D / 2, 'patch_icpp(3)%alter_patch(1)': 'T',
        'patch_icpp(3)%vel(1)': 0., 'patch_icpp(3)%vel(2)': 0.0, 'patch_icpp(3)%pres': 101325.,
        'patch_icpp(3)%alpha_rho(1)': rho_w, 'patch_icpp(3)%alpha_rho(2)': 0,
        'patch_icpp(3)%alpha(1)': 1., 'patch_icpp(3)%alpha(2)': 0.,
        'fluid_pp(1)%gamma': 1.0 / (gam_w - 1.0), 'fluid_pp(1)%pi_inf': pi_w * gam_w / (gam_w - 1.0),
        'fluid_pp(2)%gamma': 1.0 / (gam_a - 1.0), 'fluid_pp(2)%pi_inf': 0.0,
    }
    if viscous:
        for i in (1, 2):
            d[f'fluid_pp({i})%Re(1)'] = Re
            d[f'fluid_pp({i})%Re(2)'] = Re
    return d


def viscous_2d(N: int = 50, Nt: int = 10, weno_Re_flux: bool = True) -> Dict:
    """examples/2D_viscous/case.py (periodic in x, weno_Re_flux = T)"""
    p_l, p_g = 1E+06, 1E+06
    rho_l, rho_g = 1000., 1.
    v_l, v_g = 500., -500.
    return {
        'run_time_info': 'T', 'x_domain%beg': -0.5, 'x_domain%end': 0.5,
        'y_domain%beg': -0.5, 'y_domain%end': 0.5, 'm': N, 'n': N,
        'dt': 5.E-10 * (51.0 / (N + 1)), 't_step_start': 0, 't_step_stop': Nt, 't_step_save': 1,
        'num_patches': 2, 'num_fluids': 2, 'time_stepper': 3, 'weno_order': 5, 'weno_eps': 1.E-16,
        'weno_Re_flux': 'T' if weno_Re_flux else 'F',
        'bc_x%beg': -1, 'bc_x%end': -1, 'bc_y%beg': -6, 'bc_y%end': -6, 'parallel_io': 'T',
        'patch_icpp(1)%geometry': 3, 'patch_icpp(1)%x_centroid': 0., 'patch_icpp(1)%y_centroid': 0,
        'patch_icpp(1)%length_x': 1.0, 'patch_icpp(1)%length_y': 1.,
        'patch_icpp(1)%vel(1)': v_l, 'patch_icpp(1)%vel(2)': 0.0, 'patch_icpp(1)%pres': p_l,
        'patch_icpp(1)%alpha_rho(1)': rho_l, 'patch_icpp(1)%alpha_rho(2)': rho_l,
        'patch_icpp(1)%alpha(1)': 0.5, 'patch_icpp(1)%alpha(2)': 0.5,
        'patch_icpp(2)%geometry': 3, 'patch_icpp(2)%x_centroid': 0., 'patch_icpp(2)%y_centroid': 0.25,
        'patch_icpp(2)%length_x': 1.0, 'patch_icpp(2)%length_y': 0.5, 'patch_icpp(2)%alter_patch(1)': 'T',
        'patch_icpp(2)%vel(1)': v_g, 'patch_icpp(2)%vel(2)': 0.0, 'patch_icpp(2)%pres': p_g,
        'patch_icpp(2)%alpha_rho(1)': 0., 'patch_icpp(2)%alpha_rho(2)': rho_g,
        'patch_icpp(2)%alpha(1)': 0., 'patch_icpp(2)%alpha(2)': 1.0,
        'fluid_pp(1)%gamma': 1.0 / (4.4 - 1.0), 'fluid_pp(2)%gamma': 1.0 / (4.4 - 1.0),
        'fluid_pp(1)%pi_inf': 4.4 * 6.E+08 / (4.4 - 1.0), 'fluid_pp(2)%pi_inf': 4.4 * 6.E+08 / (4.4 - 1.0),
        'fluid_pp(1)%Re(1)': 0.0001, 'fluid_pp(1)%Re(2)': 0.0001,
        'fluid_pp(2)%Re(1)': 0.0001, 'fluid_pp(2)%Re(2)': 0.0001,
    }


def viscous_wave_2d(N: int = 48, Nx: int = 32, Nt: int = 100, weno_Re_flux: bool = True, bc_y: int = -1) -> Dict:
    """A smooth two-fluid low-Mach state for the viscous terms (no reference example has one: in
    examples/2D_viscous the velocity is piecewise constant, so the WENO-reconstructed
    divergence-theorem gradients of m_viscous.fpp:200-217 vanish identically).  The patches only
    lay down a uniform background; the parity tests overwrite the state with
    :func:`viscous_wave_state`, whose velocity, density and volume fraction vary in x AND y so that
    every stress component (shear and bulk, both fluids' Re) is active."""
    d = viscous_2d(N=N - 1, Nt=Nt, weno_Re_flux=weno_Re_flux)
    d.update({'m': Nx - 1, 'n': N - 1, 'x_domain%beg': 0.0, 'x_domain%end': Nx / float(N), 'y_domain%beg': 0.0,
              'y_domain%end': 1.0, 'bc_x%beg': -1, 'bc_x%end': -1, 'bc_y%beg': bc_y, 'bc_y%end': bc_y,
              'run_time_info': 'T', 'fluid_pp(1)%gamma': 2.5, 'fluid_pp(2)%gamma': 1.0 / (1.6 - 1.0),
              'fluid_pp(1)%pi_inf': 0.0, 'fluid_pp(2)%pi_inf': 0.0, 'dt': 0.2 * (1.0 / N) / math.sqrt(1.4),
              'fluid_pp(1)%Re(1)': 600.0, 'fluid_pp(1)%Re(2)': 360.0, 'fluid_pp(2)%Re(1)': 210.0, 'fluid_pp(2)%Re(2)': 480.0})
    return d


def viscous_wave_state(cfg: CaseConfig, cb):
    """Conservative state (E, 1, Ny, Nx) of :func:`viscous_wave_2d`: p = 1, Mach ~ 0.05."""
    import numpy as np
    x = (cb[0][1:] + cb[0][:-1]) / 2
    y = (cb[1][1:] + cb[1][:-1]) / 2
    Lx = cb[0][-1] - cb[0][0]
    X, Y = np.meshgrid(2 * np.pi * x / Lx, 2 * np.pi * y)          # (Ny, Nx)
    U0 = 0.05
    a1 = 0.5 + 0.2 * np.sin(Y) * np.cos(X)
    r1, r2 = 1.0 + 0.1 * np.cos(Y), 0.6 + 0.05 * np.sin(X + Y)
    u = U0 * (np.sin(Y) + 0.5 * np.cos(X))
    v = 0.3 * U0 * np.sin(X + Y)
    q = np.zeros((cfg.sys_size, 1) + X.shape)
    q[0, 0], q[1, 0] = a1 * r1, (1.0 - a1) * r2
    rho = q[0, 0] + q[1, 0]
    q[2, 0], q[3, 0] = rho * u, rho * v
    gamma = a1 * cfg.gamma[0] + (1.0 - a1) * cfg.gamma[1]
    q[4, 0] = gamma * 1.0 + 0.5 * rho * (u * u + v * v)
    q[5, 0], q[6, 0] = a1, 1.0 - a1
    return q


def shearlayer_2d(Nx: int = 319, Ny: int = 159, Nt: int = 100) -> Dict:
    """examples/2D_shearlayer/case.py (periodic x, sweep-line patch, bc_y = -5)"""
    myv = 1.
    return {
        'run_time_info': 'T', 'x_domain%beg': -0.5, 'x_domain%end': 0.5,
        'y_domain%beg': -0.25, 'y_domain%end': 0.25, 'm': Nx, 'n': Ny,
        'dt': 10.0E-7 * (320.0 / (Nx + 1)), 't_step_start': 0, 't_step_stop': Nt, 't_step_save': max(1, Nt // 10),
        'num_fluids': 2, 'num_patches': 2, 'time_stepper': 3, 'weno_order': 5, 'weno_eps': 1.0E-16,
        'bc_x%beg': -1, 'bc_x%end': -1, 'bc_y%beg': -5, 'bc_y%end': -5, 'parallel_io': 'T',
        'patch_icpp(2)%geometry': 3, 'patch_icpp(2)%x_centroid': 0.0, 'patch_icpp(2)%y_centroid': 0.0,
        'patch_icpp(2)%length_x': 2.0, 'patch_icpp(2)%length_y': 2.0,
        'patch_icpp(2)%vel(1)': myv, 'patch_icpp(2)%vel(2)': 0.0, 'patch_icpp(2)%pres': 1.01325E+05,
        'patch_icpp(2)%alpha_rho(1)': 1000.0, 'patch_icpp(2)%alpha_rho(2)': 1000. * 1E-12,
        'patch_icpp(2)%alpha(1)': 1.0 - 1.E-12, 'patch_icpp(2)%alpha(2)': 1.E-12,
        'patch_icpp(1)%geometry': 4, 'patch_icpp(1)%x_centroid': 0.0, 'patch_icpp(1)%y_centroid': 0.0,
        'patch_icpp(1)%normal(1)': 0.00624987793326E+00, 'patch_icpp(1)%normal(2)': -0.99998046932219E+00,
        'patch_icpp(1)%vel(1)': -myv, 'patch_icpp(1)%vel(2)': 0.0, 'patch_icpp(1)%pres': 1.01325E+05,
        'patch_icpp(1)%alpha_rho(1)': 1000 * 1.E-12, 'patch_icpp(1)%alpha_rho(2)': 1000.0,
        'patch_icpp(1)%alpha(1)': 1.0E-12, 'patch_icpp(1)%alpha(2)': 1 - 1.0E-12,
        'fluid_pp(1)%gamma': 1.0 / (4.4 - 1.0), 'fluid_pp(1)%pi_inf': 4.4 * 6.0E+08 / (4.4 - 1.0),
        'fluid_pp(2)%gamma': 1.0 / (4.4 - 1.0), 'fluid_pp(2)%pi_inf': 4.4 * 6.0E+08 / (4.4 - 1.0),
    }


def shockbubble_3d(nc: int = 64, Nt: int = 10, ncx: int | None = None, ncy: int | None = None,
                   ncz: int | None = None, periodic_z: bool = False, z_invariant: bool = False,
                   smooth: bool = True) -> Dict:
    """EXTENSION (no reference): the 2D_shockbubble fluids and states on a cube [-0.5,0.5]^3
    of ncx x ncy x ncz cells, the helium circle extruded to a sphere (BASELINE config 5:
    512^3 per GPU).  z_invariant=True keeps the 2-D cylinder (for the 3-D-vs-2-D cross-check)."""
    ncx = ncx or nc; ncy = ncy or nc; ncz = ncz or nc
    leng, vel, ps = 1., 230., 248758.567
    dxm = min(leng / ncx, leng / ncy, leng / ncz)
    dt = 0.1 * dxm / math.sqrt(1.4 * ps / 1.)
    bz = -1 if periodic_z else -6
    d = {
        'run_time_info': 'T',
        'x_domain%beg': -leng / 2., 'x_domain%end': leng / 2., 'y_domain%beg': -leng / 2., 'y_domain%end': leng / 2.,
        'z_domain%beg': -leng / 2., 'z_domain%end': leng / 2.,
        'm': ncx - 1, 'n': ncy - 1, 'p': ncz - 1,
        'dt': dt, 't_step_start': 0, 't_step_stop': Nt, 't_step_save': max(1, Nt),
        'num_patches': 3, 'num_fluids': 2, 'time_stepper': 3, 'weno_order': 5, 'weno_eps': 1.E-16,
        'bc_x%beg': -6, 'bc_x%end': -6, 'bc_y%beg': -6, 'bc_y%end': -6, 'bc_z%beg': bz, 'bc_z%end': bz,
        'parallel_io': 'T',
        'fluid_pp(1)%gamma': 1.0 / (1.4 - 1.0), 'fluid_pp(1)%pi_inf': 0.,
        'fluid_pp(2)%gamma': 1.0 / (1.6666 - 1.0), 'fluid_pp(2)%pi_inf': 0.0,
    }
    base = {'vel(2)': 0.0, 'vel(3)': 0.0, 'y_centroid': 0., 'z_centroid': 0.}
    p1 = dict(base, geometry=9, x_centroid=0., length_x=10 * leng, length_y=10 * leng, length_z=10 * leng)
    p1.update({'vel(1)': vel, 'pres': 101325., 'alpha_rho(1)': 1.29, 'alpha_rho(2)': 0.0, 'alpha(1)': 1.0, 'alpha(2)': 0.0})
    p2 = dict(base, geometry=9, x_centroid=-3 * leng / 8., length_x=leng / 4., length_y=10 * leng, length_z=10 * leng)
    p2.update({'vel(1)': vel, 'pres': ps, 'alpha_rho(1)': 2.4, 'alpha_rho(2)': 0.0, 'alpha(1)': 1.0, 'alpha(2)': 0.0,
               'alter_patch(1)': 'T'})
    p3 = dict(base, geometry=(10 if z_invariant else 8), x_centroid=0., radius=leng / 5.)
    p3.update({'vel(1)': 0., 'pres': 101325., 'alpha_rho(1)': 0.0, 'alpha_rho(2)': 0.167, 'alpha(1)': 0.0, 'alpha(2)': 1.0,
               'alter_patch(1)': 'T'})
    if smooth:
        # a sharp sphere always has 1-2 cell wide chords near its caps, where component-wise
        # WENO5 of (1.29 -> 0) and (0 -> 0.167) undershoots to a negative mixture density;
        # smear the interface over a few cells as 3-D MFC cases customarily do
        p3.update({'smoothen': 'T', 'smooth_patch_id': 1, 'smooth_coeff': 0.5})
    for i, pt in enumerate((p1, p2, p3), start=1):
        for k, v in pt.items():
            d[f'patch_icpp({i})%{k}'] = v
    return d


def config(d: Dict) -> CaseConfig:
    c = parse_case(d)
    c.check()
    return c
