"""The CPU oracle against an independent, textbook-form statement of the same scheme
(tests/textbook_scheme.py: WENO5-JS, Toro's HLLC, the five-equation model, SSP-RK3 written from the
published algorithms, not from the reference's source structure).  With no reference-held fixtures
and no way to build the reference here this is the second, structurally different pin
of the oracle besides the analytic known answers: RHS and 20 RK3 steps agree to rounding (1e-11; the
reference's grid-dependent WENO coefficients equal the textbook constants only up to ~1e-15)."""
import numpy as np
import pytest

from microfc_b200 import cases

import oracle_lib
from common import norm_linf, oracle_run, setup_case
from textbook_scheme import Textbook

CASES = {
    "sod_1d": lambda: cases.sod_1d(),                                   # bc -3
    "kapila_1d": lambda: cases.kapila_1d(Nx=399),                       # water | air, stiffened gas
    "advection_2d": lambda: cases.advection_2d(N=63),                   # two fluids, smoothed interface
    "shockbubble_2d": lambda: cases.shockbubble_2d(Ny=40),              # bc -6 (extrapolation)
    "shearlayer_2d": lambda: cases.shearlayer_2d(Nx=79, Ny=39),         # periodic x, bc_y -5
    "shockdroplet_2d": lambda: cases.shockdroplet_2d(Nx=199, Ny=59),    # reflective y, water/air
    "three_fluids_2d": lambda: cases.three_fluids(cases.shockbubble_2d(Ny=40)),
    # the 3-D EXTENSION (the reference is 1-D/2-D): the oracle's direction-by-direction generalisation
    "shockbubble_3d": lambda: cases.shockbubble_3d(ncx=30, ncy=28, ncz=26),
    "shockbubble_3d_periodic_z": lambda: cases.shockbubble_3d(nc=26, periodic_z=True),
    # viscous terms (weno_Re_flux = F: central differences): Navier-Stokes stress with shear and bulk
    # viscosity, different Reynolds numbers per fluid, every component active
    "viscous_wave_2d": lambda: (cases.viscous_wave_2d(weno_Re_flux=False), cases.viscous_wave_state),
    "viscous_wave_2d_extrap_y": lambda: (cases.viscous_wave_2d(weno_Re_flux=False, bc_y=-6), cases.viscous_wave_state),
    "shockdroplet_2d_viscous": lambda: cases.shockdroplet_2d(Nx=199, Ny=59, viscous=True),
}


def _textbook(cfg, cb):
    dx = [float(cb[d][1] - cb[d][0]) for d in range(cfg.num_dims)]
    return Textbook(cfg.num_fluids, cfg.num_dims, cfg.gamma, cfg.pi_inf, dx, cfg.bc, cfg.weno_eps,
                    Re=cfg.Re if cfg.viscous else None)


@pytest.mark.parametrize("name", list(CASES))
def test_rhs_and_steps_agree_with_the_textbook_scheme(name):
    cfg, cb, q0 = setup_case(CASES[name](), n_steps=20)
    assert all(c != -4 for side in cfg.bc[:cfg.num_dims] for c in side)
    tb = _textbook(cfg, cb)
    o = oracle_lib.Oracle(cfg, cb)
    o.set_q(q0)
    drop = (slice(None), 0) if cfg.num_dims < 3 else (slice(None),)      # 1-D / 2-D arrays carry a z axis of length 1
    rhs_o = o.compute_rhs(0)[drop]
    rhs_t = tb.rhs(q0[drop])
    nf, nd = cfg.num_fluids, cfg.num_dims
    diff = np.abs(rhs_t - rhs_o).reshape(cfg.sys_size, -1).max(axis=1)
    scale = np.abs(rhs_o).reshape(cfg.sys_size, -1).max(axis=1)
    scale[nf:nf + nd] = scale[nf:nf + nd].max()
    qmax = np.abs(q0).reshape(cfg.sys_size, -1).max(axis=1)
    qmax[nf:nf + nd] = max(qmax[nf:nf + nd].max(), 1e-300)
    # Per variable: the two RHS agree to 1e-11 of the RHS's own scale -- or, where the RHS is the small
    # difference of large fluxes (a nearly uniform flow: the rounding of p ~ 1e5 fluxes shows at 1e-8 of
    # a ~1e-3 RHS) or pure round-off (a uniform volume fraction), their difference changes the variable
    # by less than 1e-15 of its magnitude per time step.
    # (stiffened-gas liquids: p = (E - rho u^2/2 - Pi)/Gamma with Pi ~ 1e9, p ~ 1e5 carries 1e-12 of relative
    # rounding noise, whose cell-to-cell differences ARE the y-momentum RHS of a horizontal shear layer)
    per_step = 1e-12 if max(cfg.pi_inf[:nf]) > 0 else 1e-15
    ok = (diff <= 1e-11 * scale) | (cfg.dt * diff <= per_step * qmax)
    assert ok.all(), (diff / np.where(scale == 0, 1.0, scale), cfg.dt * diff / qmax)
    q_o, _ = oracle_run(cfg, cb, q0)
    q_t = q0[drop].copy()
    for _ in range(20):
        q_t = tb.step(q_t, cfg.dt)
    # water/air: the cancellation in p = (E - ...)/Gamma amplifies the rounding differences
    tol = 1e-9 if max(cfg.pi_inf[:cfg.num_fluids]) > 0 else 1e-11
    q_t = q_t.reshape(q_o.shape)
    assert (norm_linf(q_t, q_o, cfg) <= tol).all(), norm_linf(q_t, q_o, cfg)
