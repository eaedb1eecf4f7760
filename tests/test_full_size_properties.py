"""BASELINE.json's configurations at their FULL sizes on the GPU, checked through properties
that do not depend on the size (the CPU oracle needs minutes to hours at these sizes; the same
properties pin the oracle itself at small sizes in tests/test_oracle_known_answers.py):

* configs[1]  2D_advection at 1024^2: the quasi-conservative 5-equation scheme keeps pressure
              and velocity uniform to round-off while the interface moves;
* configs[2]  2D_shockbubble at 4096^2: mirror symmetry in y;
* configs[4]  3-D shock-bubble at 512^3: mirror symmetry in y and z and the y<->z exchange
              symmetry of the sphere;
* periodic 2048 x 1024 two-fluid shear layer: discrete conservation of partial densities,
  momenta and energy (flux-difference form);
* the device-generated initial condition + a repeated run is deterministic bit for bit.

All through the C ABI (fast build: the one bench.py measures), initial condition laid out on the
device (mfc_b200_generate_initial_condition)."""
import dataclasses

import numpy as np
import pytest

from microfc_b200 import cases, pre_process

from common import norm_linf

pytestmark = pytest.mark.gpu


def run(case_dict, n_steps, prim=False, strict=False, keep_q0=True):
    from microfc_b200.simulation import Simulation
    cfg = dataclasses.replace(cases.config(case_dict), t_step_stop=n_steps)
    cb = pre_process.generate_grid(cfg)
    sim = Simulation(cfg, cb, strict=strict)
    try:
        sim.generate_initial_condition(cb)
        q0 = sim.download() if keep_q0 else None
        rows = sim.run()
        q = sim.download()
        p = sim.download_prim() if prim else None
    finally:
        sim.close()
    assert np.isfinite(q).all()
    return cfg, cb, q0, q, p, rows


def test_advection_1024_keeps_pressure_and_velocity_uniform():
    cfg, cb, q0, q, prim, rows = run(cases.advection_2d(N=1023), 40, prim=True)
    nf, nd = cfg.num_fluids, cfg.num_dims
    u, v, p = prim[nf], prim[nf + 1], prim[nf + nd]
    assert np.abs(u / 100.0 - 1).max() < 1e-9
    assert np.abs(v / 100.0 - 1).max() < 1e-9
    assert np.abs(p / 1e5 - 1).max() < 1e-9
    assert np.abs(q[0] - q0[0]).max() > 1e-3                      # the interface really moved
    assert all(0.0 < r[2][0] < 1.0 for r in rows)                 # ICFL rows of run_time.inf


def test_shockbubble_4096_is_symmetric_in_y():
    cfg, cb, q0, q, _, _ = run(cases.shockbubble_2d_cells(4096, 4096), 12)
    assert np.allclose(cb[1], -cb[1][::-1], atol=1e-15)
    sign = np.ones(cfg.sys_size)
    sign[cfg.num_fluids + 1] = -1.0                                # y-momentum is odd
    err = norm_linf(q[:, :, ::-1, :] * sign[:, None, None, None], q, cfg)
    assert (err < 1e-11).all(), err
    assert np.abs(q - q0).max() > 0


def test_shockbubble_3d_512_symmetries():
    cfg, cb, _, q, _, _ = run(cases.shockbubble_3d(nc=512), 3, keep_q0=False)   # 8.6 GB on the host: no second copy
    nf = cfg.num_fluids
    E = cfg.sys_size
    mom_scale = max(np.abs(q[nf + d]).max() for d in range(3))
    perm = list(range(E))
    perm[nf + 1], perm[nf + 2] = nf + 2, nf + 1                    # y <-> z exchange swaps the v and w momenta
    for v in range(E):                                             # one variable at a time (1 GiB each)
        b = q[v]
        scale = mom_scale if nf <= v < nf + 3 else max(np.abs(b).max(), 1e-300)
        sy = -1.0 if v == nf + 1 else 1.0
        sz = -1.0 if v == nf + 2 else 1.0
        assert np.abs(sy * b[:, ::-1, :] - b).max() / scale < 1e-11, ("y mirror", v)
        assert np.abs(sz * b[::-1, :, :] - b).max() / scale < 1e-11, ("z mirror", v)
        assert np.abs(q[perm[v]].transpose(1, 0, 2) - b).max() / scale < 1e-11, ("y<->z", v)
    # not vacuous: the shock has started to wrap around the bubble (transverse momentum is 0 at t = 0)
    assert np.abs(q[nf + 1]).max() > 1e-6 * mom_scale


def test_periodic_2048x1024_conserves_mass_momentum_energy():
    d = cases.shearlayer_2d(Nx=2047, Ny=1023, Nt=20)         # dx = dy like the shipped 320 x 160
    d['bc_y%beg'] = -1
    d['bc_y%end'] = -1
    cfg, cb, q0, q, _, _ = run(d, 20)
    nf, nd = cfg.num_fluids, cfg.num_dims
    for v in list(range(nf)) + list(range(nf, nf + nd + 1)):       # partial densities, momenta, energy
        s0, s1 = np.sum(q0[v], dtype=np.longdouble), np.sum(q[v], dtype=np.longdouble)
        scale = np.abs(q0[nf:nf + nd]).sum() if nf <= v < nf + nd else np.abs(q0[v]).sum()
        assert abs(float(s1 - s0)) / scale < 1e-13, (v, s0, s1)
    assert np.abs(q - q0).max() > 0


def test_repeated_run_is_bitwise_deterministic():
    a = run(cases.shockbubble_2d_cells(1024, 1024), 10)[3]
    b = run(cases.shockbubble_2d_cells(1024, 1024), 10)[3]
    assert np.array_equal(a, b)
