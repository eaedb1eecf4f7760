"""the reference's pre_process (patch geometries, smoothing, patch
permissions, primitive -> conservative) evaluated on the device by
``mfc_b200_generate_initial_condition`` against the host restatement
``microfc_b200.pre_process`` (numpy, statement by statement after m_create_patches.fpp /
m_assign_patches.fpp) on every case of the parity suite plus synthetic patch stacks that
exercise each geometry, ``alter_patch`` and ``smooth_patch_id``.

Bar: hard-edged patches BITWISE (the kernel is built without FMA contraction and keeps the
Fortran operand order); smoothed boundaries within 4 ulp of the blended value -- tanh comes
from the CUDA math library instead of glibc (<= 2 ulp apart), everything else is identical."""
import dataclasses

import numpy as np
import pytest

from microfc_b200 import cases, pre_process
from microfc_b200.case import Patch

from common import setup_case
from test_gpu_parity import CASES

pytestmark = pytest.mark.gpu


def device_ic(cfg, cb):
    from microfc_b200.simulation import Simulation
    sim = Simulation(cfg, cb)
    try:
        sim.generate_initial_condition(cb)
        return sim.download()
    finally:
        sim.close()


def check(cfg, cb, q_ref):
    q = device_ic(cfg, cb)
    assert q.shape == q_ref.shape and np.isfinite(q).all()
    smooth = any(p.smoothen or p.geometry in (7, 15) for p in cfg.patches)    # tanh / exp: CUDA vs glibc, <= 2 ulp
    if not smooth:
        assert np.array_equal(q, q_ref)
        return
    scale = np.abs(q_ref).reshape(q_ref.shape[0], -1).max(axis=1).reshape((-1,) + (1,) * (q.ndim - 1))
    err = np.abs(q - q_ref) / np.where(scale == 0, 1.0, scale)
    assert err.max() <= 4 * np.finfo(np.float64).eps, err.max()
    # away from the smeared interfaces (eta saturated to exactly 0 or 1) the fields are identical
    assert (q == q_ref).mean() > 0.5


@pytest.mark.parametrize("name", [n for n in CASES if ("weno" not in n or n.startswith("viscous")) and "viscous_wave" not in n])   # viscous_wave: state not from patches
def test_device_initial_condition_matches_pre_process(name):
    cfg, cb, q0 = setup_case(CASES[name]())
    check(cfg, cb, q0)


def _stack(nd):
    """Every geometry of this dimensionality in one patch list, with overwrite permissions and
    smearing against an earlier patch (m_create_patches.fpp:131-137)."""
    base = cases.config(cases.advection_2d(N=63) if nd == 2 else cases.shockbubble_3d(nc=32))
    def mk(geo, **kw):
        p = Patch(geometry=geo, alpha_rho=[0.3, 0.7, 0, 0], alpha=[0.25, 0.75, 0, 0], vel=[1.0, -2.0, 0.5], pres=1.5)
        for k, v in kw.items():
            setattr(p, k, v)
        return p
    lo = [base.domain[d][0] for d in range(3)]
    hi = [base.domain[d][1] for d in range(3)]
    cx, cy, cz = [(lo[d] + hi[d]) / 2 for d in range(3)]
    L = hi[0] - lo[0]
    ps = [mk(3 if nd == 2 else 9, x_centroid=cx, y_centroid=cy, z_centroid=cz, length_x=2 * L, length_y=2 * L, length_z=2 * L,
             alter_patch={0: True})]
    ps.append(mk(2 if nd == 2 else 8, x_centroid=cx - 0.1 * L, y_centroid=cy, z_centroid=cz, radius=0.2 * L, smoothen=True,
                 smooth_patch_id=1, smooth_coeff=0.5, alpha_rho=[1.0, 0.1, 0, 0], alpha=[0.9, 0.1, 0, 0], pres=2.0,
                 alter_patch={0: False, 1: True}))
    ps.append(mk(4, x_centroid=cx + 0.25 * L, y_centroid=cy, normal=[1.0, 0.5, 0.0], smoothen=True, smooth_patch_id=1,
                 smooth_coeff=0.4, alpha_rho=[0.2, 0.2, 0, 0], alpha=[0.5, 0.5, 0, 0], alter_patch={0: False, 1: True, 2: False}))
    ps.append(mk(5, x_centroid=cx, y_centroid=cy + 0.2 * L, radii=[0.15 * L, 0.05 * L, 0.0], pres=3.0,
                 alter_patch={0: False, 1: True, 2: True, 3: False}))
    ps.append(mk(18, x_centroid=cx, y_centroid=cy - 0.2 * L, radius=0.1 * L, epsilon=0.03 * L, vel=[0.0, 0.0, 0.0],
                 alter_patch={0: False, 1: True, 2: True, 3: True, 4: True}))
    ps.append(mk(1, x_centroid=lo[0] + 0.05 * L, length_x=0.06 * L, pres=0.1,
                 alter_patch={0: False, 1: True, 2: False, 3: False, 4: False, 5: True}))
    if nd == 2:                                      # isentropic vortex (a hard circle) and the 2-D analytical patch
        ps.append(mk(6, x_centroid=cx - 0.3 * L, y_centroid=cy - 0.3 * L, radius=0.08 * L, pres=4.0, vel=[3.0, 1.0, 0.0],
                     alter_patch={k: True for k in range(7)}))
        ps.append(mk(7, x_centroid=cx + 0.3 * L, y_centroid=cy - 0.25 * L, length_x=0.3 * L, length_y=0.2 * L, pres=2.5,
                     alter_patch={k: True for k in range(8)}))
    if nd == 3:
        ps.append(mk(10, x_centroid=cx + 0.3 * L, y_centroid=cy + 0.3 * L, radius=0.07 * L, pres=7.0,
                     alter_patch={k: True for k in range(7)}))
    for i, p in enumerate(ps):
        if not p.smoothen:
            p.smooth_patch_id = i + 1
    return dataclasses.replace(base, num_patches=len(ps), patches=ps)


@pytest.mark.parametrize("nd", [2, 3])
def test_every_geometry_permission_and_smearing(nd):
    cfg = _stack(nd)
    cb = pre_process.generate_grid(cfg)
    q_ref = pre_process.generate_initial_condition(cfg, cb)
    check(cfg, cb, q_ref)


def test_1d_analytical_patch():
    """geometry 15 (s_1D_analytical, m_create_patches.fpp:424-473): a line segment whose pressure
    carries a Gaussian bump evaluated at the right cell boundaries."""
    base = cases.config(cases.sod_1d(Nx=199))
    ps = [dataclasses.replace(p) for p in base.patches]
    ps[0].geometry = 15
    ps[0].x_centroid, ps[0].length_x = 0.4, 0.6
    ps[0].alter_patch = {0: True, 1: True, 2: True}
    cfg = dataclasses.replace(base, patches=ps)
    cb = pre_process.generate_grid(cfg)
    q_ref = pre_process.generate_initial_condition(cfg, cb)
    flat = pre_process.generate_initial_condition(base, cb)
    assert np.abs(q_ref[2] - flat[2]).max() > 0.05 * np.abs(flat[2]).max()      # the bump is there
    check(cfg, cb, q_ref)


def test_unsupported_geometry_fails_loudly():
    from microfc_b200 import abi
    from microfc_b200.simulation import Simulation
    cfg = cases.config(cases.advection_2d(N=31))
    ps = [dataclasses.replace(p) for p in cfg.patches]
    ps[-1].geometry = 11         # not a geometry of m_initial_condition.fpp:50-100
    cfg = dataclasses.replace(cfg, patches=ps)
    cb = pre_process.generate_grid(cfg)
    sim = Simulation(cfg, cb)
    try:
        with pytest.raises(abi.MfcB200Error) as e:
            sim.generate_initial_condition(cb)
        assert e.value.code == -7
    finally:
        sim.close()


def test_device_ic_then_steps_equal_uploaded_ic():
    """The generated state is a drop-in for pre_process + upload: stepping from it gives the
    same result as stepping from the uploaded host fields (hard-edged case: bitwise)."""
    from microfc_b200.simulation import Simulation
    cfg, cb, q0 = setup_case(cases.sod_1d(), n_steps=20)
    out = []
    for dev in (False, True):
        sim = Simulation(cfg, cb, strict=True)
        try:
            if dev:
                sim.generate_initial_condition(cb)
            else:
                sim.upload(sim.scatter(q0))
            sim.run()
            out.append(sim.download())
        finally:
            sim.close()
    assert np.array_equal(out[0], out[1])
