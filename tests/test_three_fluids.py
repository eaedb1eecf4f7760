"""num_fluids = 3 (the reference allows up to num_fluids_max = 10, every shipped example uses 1 or 2;
the sweeps are instantiated for 1..MFC_B200_BUILT_FLUIDS = 3): three-fluid variants of the 1-D, 2-D,
3-D and viscous parity cases against the oracle -- strict build bitwise, fast build within the gate --
and the device-side initial condition with three fluids."""
import numpy as np
import pytest

from microfc_b200 import cases

from common import gpu_run, norm_linf, oracle_run, setup_case

pytestmark = pytest.mark.gpu

CASES = {
    "kapila_1d": (lambda: cases.three_fluids(cases.kapila_1d(Nx=399)), 100, 1e-10),
    "shockbubble_2d": (lambda: cases.three_fluids(cases.shockbubble_2d(Ny=60)), 100, 1e-10),
    "shockbubble_3d": (lambda: cases.three_fluids(cases.shockbubble_3d(ncx=40, ncy=34, ncz=30)), 20, 1e-10),
    # water/air: the documented bound of the two-fluid case (tests/test_gpu_parity.py: FAST_TOL)
    "shockdroplet_2d_viscous": (lambda: cases.three_fluids(cases.shockdroplet_2d(Nx=199, Ny=59, viscous=True)), 100, 2e-9),
}


@pytest.mark.parametrize("name", list(CASES))
def test_three_fluids_match_the_oracle(name):
    mk, steps, tol = CASES[name]
    cfg, cb, q0 = setup_case(mk(), n_steps=steps)
    assert cfg.num_fluids == 3 and cfg.sys_size == 2 * 3 + cfg.num_dims + 1
    q_ref, _ = oracle_run(cfg, cb, q0)
    q_strict, _ = gpu_run(cfg, cb, q0, strict=True)
    assert np.array_equal(q_strict, q_ref), norm_linf(q_strict, q_ref, cfg)
    q_fast, _ = gpu_run(cfg, cb, q0, strict=False)
    err = norm_linf(q_fast, q_ref, cfg)
    print(f"FASTLINF three_fluids_{name} max {err.max():.3e} per-variable {norm_linf(q_fast, q_ref).max():.3e} gate {tol:.1e}")
    assert (err <= tol).all(), err


def test_three_fluid_initial_condition_on_the_device():
    from microfc_b200.simulation import Simulation
    cfg, cb, q0 = setup_case(cases.three_fluids(cases.shockbubble_2d(Ny=40)))
    sim = Simulation(cfg, cb)
    try:
        sim.generate_initial_condition(cb)
        q = sim.download()
    finally:
        sim.close()
    assert np.array_equal(q, q0)
