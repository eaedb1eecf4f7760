"""Oracle parity at sizes close to the benchmarked ones, and on a stretched grid.

tests/test_gpu_parity.py compares with the strict (single-threaded, -ffp-contract=off) oracle,
which limits it to ~1e4 cells.  Here the reference is the oracle's TIMING build (the same
source, -O3 -march=native -fopenmp: all host threads), first cross-checked against the strict
build on a small case, then run at 256^3 x 5 steps and 1024^2 x 10 steps -- sizes where every
launch-geometry decision of the sweeps (x tiles and row blocks, pencil segments, partly filled
waves) differs from the small cases.  Gate: 1e-10, the north-star tolerance.

The stretched-grid cases run the per-cell coefficient tables (COEF = 1 kernels,
m_weno.fpp:168-363) that a uniform grid never touches in the fast build."""
import numpy as np
import pytest

from microfc_b200 import cases

from common import gpu_run, norm_linf, oracle_run, setup_case

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _advection(n):
    d = cases.advection_2d(N=n - 1)
    d['dt'] = d['dt'] * (100.0 / n)
    return d


def test_timing_oracle_equals_strict_oracle_to_rounding():
    cfg, cb, q0 = setup_case(cases.shockbubble_3d(nc=32), n_steps=5)
    a, _ = oracle_run(cfg, cb, q0, kind="strict")
    b, _ = oracle_run(cfg, cb, q0, kind="timing")
    assert (norm_linf(b, a, cfg) <= 1e-12).all()


@pytest.mark.parametrize("name,mk,steps", [
    ("shockbubble_3d_256", lambda: cases.shockbubble_3d(nc=256), 5),
    ("advection_2d_1024", lambda: _advection(1024), 10),
    ("shockbubble_2d_1536x512", lambda: cases.shockbubble_2d_cells(1536, 512), 10),
])
def test_fast_build_matches_the_oracle_at_size(name, mk, steps):
    cfg, cb, q0 = setup_case(mk(), n_steps=steps)
    q_ref, rows_ref = oracle_run(cfg, cb, q0, kind="timing")
    q_gpu, rows_gpu = gpu_run(cfg, cb, q0, strict=False)
    err = norm_linf(q_gpu, q_ref, cfg)
    print(f"SIZELINF {name} cells {int(np.prod(cfg.shape_glb))} steps {steps} max {err.max():.3e} gate {TOL:.0e}")
    assert np.isfinite(q_gpu).all()
    assert (err <= TOL).all(), err
    if cfg.run_time_info:
        assert abs(rows_gpu[-2][2][0] - rows_ref[-2][2][0]) <= 1e-9 * max(1.0, abs(rows_ref[-2][2][0]))


def _stretched(d, dirs):
    """Cluster cells towards the middle of the domain like m_grid.f90:171-187 (stretch_x, a_x, x_a, x_b)."""
    out = dict(d)
    for ax in dirs:
        lo, hi = d[f'{ax}_domain%beg'], d[f'{ax}_domain%end']
        out.update({f'stretch_{ax}': 'T', f'a_{ax}': 2.0, f'{ax}_a': lo + 0.3 * (hi - lo), f'{ax}_b': lo + 0.7 * (hi - lo), f'loops_{ax}': 1})
    out['dt'] = d['dt'] * 0.25                       # the smallest cells are ~3x narrower than the uniform ones
    return out


def _wide(d, patch, **kw):
    """The stretching of m_grid.f90:171-187 moves the end of a [0, 1] domain to ~1.4: let the
    background patches reach that far."""
    out = dict(d)
    for k, v in kw.items():
        out[f'patch_icpp({patch})%{k}'] = v
    return out


STRETCHED = {
    "sod_1d_stretched": lambda: _stretched(_wide(cases.sod_1d(), 2, x_centroid=1.5, length_x=2.0), "x"),
    "advection_2d_stretched": lambda: _stretched(_wide(cases.advection_2d(N=79), 1, length_x=4.0, length_y=4.0), "xy"),
    "shockbubble_3d_stretched": lambda: _stretched(cases.shockbubble_3d(ncx=40, ncy=36, ncz=34), "xyz"),
}


@pytest.mark.parametrize("name", list(STRETCHED))
def test_stretched_grid_tables(name):
    cfg, cb, q0 = setup_case(STRETCHED[name](), n_steps=40 if "3d" not in name else 10)
    w = np.diff(cb[0])
    assert w.max() / w.min() > 1.2, "the grid is not stretched"
    q_ref, _ = oracle_run(cfg, cb, q0)
    q_strict, _ = gpu_run(cfg, cb, q0, strict=True)
    assert np.array_equal(q_strict, q_ref), norm_linf(q_strict, q_ref, cfg)
    q_fast, _ = gpu_run(cfg, cb, q0, strict=False)
    err = norm_linf(q_fast, q_ref, cfg)
    print(f"SIZELINF {name} stretched max {err.max():.3e}")
    assert (err <= TOL).all(), err
