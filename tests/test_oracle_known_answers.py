"""Pins the CPU oracle (oracle/) with first-principles known answers.

The reference ships no tests, golden vectors or fixtures for this path and cannot be compiled
here -- "parity unpinned" -- so the restatement is pinned by what can be
derived independently of any implementation:

  1. examples/1D_sodshocktube against the exact Riemann solution
  2. examples/2D_advection: pressure / velocity equilibrium preserved to round-off
  3. discrete conservation with periodic boundaries
  4. y-symmetry of examples/2D_shockbubble
  5. WENO coefficients: classical Jiang-Shu constants on uniform grids, polynomial exactness
     on stretched grids (m_weno.fpp:168-363)
  6. HLLC consistency: a uniform state has zero RHS; a supersonic uniform advection is upwind
  7. 3-D extension: z-invariant data reproduces the 2-D run plane by plane (all three axes)
  8. emulated multi-rank run == single-rank run bitwise (m_mpi_proxy.fpp halo semantics)
  9. the -O3/OpenMP timing build (bench.py's cpu_baseline) agrees with the strict build
 10. viscous terms (both weno_Re_flux branches): a low-Mach sinusoidal shear wave decays like
     exp(-nu k^2 t), nu = 1/(Re rho); the VCFL row of run_time.inf is dt/(Re dx^2)
 11. order of accuracy: a smooth density wave advected with uniform u, p converges at fifth
     order in the cell averages (WENO5 reconstruction + flux differencing + RK3 at small dt);
     second order for WENO3-JS + RK2, first order for WENO1 + RK1
 12. the symmetry boundary (bc = -2, m_rhs.fpp:704-720,822-835): the upper half of the symmetric
     shock-bubble run on a half domain with a reflecting wall equals the full-domain run
 13. the alpha div(u) source (m_rhs.fpp:582-586): d(alpha)/dt = -div(alpha u) + alpha div(u) is
     the advection equation, so a UNIFORM volume fraction stays uniform in a compressing flow
 14. examples/1D_kapilashocktube (water | air, stiffened gas) against the exact two-material
     Riemann solution: star pressure and velocity, contact and shock positions, density in L1
 15. stretched grids: the smoothness indicators are Jiang & Shu's integral definition for the
     quadratic through the stencil's cell averages (beta_coef, m_weno.fpp:283-345)
 16. three fluids: a third fluid that duplicates the first reproduces the two-fluid solution
(tests/test_oracle_vs_textbook.py adds the structurally independent pin: a textbook-form numpy
statement of the whole scheme, inviscid and viscous, 1-D to 3-D.)
"""
import dataclasses

import numpy as np
import pytest

from microfc_b200 import cases, pre_process

import oracle_lib
from common import norm_linf, oracle_run, setup_case


# ---- 1. Sod ------------------------------------------------------------------------------------
def sod_exact(x, t, g=1.4, rl=1.0, pl=1.0, rr=0.125, pr=0.1, x0=0.5):
    """Exact solution of the Sod problem (Toro, ch. 4), density only."""
    cl, cr = np.sqrt(g * pl / rl), np.sqrt(g * pr / rr)

    def f(p, rk, pk, ck):
        if p > pk:
            A, B = 2 / ((g + 1) * rk), (g - 1) / (g + 1) * pk
            return (p - pk) * np.sqrt(A / (p + B))
        return 2 * ck / (g - 1) * ((p / pk) ** ((g - 1) / (2 * g)) - 1)

    lo, hi = 1e-8, 10.0
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if f(mid, rl, pl, cl) + f(mid, rr, pr, cr) > 0:
            hi = mid
        else:
            lo = mid
    ps = 0.5 * (lo + hi)
    us = 0.5 * (f(ps, rr, pr, cr) - f(ps, rl, pl, cl))
    rsl = rl * (ps / pl) ** (1 / g)
    csl = cl * (ps / pl) ** ((g - 1) / (2 * g))
    rsr = rr * ((ps / pr + (g - 1) / (g + 1)) / ((g - 1) / (g + 1) * ps / pr + 1))
    S = cr * np.sqrt((g + 1) / (2 * g) * ps / pr + (g - 1) / (2 * g))
    xi = (x - x0) / t
    rho = np.where(xi < -cl, rl, 0.0)
    fan = (xi >= -cl) & (xi < us - csl)
    rho = np.where(fan, rl * (2 / (g + 1) + (g - 1) / ((g + 1) * cl) * (-xi)) ** (2 / (g - 1)), rho)
    rho = np.where((xi >= us - csl) & (xi < us), rsl, rho)
    rho = np.where((xi >= us) & (xi < S), rsr, rho)
    rho = np.where(xi >= S, rr, rho)
    return rho, ps, us, S


def test_sod_matches_exact_riemann_solution():
    cfg, cb, q0 = setup_case(cases.sod_1d(), n_steps=1000)       # t = 0.1
    q, rows = oracle_run(cfg, cb, q0)
    x = 0.5 * (cb[0][1:] + cb[0][:-1])
    t_end = sum(r[1] for r in rows[:-1])
    assert abs(t_end - 0.1) < 1e-12
    rho_ex, ps, us, S = sod_exact(x, t_end)
    assert abs(ps - 0.30313) < 1e-4 and abs(us - 0.92745) < 1e-4
    rho = q[0, 0, 0]
    dx = x[1] - x[0]
    l1 = np.abs(rho - rho_ex).sum() * dx
    assert l1 < 2.5e-3, l1                                         # WENO5 at 400 cells: ~1.5e-3
    # shock position: where density crosses the middle of the jump
    mid = 0.5 * (0.125 + 0.26557)
    i_shock = np.where(rho[200:] > mid)[0].max() + 200
    assert abs(x[i_shock] - (0.5 + S * t_end)) < 2 * dx
    # plateau values between contact and shock / fan and contact
    assert abs(rho[np.argmin(np.abs(x - (0.5 + 0.5 * (us + S) * t_end)))] - 0.26557) < 2e-3
    assert abs(rho[np.argmin(np.abs(x - (0.5 + 0.5 * (us - 0.5) * t_end)))] - 0.42632) < 5e-3
    # ICFL row of run_time.inf is the analytic (|u|+c) dt/dx of the initial state on step 0
    assert abs(rows[0][2][0] - np.sqrt(1.4) * cfg.dt / dx) < 1e-12


# ---- 2. interface advection --------------------------------------------------------------------
def test_advection_keeps_pressure_and_velocity_uniform():
    cfg, cb, q0 = setup_case(cases.advection_2d(N=49), n_steps=50)
    o = oracle_lib.Oracle(cfg, cb)
    o.set_q(q0)
    oracle_lib.run_p_main(o, cfg)
    prim = o.get_prim()
    nf, nd = cfg.num_fluids, cfg.num_dims
    u, v, p = prim[nf], prim[nf + 1], prim[nf + nd]
    assert np.abs(u / 100.0 - 1).max() < 1e-9
    assert np.abs(v / 100.0 - 1).max() < 1e-9
    assert np.abs(p / 1e5 - 1).max() < 1e-9
    # the interface really moved (the test is not vacuous)
    assert np.abs(o.get_q()[0] - q0[0]).max() > 1e-3


# ---- 3. conservation ---------------------------------------------------------------------------
def test_conservation_with_periodic_boundaries():
    d = cases.shearlayer_2d(Nx=39, Ny=39, Nt=40)
    d['bc_y%beg'] = -1
    d['bc_y%end'] = -1
    cfg, cb, q0 = setup_case(d, n_steps=40)
    q, _ = oracle_run(cfg, cb, q0)
    nf, nd = cfg.num_fluids, cfg.num_dims
    for v in list(range(nf)) + list(range(nf, nf + nd + 1)):       # partial densities, momenta, energy
        s0, s1 = q0[v].sum(), q[v].sum()
        scale = np.abs(q0[nf:nf + nd]).sum() if nf <= v < nf + nd else np.abs(q0[v]).sum()
        assert abs(s1 - s0) / scale < 1e-13, (v, s0, s1)
    assert np.abs(q - q0).max() > 0


# ---- 4. symmetry -------------------------------------------------------------------------------
def test_shockbubble_is_symmetric_in_y():
    cfg, cb, q0 = setup_case(cases.shockbubble_2d(Ny=40), n_steps=60)
    assert np.allclose(cb[1], -cb[1][::-1], atol=1e-15)
    q, _ = oracle_run(cfg, cb, q0)
    nf = cfg.num_fluids
    sign = np.ones(cfg.sys_size)
    sign[nf + 1] = -1.0                                            # y-momentum is odd
    qm = q[:, :, ::-1, :] * sign[:, None, None, None]
    err = norm_linf(qm, q, cfg)
    assert (err < 1e-11).all(), err


# ---- 5. WENO coefficients ----------------------------------------------------------------------
def test_weno_coefficients_uniform_grid_are_jiang_shu():
    cfg, cb, _ = setup_case(cases.sod_1d(Nx=63), n_steps=1)
    o = oracle_lib.Oracle(cfg, cb)
    c = o.weno_coefficients(0, 0)
    assert np.allclose(c["d_R"], [[0.3, 0.6, 0.1]], atol=1e-13)
    assert np.allclose(c["d_L"], [[0.1, 0.6, 0.3]], atol=1e-13)
    # classical smoothness indicators in first-difference form, for v = x^2 on dx = 1:
    # dvd = (2j+1) ..., checked through the generic exactness test below; here: symmetry
    assert np.allclose(c["poly_R"][:, ::-1, ::-1], -c["poly_L"], atol=1e-13)


def _reconstruct(c, v, j, cell0):
    """m_weno.fpp:476-531 for one cell with the oracle's coefficient arrays."""
    i = j - cell0
    dvd = {1: v[j + 2] - v[j + 1], 0: v[j + 1] - v[j], -1: v[j] - v[j - 1], -2: v[j - 1] - v[j - 2]}
    out = []
    for poly in (c["poly_L"], c["poly_R"]):
        out.append([v[j] + poly[i, 0, 0] * dvd[1] + poly[i, 0, 1] * dvd[0],
                    v[j] + poly[i, 1, 0] * dvd[0] + poly[i, 1, 1] * dvd[-1],
                    v[j] + poly[i, 2, 0] * dvd[-1] + poly[i, 2, 1] * dvd[-2]])
    bt = c["beta"][i]
    beta = [bt[0, 0] * dvd[1] ** 2 + bt[0, 1] * dvd[1] * dvd[0] + bt[0, 2] * dvd[0] ** 2,
            bt[1, 0] * dvd[0] ** 2 + bt[1, 1] * dvd[0] * dvd[-1] + bt[1, 2] * dvd[-1] ** 2,
            bt[2, 0] * dvd[-1] ** 2 + bt[2, 1] * dvd[-1] * dvd[-2] + bt[2, 2] * dvd[-2] ** 2]
    return out[0], out[1], beta


def test_weno_polynomial_exactness_on_a_stretched_grid():
    cfg = cases.config(cases.sod_1d(Nx=39))
    rng = np.random.default_rng(3)
    w = 1.0 + 0.6 * rng.random(cfg.m + 1)                          # irregular cell widths
    cbx = np.concatenate([[0.0], np.cumsum(w)])
    cbx /= cbx[-1]
    o = oracle_lib.Oracle(cfg, [cbx])
    c = o.weno_coefficients(0, 0)
    cb_g, _, _ = o.rank_metrics(0, 0)
    b = cfg.buff_size
    lo = -b + 2                                                    # first cell with coefficients
    # cell averages of a quadratic: every 3-cell candidate stencil reproduces the face value exactly
    a0, a1, a2 = 0.3, -1.2, 2.5
    P = lambda x: a0 * x + a1 * x ** 2 / 2 + a2 * x ** 3 / 3
    f = lambda x: a0 + a1 * x + a2 * x ** 2
    left, right = cb_g[:-1], cb_g[1:]                              # cells -b .. m+b
    avg = (P(right) - P(left)) / (right - left)
    for cell in range(lo + 1, cfg.m + b - 2):
        j = cell + b                                               # array index of the cell
        pl, pr, beta = _reconstruct(c, avg, j, lo + b)
        for k in range(3):
            assert abs(pl[k] - f(left[j])) < 1e-11, (cell, k)
            assert abs(pr[k] - f(right[j])) < 1e-11, (cell, k)
        i = cell - lo
        # ideal weights turn the three quadratics into the 5-cell quartic-exact value
        b0, b1, b2, b3, b4 = 0.7, 0.2, -0.4, 1.1, 0.9
        P4 = lambda x: b0 * x + b1 * x ** 2 / 2 + b2 * x ** 3 / 3 + b3 * x ** 4 / 4 + b4 * x ** 5 / 5
        f4 = lambda x: b0 + b1 * x + b2 * x ** 2 + b3 * x ** 3 + b4 * x ** 4
        avg4 = (P4(right) - P4(left)) / (right - left)
        pl4, pr4, _ = _reconstruct(c, avg4, j, lo + b)
        assert abs(np.dot(c["d_L"][i], pl4) - f4(left[j])) < 1e-10
        assert abs(np.dot(c["d_R"][i], pr4) - f4(right[j])) < 1e-10
        assert abs(c["d_L"][i].sum() - 1) < 1e-14 and (c["d_L"][i] > 0).all()
        # smoothness indicators vanish for constants and are positive otherwise
        assert all(x > 0 for x in beta)


def test_weno_beta_on_a_stretched_grid_is_the_jiang_shu_integral():
    """On a NON-uniform grid the smoothness indicator of candidate k must be Jiang & Shu's definition
    beta_k = sum_{l=1,2} int_cell h^(2l-1) (d^l p_k / dx^l)^2 dx for the quadratic p_k whose cell averages
    match the three cells of the stencil: pins beta_coef (m_weno.fpp:283-345) where the uniform-grid
    constants say nothing."""
    cfg = cases.config(cases.sod_1d(Nx=39))
    rng = np.random.default_rng(5)
    w = 1.0 + 0.6 * rng.random(cfg.m + 1)
    cbx = np.concatenate([[0.0], np.cumsum(w)])
    cbx /= cbx[-1]
    o = oracle_lib.Oracle(cfg, [cbx])
    c = o.weno_coefficients(0, 0)
    cb_g, _, _ = o.rank_metrics(0, 0)
    b = cfg.buff_size
    lo = -b + 2
    left, right = cb_g[:-1], cb_g[1:]
    v = rng.random(len(left))

    def integral(cells, j):
        A = [[(right[q] ** (p + 1) - left[q] ** (p + 1)) / ((p + 1) * (right[q] - left[q])) for p in range(3)] for q in cells]
        a = np.linalg.solve(np.array(A), v[list(cells)])           # p(x) = a0 + a1 x + a2 x^2
        xl, xr = left[j], right[j]
        h = xr - xl
        return (h * (a[1] ** 2 * (xr - xl) + 2 * a[1] * a[2] * (xr ** 2 - xl ** 2) + 4 * a[2] ** 2 * (xr ** 3 - xl ** 3) / 3)
                + h ** 3 * 4 * a[2] ** 2 * (xr - xl))
    for cell in range(lo + 1, cfg.m + b - 2):
        j = cell + b
        _, _, beta = _reconstruct(c, v, j, lo + b)
        want = [integral((j, j + 1, j + 2), j), integral((j - 1, j, j + 1), j), integral((j - 2, j - 1, j), j)]
        assert np.allclose(beta, want, rtol=1e-9, atol=1e-14), (cell, beta, want)


def test_weno_beta_uniform_grid_matches_classical_formula():
    cfg, cb, _ = setup_case(cases.sod_1d(Nx=31), n_steps=1)
    o = oracle_lib.Oracle(cfg, cb)
    c = o.weno_coefficients(0, 0)
    rng = np.random.default_rng(0)
    v = rng.random(cfg.m + 1 + 2 * cfg.buff_size)
    lo = -cfg.buff_size + 2
    j = 12
    _, _, beta = _reconstruct(c, v, j, lo + cfg.buff_size)
    vm2, vm1, v0, vp1, vp2 = v[j - 2:j + 3]
    js = [13 / 12 * (v0 - 2 * vp1 + vp2) ** 2 + 0.25 * (3 * v0 - 4 * vp1 + vp2) ** 2,
          13 / 12 * (vm1 - 2 * v0 + vp1) ** 2 + 0.25 * (vm1 - vp1) ** 2,
          13 / 12 * (vm2 - 2 * vm1 + v0) ** 2 + 0.25 * (vm2 - 4 * vm1 + 3 * v0) ** 2]
    assert np.allclose(beta, js, rtol=1e-11)


# ---- 6. HLLC consistency -----------------------------------------------------------------------
@pytest.mark.parametrize("name", ["advection_2d", "shockbubble_3d", "viscous_2d"])
def test_uniform_state_has_zero_rhs(name):
    d = {"advection_2d": lambda: cases.advection_2d(N=24), "shockbubble_3d": lambda: cases.shockbubble_3d(nc=25),
         "viscous_2d": lambda: cases.viscous_2d(N=25)}[name]()
    cfg, cb, q0 = setup_case(d, n_steps=2)
    q = np.broadcast_to(q0[:, :1, :1, :1], q0.shape).copy()        # the corner cell everywhere
    o = oracle_lib.Oracle(cfg, cb)
    o.set_q(q)
    rhs = o.compute_rhs(0)
    # round-off of (F_{j-1/2} - F_{j+1/2})/dx with |F_v| ~ |q_v| (|u| + c) (+ p for momentum, energy)
    prim = o.get_prim()
    nf, nd = cfg.num_fluids, cfg.num_dims
    rho = q[:nf].sum(axis=0).max()
    pres = np.abs(prim[nf + nd]).max()
    speed = np.abs(prim[nf:nf + nd]).max() + np.sqrt((1 + 1 / min(cfg.gamma[:nf])) * (pres + max(cfg.pi_inf[:nf])) / rho)
    fscale = np.abs(q).reshape(cfg.sys_size, -1).max(axis=1) * speed
    fscale[nf:nf + nd + 1] += pres * max(1.0, speed)
    fscale = np.where(fscale == 0, 1, fscale) / min(np.diff(c).min() for c in cb)
    assert (np.abs(rhs).reshape(cfg.sys_size, -1).max(axis=1) < 1e-13 * fscale).all()


def test_supersonic_advection_is_fully_upwind():
    """s_L > 0 => HLLC flux is the left physical flux: a density bump advected at Mach 3 moves
    without any upstream influence: cells upstream of the bump keep a zero RHS."""
    d = cases.sod_1d(Nx=99)
    cfg, cb, q0 = setup_case(d, n_steps=2)
    g = cfg.gamma[0]                                               # 1/(gamma-1)
    rho = np.ones(cfg.m + 1)
    rho[50:55] = 2.0
    u, p = 3.0 * np.sqrt(1.4), 1.0
    q = np.zeros_like(q0)
    q[0, 0, 0] = rho
    q[1, 0, 0] = rho * u
    q[2, 0, 0] = g * p + 0.5 * rho * u * u
    q[3, 0, 0] = 1.0
    o = oracle_lib.Oracle(cfg, cb)
    o.set_q(q)
    rhs = o.compute_rhs(0)
    assert np.abs(rhs[:, 0, 0, :47]).max() < 1e-10                 # 50 - 3 stencil cells: round-off of f/dx only
    assert np.abs(rhs[0, 0, 0, 48:60]).max() > 1.0


# ---- 7. 3-D extension vs 2-D -------------------------------------------------------------------
@pytest.mark.parametrize("axis", [0, 1, 2])
def test_3d_invariant_direction_reproduces_2d(axis):
    """Extrude the 2-D shock-bubble along one axis of a 3-D box (periodic in that axis): every
    plane must equal the 2-D solution.  axis = index of the invariant direction (0: x, 1: y, 2: z);
    the 2-D case's (x, y) are mapped onto the two remaining axes in order."""
    n2 = 25
    d2 = cases.shockbubble_2d_cells(3 * n2, n2, Nt=12)
    cfg2, cb2, q2 = setup_case(d2, n_steps=12)
    ref, _ = oracle_run(cfg2, cb2, q2)
    nf = cfg2.num_fluids
    ninv = 6
    d3 = cases.shockbubble_3d(nc=25, Nt=12)
    cfg3 = cases.config(d3)
    other = [a for a in (0, 1, 2) if a != axis]                    # 3-D axes carrying 2-D x and y
    N3 = [0, 0, 0]
    N3[axis] = ninv - 1
    N3[other[0]], N3[other[1]] = cfg2.m, cfg2.n
    bc3 = [None, None, None]
    bc3[axis] = [-1, -1]
    bc3[other[0]], bc3[other[1]] = list(cfg2.bc[0]), list(cfg2.bc[1])
    cfg3 = dataclasses.replace(cfg3, m=N3[0], n=N3[1], p=N3[2], bc=bc3, dt=cfg2.dt, t_step_stop=12,
                               gamma=cfg2.gamma, pi_inf=cfg2.pi_inf, run_time_info=cfg2.run_time_info)
    cb3 = [None, None, None]
    cb3[axis] = np.linspace(0.0, ninv * (cb2[0][1] - cb2[0][0]), ninv + 1)
    cb3[other[0]], cb3[other[1]] = cb2[0], cb2[1]
    # state (E3, z, y, x): 2-D field f[y2, x2] placed on axes (other[1], other[0])
    E3 = cfg3.sys_size
    shape3 = (N3[2] + 1, N3[1] + 1, N3[0] + 1)
    q3 = np.zeros((E3,) + shape3)

    def lift(f2):                                                  # f2[y2, x2] -> 3-D array
        a = f2.T if other[0] > other[1] else f2                    # never true (other is sorted)
        src_axes = {other[1]: 0, other[0]: 1}                      # 3-D axis -> axis of f2
        out = np.empty(shape3)
        idx = np.indices(shape3)                                   # idx[0]=z, idx[1]=y, idx[2]=x
        ax_of = {0: idx[2], 1: idx[1], 2: idx[0]}
        out[...] = a[ax_of[other[1]], ax_of[other[0]]]
        return out

    def map_var(v2):                                               # 2-D variable index -> 3-D index
        if v2 < nf:
            return v2
        if v2 < nf + 2:
            return nf + other[v2 - nf]
        return v2 + 1
    for v2 in range(cfg2.sys_size):
        q3[map_var(v2)] = lift(q2[v2, 0])
    out3, _ = oracle_run(cfg3, cb3, q3)
    for v2 in range(cfg2.sys_size):
        want = lift(ref[v2, 0])
        got = out3[map_var(v2)]
        den = np.abs(want).max()
        assert np.abs(got - want).max() <= 1e-12 * max(den, 1e-300) + (1e-9 if nf <= v2 < nf + 2 else 0), (axis, v2)
    assert np.abs(out3[nf + axis]).max() < 1e-9                    # no flow along the invariant axis


# ---- 8. emulated MPI ranks ---------------------------------------------------------------------
@pytest.mark.parametrize("name,nprocs", [("sod_1d", 2), ("sod_1d", 3), ("shockbubble_2d", 2), ("shockbubble_2d", 4),
                                         ("shearlayer_2d", 4), ("shockbubble_3d", 8), ("viscous_2d", 2)])
def test_multi_rank_equals_single_rank_bitwise(name, nprocs):
    d = {"sod_1d": lambda: cases.sod_1d(Nx=99), "shockbubble_2d": lambda: cases.shockbubble_2d_cells(100, 52, Nt=10),
         "shearlayer_2d": lambda: cases.shearlayer_2d(Nx=63, Ny=55), "shockbubble_3d": lambda: cases.shockbubble_3d(nc=52),
         "viscous_2d": lambda: cases.viscous_2d(N=59, Nt=6)}[name]()
    n = 3 if "3d" in name else 8
    cfg, cb, q0 = setup_case(d, n_steps=n)
    q1, r1 = oracle_run(cfg, cb, q0, num_procs=1)
    qn, rn = oracle_run(cfg, cb, q0, num_procs=nprocs)
    assert np.array_equal(q1, qn)
    if cfg.run_time_info:
        assert [r[2][0] for r in r1] == [r[2][0] for r in rn]


# ---- 9. timing build ---------------------------------------------------------------------------
def test_timing_build_agrees_with_strict_build():
    cfg, cb, q0 = setup_case(cases.shockbubble_2d(Ny=30), n_steps=30)
    qs, _ = oracle_run(cfg, cb, q0, kind="strict")
    qt, _ = oracle_run(cfg, cb, q0, kind="timing")
    assert oracle_lib.load("strict").orc_is_strict() == 1
    assert oracle_lib.load("timing").orc_is_strict() == 0
    assert (norm_linf(qt, qs, cfg) < 1e-10).all()


# ---- 10. viscous shear-wave decay --------------------------------------------------------------
@pytest.mark.parametrize("weno_Re_flux", [False, True])
def test_viscous_shear_wave_decays_at_the_analytic_rate(weno_Re_flux):
    """u(y, t) = U0 sin(2 pi y) exp(-nu (2 pi)^2 t) solves the viscous momentum equation for a
    uniform low-Mach state (div u = 0, pressure uniform up to O(Mach^2) heating).  Pins
    m_viscous.fpp / s_compute_viscous_source_flux in the oracle: tau_xy = (du/dy + dv/dx)/Re,
    rhs(mom_x) += d tau_xy / dy, with 1/Re = sum alpha_i/Re_i."""
    N, Nx, Re, U0, nsteps = 64, 32, 100.0, 0.01, 200
    d = cases.viscous_2d(N=N - 1, Nt=nsteps, weno_Re_flux=weno_Re_flux)
    d.update({'m': Nx - 1, 'n': N - 1, 'x_domain%beg': 0.0, 'x_domain%end': Nx / float(N), 'y_domain%beg': 0.0,
              'y_domain%end': 1.0, 'bc_x%beg': -1, 'bc_x%end': -1, 'bc_y%beg': -1, 'bc_y%end': -1, 'run_time_info': 'T',
              'fluid_pp(1)%gamma': 2.5, 'fluid_pp(2)%gamma': 2.5, 'fluid_pp(1)%pi_inf': 0.0, 'fluid_pp(2)%pi_inf': 0.0,
              'dt': 0.2 * (1.0 / N) / np.sqrt(1.4)})
    for i in (1, 2):
        d[f'fluid_pp({i})%Re(1)'] = Re                             # shear viscosity only
        d.pop(f'fluid_pp({i})%Re(2)', None)
    cfg = dataclasses.replace(cases.config(d), t_step_stop=nsteps)
    cb = pre_process.generate_grid(cfg)
    y = (cb[1][1:] + cb[1][:-1]) / 2
    q = np.zeros((cfg.sys_size, 1, N, Nx))
    u = U0 * np.sin(2 * np.pi * y)[None, :, None] * np.ones((1, N, Nx))
    q[0], q[1] = 0.5, 0.5                                          # rho = 1, two identical fluids
    q[2], q[3] = u, 0.0
    q[4] = 2.5 * 1.0 + 0.5 * u * u                                 # p = 1, gamma = 1.4
    q[5], q[6] = 0.5, 0.5
    o = oracle_lib.Oracle(cfg, cb)
    o.set_q(q)
    rows = oracle_lib.run_p_main(o, cfg)
    out = o.get_q()
    uu = out[2] / (out[0] + out[1])
    amp = 2 * np.mean(uu[0, :, 0] * np.sin(2 * np.pi * y))
    exact = np.exp(-(1.0 / Re) * (2 * np.pi) ** 2 * nsteps * cfg.dt)
    assert abs(amp / U0 - exact) < (1e-5 if weno_Re_flux else 5e-4) * exact, (amp / U0, exact)
    assert np.abs(uu - uu[:, :, :1]).max() == 0.0                  # x-invariant stays x-invariant, bitwise
    assert np.abs(out[3]).max() < 1e-5 * U0 * 100                  # no spurious transverse momentum
    # run_time.inf: VCFL = dt max(1/Re_1, 1/Re_2) / min(dx, dy)^2 (m_data_output.fpp:223-229)
    dx = 1.0 / N
    assert abs(rows[0][2][1] - cfg.dt / Re / dx ** 2) < 1e-12


@pytest.mark.parametrize("weno_Re_flux,bc_y", [(True, -1), (False, -1), (True, -6), (False, -6)])
def test_viscous_wave_case_exercises_every_viscous_term(weno_Re_flux, bc_y):
    """The parity case of the viscous path (cases.viscous_wave_2d, used by tests/test_gpu_parity.py,
    tests/nccl_worker.py and two golden vectors) must actually be moved by the viscous terms: in
    examples/2D_viscous the velocity is piecewise constant, the reconstructed gradients of the
    weno_Re_flux branch are exactly zero and a kernel that did nothing would pass.  Here the
    viscous run differs from the inviscid run of the same state by several per cent in both
    momenta and in the energy, and the shear-only / bulk-only runs differ from each other."""
    def run(keep):
        d = cases.viscous_wave_2d(N=32, Nx=26, Nt=40, weno_Re_flux=weno_Re_flux, bc_y=bc_y)
        for k in list(d):
            if '%Re(' in k and not k.endswith(keep):
                d.pop(k)
        cfg = cases.config(d)
        cb = pre_process.generate_grid(cfg)
        o = oracle_lib.Oracle(cfg, cb)
        o.set_q(cases.viscous_wave_state(cfg, cb))
        oracle_lib.run_p_main(o, cfg)
        return o.get_q()
    full, shear, bulk, none = run(")"), run("Re(1)"), run("Re(2)"), run("none")
    scale = np.abs(none[2:4]).max()
    for other in (none, shear, bulk):
        assert np.abs(full[2] - other[2]).max() > 1e-4 * scale       # x-momentum
        assert np.abs(full[3] - other[3]).max() > 1e-4 * scale       # y-momentum
    assert np.abs(full[4] - none[4]).max() > 1e-9 * np.abs(none[4]).max()
    assert np.isfinite(full).all()


# ---- 11. order of accuracy ---------------------------------------------------------------------
@pytest.mark.parametrize("weno_order,time_stepper,err40,order", [
    (5, 3, 5e-6, 4.7),      # WENO5-JS: fifth order
    (3, 2, 5e-3, 1.9),      # WENO3-JS with eps = 1e-16 falls to second order at the smooth extrema (known)
    (1, 1, 3e-2, 0.9),      # piecewise constant + Euler: first order
])
def test_smooth_advection_converges_at_the_design_order(weno_order, time_stepper, err40, order):
    """rho(x, 0) = 1 + 0.2 sin(2 pi x), u = 1, p = 1, periodic: the exact solution is the
    translated profile.  Cell AVERAGES are compared (the scheme is a finite-volume method:
    m_rhs.fpp:567-576 differences face fluxes), so the error is the scheme's alone."""
    errs = []
    for N in (40, 80, 160):
        d = cases.sod_1d(Nx=N - 1, Nt=10)
        d.update({'bc_x%beg': -1, 'bc_x%end': -1, 'x_domain%beg': 0.0, 'x_domain%end': 1.0,
                  'weno_order': weno_order, 'time_stepper': time_stepper})
        u0, T = 1.0, 0.25
        dt = 0.2 * (1.0 / N) / (u0 + np.sqrt(1.4 / 0.8))
        nsteps = int(round(T / dt))
        d['dt'] = T / nsteps
        cfg = dataclasses.replace(cases.config(d), t_step_stop=nsteps)
        cb = pre_process.generate_grid(cfg)
        xl, xr = cb[0][:-1], cb[0][1:]

        def avg(a, b):
            return 1.0 + 0.2 * (np.cos(2 * np.pi * a) - np.cos(2 * np.pi * b)) / (2 * np.pi * (b - a))
        rho = avg(xl, xr)
        q = np.zeros((cfg.sys_size, 1, 1, N))
        q[0], q[1], q[2], q[3] = rho, rho * u0, 2.5 + 0.5 * rho * u0 * u0, 1.0
        o = oracle_lib.Oracle(cfg, cb)
        o.set_q(q)
        oracle_lib.run_p_main(o, cfg)
        errs.append(np.abs(o.get_q()[0, 0, 0] - avg(xl - u0 * T, xr - u0 * T)).mean())
    orders = [np.log2(errs[i] / errs[i + 1]) for i in range(2)]
    assert errs[0] < err40 and all(o > order for o in orders), (errs, orders)


# ---- 12. reflecting wall -----------------------------------------------------------------------
def test_reflecting_wall_equals_the_mirrored_full_domain():
    full = cases.shockbubble_2d_cells(180, 60, Nt=100)
    cfgF = dataclasses.replace(cases.config(full), t_step_stop=60)
    cbF = pre_process.generate_grid(cfgF)
    q0F = pre_process.generate_initial_condition(cfgF, cbF)
    qF, _ = oracle_run(cfgF, cbF, q0F)
    half = dict(full)
    half.update({'n': 29, 'y_domain%beg': 0.0, 'bc_y%beg': -2})   # y in [0, 0.5], wall at y = 0
    cfgH = dataclasses.replace(cases.config(half), t_step_stop=60, dt=cfgF.dt)
    cbH = pre_process.generate_grid(cfgH)
    q0H = pre_process.generate_initial_condition(cfgH, cbH)
    assert np.array_equal(q0H, q0F[:, :, 30:, :])
    qH, _ = oracle_run(cfgH, cbH, q0H)
    err = norm_linf(qH, qF[:, :, 30:, :], cfgH)
    assert (err < 1e-11).all(), err
    assert np.abs(qH - q0H).max() > 0


# ---- 13. volume-fraction source ----------------------------------------------------------------
def test_uniform_volume_fraction_survives_compression():
    """Water/air mixture with uniform alpha in a sinusoidal velocity field: the densities change
    (div u != 0) but -div(alpha u) + alpha div(u) must cancel to round-off for uniform alpha."""
    d = cases.kapila_1d(Nx=199, Nt=100)
    d.update({'bc_x%beg': -1, 'bc_x%end': -1})
    cfg = dataclasses.replace(cases.config(d), t_step_stop=80)
    cb = pre_process.generate_grid(cfg)
    x = (cb[0][1:] + cb[0][:-1]) / 2
    N, a1 = cfg.m + 1, 0.3
    ar1, ar2 = 1000.0 * a1, 50.0 * (1 - a1)
    rho = ar1 + ar2
    u = 20.0 * np.sin(2 * np.pi * (x - cfg.domain[0][0]) / (cfg.domain[0][1] - cfg.domain[0][0]))
    G = a1 * cfg.gamma[0] + (1 - a1) * cfg.gamma[1]
    Pi = a1 * cfg.pi_inf[0] + (1 - a1) * cfg.pi_inf[1]
    q = np.zeros((cfg.sys_size, 1, 1, N))
    q[0], q[1], q[2], q[3], q[4], q[5] = ar1, ar2, rho * u, G * 1e5 + Pi + 0.5 * rho * u * u, a1, 1 - a1
    o = oracle_lib.Oracle(cfg, cb)
    o.set_q(q)
    oracle_lib.run_p_main(o, cfg)
    out = o.get_q()
    assert np.isfinite(out).all()
    assert np.abs(out[4] - a1).max() < 1e-14 and np.abs(out[5] - (1 - a1)).max() < 1e-14
    assert np.abs(out[0] + out[1] - rho).max() / rho > 1e-3       # the flow really compressed the mixture


# ---- 14. water / air shock tube ----------------------------------------------------------------
def stiffened_riemann_exact(x, t, x0, gl, pil, rl, pl, gr, pir, rr, pr):
    """Exact solution (density) of the Riemann problem between two stiffened gases at rest,
    p_l > p_r: left rarefaction, contact, right shock (Toro ch. 4 with p -> p + pi_inf)."""
    Pl, Pr = pl + pil, pr + pir
    cl, cr = np.sqrt(gl * Pl / rl), np.sqrt(gr * Pr / rr)

    def f(p, g, pik, rk, pk, ck):
        P, Pk = p + pik, pk + pik
        if p > pk:
            return (p - pk) * np.sqrt(2 / ((g + 1) * rk) / (P + (g - 1) / (g + 1) * Pk))
        return 2 * ck / (g - 1) * ((P / Pk) ** ((g - 1) / (2 * g)) - 1)

    lo, hi = pr, pl
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if f(mid, gl, pil, rl, pl, cl) + f(mid, gr, pir, rr, pr, cr) > 0:
            hi = mid
        else:
            lo = mid
    ps = 0.5 * (lo + hi)
    us = 0.5 * (f(ps, gr, pir, rr, pr, cr) - f(ps, gl, pil, rl, pl, cl))
    rsl = rl * ((ps + pil) / Pl) ** (1 / gl)
    csl = cl * ((ps + pil) / Pl) ** ((gl - 1) / (2 * gl))
    Pq, k = (ps + pir) / Pr, (gr - 1) / (gr + 1)
    rsr = rr * (Pq + k) / (k * Pq + 1)
    S = cr * np.sqrt((gr + 1) / (2 * gr) * Pq + (gr - 1) / (2 * gr))
    xi = (x - x0) / t
    rho = np.where(xi < -cl, rl, 0.0)
    fan = (xi >= -cl) & (xi < us - csl)
    rho = np.where(fan, rl * (2 / (gl + 1) + (gl - 1) / ((gl + 1) * cl) * (-xi)) ** (2 / (gl - 1)), rho)
    rho = np.where((xi >= us - csl) & (xi < us), rsl, rho)
    rho = np.where((xi >= us) & (xi < S), rsr, rho)
    rho = np.where(xi >= S, rr, rho)
    return rho, ps, us, S, rsr


def test_kapila_water_air_shock_tube_matches_exact_riemann_solution():
    N = 400
    cfg, cb, q0 = setup_case(cases.kapila_1d(Nx=N - 1, Nt=int(6025 * N / 1000)))
    o = oracle_lib.Oracle(cfg, cb)
    o.set_q(q0)
    oracle_lib.run_p_main(o, cfg)
    q, prim = o.get_q(), o.get_prim()
    x = (cb[0][1:] + cb[0][:-1]) / 2
    T, dx = cfg.t_step_stop * cfg.dt, 1.0 / N
    rho_ex, ps, us, S, rsr = stiffened_riemann_exact(x, T, 0.7, 4.4, 6e8, 1000., 1e9, 1.4, 0., 50., 1e5)
    rho = (q[0] + q[1])[0, 0]
    assert np.abs(rho - rho_ex).mean() / np.abs(rho_ex).mean() < 1e-2
    i_star = np.argmin(np.abs(x - (0.7 + 0.5 * us * T)))           # star region, water side
    assert abs(prim[3][0, 0][i_star] / ps - 1) < 5e-3              # p* = 14.19 MPa
    assert abs(prim[2][0, 0][i_star] / us - 1) < 1e-3              # u* = 482.6 m/s
    i0 = int(0.72 * N)
    i_shock = np.where(rho[i0:] > 0.5 * (rsr + 50.0))[0].max() + i0
    assert abs(x[i_shock] - (0.7 + S * T)) < 3 * dx
    i_contact = np.where(q[4][0, 0] > 0.5)[0].max()                # alpha_water = 1/2
    assert abs(x[i_contact] - (0.7 + us * T)) < 2 * dx


# ---- three fluids ------------------------------------------------------------------------------
def test_three_fluids_with_two_identical_gases_reduce_to_the_two_fluid_solution():
    """Pins num_fluids = 3 in the oracle (no reference example has three fluids): fluid 3 has the
    equation of state of fluid 1 and takes the same share s of fluid 1 everywhere, so alpha_3 rho_3 =
    s X and alpha_1 rho_1 = (1 - s) X are proportional fields.  The mixture rules
    (m_variables_conversion.fpp:187-227) then see the rho, Gamma, Pi of the two-fluid case, and WENO's
    nonlinear weights are scale-invariant up to weno_eps, so total density, momenta and energy follow
    the two-fluid run and alpha_1 + alpha_3 follows its alpha_1 (to ~1e-9: weno_eps breaks the scale
    invariance where beta ~ 1e-16)."""
    s_share = 0.25
    d2 = cases.shockbubble_2d(Ny=40)
    d3 = dict(d2)
    d3['num_fluids'] = 3
    d3['fluid_pp(3)%gamma'], d3['fluid_pp(3)%pi_inf'] = d2['fluid_pp(1)%gamma'], d2['fluid_pp(1)%pi_inf']
    for i in range(1, d2['num_patches'] + 1):
        a1, r1 = d2[f'patch_icpp({i})%alpha(1)'], d2[f'patch_icpp({i})%alpha_rho(1)']
        d3[f'patch_icpp({i})%alpha(3)'], d3[f'patch_icpp({i})%alpha_rho(3)'] = s_share * a1, s_share * r1
        d3[f'patch_icpp({i})%alpha(1)'], d3[f'patch_icpp({i})%alpha_rho(1)'] = a1 - s_share * a1, r1 - s_share * r1
    out = []
    for d in (d2, d3):
        cfg = dataclasses.replace(cases.config(d), t_step_stop=40)
        cb = pre_process.generate_grid(cfg)
        o = oracle_lib.Oracle(cfg, cb)
        o.set_q(pre_process.generate_initial_condition(cfg, cb))
        oracle_lib.run_p_main(o, cfg)
        out.append(o.get_q())
    q2, q3 = out
    rho2, rho3 = q2[0] + q2[1], q3[0] + q3[1] + q3[2]
    tol = 1e-9
    assert np.abs(rho3 - rho2).max() <= tol * np.abs(rho2).max()
    for k in range(3):                                           # mom_x, mom_y, E
        a, b = q2[2 + k], q3[3 + k]
        assert np.abs(b - a).max() <= tol * max(np.abs(a).max(), 1.0)
    assert np.abs((q3[6] + q3[8]) - q2[5]).max() <= tol           # alpha_1 + alpha_3 == alpha_1 of the two-fluid run
    assert np.abs(q3[7] - q2[6]).max() <= tol
    assert np.abs(q3[2] - s_share / (1 - s_share) * q3[0]).max() <= tol * np.abs(q3[0]).max()   # the share is advected unchanged
