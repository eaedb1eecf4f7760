"""Host-side logic of the path (no GPU): case front-end, pre_process restatement, domain
decomposition and ghosted metrics -- each checked against the reference's rules
(file:line in the docstrings of the code under test) and, where the CPU oracle has its own
independent restatement (decomposition, metrics), against the oracle."""
import ctypes as C
import dataclasses
import os

import numpy as np
import pytest

from microfc_b200 import cases, domain, pre_process
from microfc_b200.case import CaseConfig, load_case_file, parse_case

import oracle_lib
from common import setup_case

REF_EXAMPLES = "/root/reference/examples"


# ---- case front-end -----------------------------------------------------------------------------
def test_parse_case_defaults_and_indices():
    d = cases.advection_2d(N=99)
    cfg = cases.config(d)
    assert (cfg.m, cfg.n, cfg.p) == (99, 99, 0)
    assert cfg.num_dims == 2 and cfg.num_fluids == 2 and cfg.sys_size == 7      # m_global_parameters.fpp:302-310
    assert cfg.weno_polyn == 2 and cfg.buff_size == 4 and not cfg.viscous         # :356-360
    v = cases.config(cases.viscous_2d(N=50))
    assert v.viscous and v.buff_size == 6
    assert cases.config(cases.sod_1d()).sys_size == 4
    assert cases.config(cases.shockbubble_3d(nc=32)).sys_size == 8


@pytest.mark.parametrize("field,value,msg", [
    ("m", 0, "m"), ("weno_order", 4, "weno_order"), ("time_stepper", 5, "time_stepper"),
    ("weno_eps", 1e-3, "weno_eps"), ("dt", -1.0, "dt"), ("num_fluids", 9, "num_fluids"),
])
def test_input_checks_mirror_s_check_input_file(field, value, msg):
    cfg = cases.config(cases.sod_1d())
    bad = dataclasses.replace(cfg, **{field: value})
    with pytest.raises(ValueError, match=f"Unsupported value of {msg}"):
        bad.check()


def test_too_few_cells_for_weno_order_is_rejected():
    d = cases.sod_1d(Nx=23)                                     # 24 cells < 5*weno_order
    with pytest.raises(ValueError, match="m and weno_order"):
        cases.config(d)


@pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("ex,builder", [("1D_sodshocktube", lambda: cases.sod_1d()),
                                        ("2D_advection", lambda: cases.advection_2d(N=99)),
                                        ("2D_shockbubble", lambda: cases.shockbubble_2d())])
def test_case_builders_equal_the_reference_case_files(ex, builder):
    """The unchanged reference case.py scripts parse to the same configuration as our builders
    at the shipped resolution (input files stay unchanged, BASELINE.json north_star)."""
    ref = load_case_file(os.path.join(REF_EXAMPLES, ex, "case.py"))
    ours = cases.config(builder())
    for f in ("m", "n", "p", "num_fluids", "weno_order", "time_stepper", "bc", "num_patches", "run_time_info"):
        assert getattr(ref, f) == getattr(ours, f), f
    assert np.allclose(ref.dt, ours.dt, rtol=1e-14)
    assert np.allclose(ref.gamma[:ref.num_fluids], ours.gamma[:ours.num_fluids], rtol=1e-14)
    assert np.allclose(ref.pi_inf[:ref.num_fluids], ours.pi_inf[:ours.num_fluids], rtol=1e-14)
    for pr, po in zip(ref.patches, ours.patches):
        assert pr.geometry == po.geometry
        assert np.allclose(pr.alpha_rho, po.alpha_rho, rtol=1e-14) and np.allclose(pr.vel, po.vel, rtol=1e-14)
        assert np.isclose(pr.pres, po.pres, rtol=1e-14)


# ---- pre_process ---------------------------------------------------------------------------------
def test_uniform_grid_matches_m_grid():
    cfg = cases.config(cases.sod_1d(Nx=99))
    cb = pre_process.generate_grid(cfg)[0]
    assert cb.shape == (101,) and cb[0] == 0.0 and cb[-1] == 1.0          # x_cb(-1) = beg, x_cb(m) = end
    dx = (1.0 - 0.0) / 100.0
    assert np.array_equal(cb[:-1], 0.0 + dx * np.arange(100.0))            # m_grid.f90:167-169


def test_sod_initial_condition_and_prim_to_cons():
    cfg, cb, q0 = setup_case(cases.sod_1d())
    x = 0.5 * (cb[0][1:] + cb[0][:-1])
    rho, mom, E, al = q0[0, 0, 0], q0[1, 0, 0], q0[2, 0, 0], q0[3, 0, 0]
    left = x < 0.5
    assert np.array_equal(rho[left], np.full(left.sum(), 1.0)) and np.array_equal(rho[~left], np.full((~left).sum(), 0.125))
    assert np.array_equal(mom, np.zeros_like(mom)) and np.array_equal(al, np.ones_like(al))
    g = cfg.gamma[0]
    assert np.allclose(E[left], g * 1.0) and np.allclose(E[~left], g * 0.1)    # E = Gamma p + Pi + 0.5 rho u^2


def test_smoothed_circle_patch_is_bounded_and_mixture_consistent():
    cfg, cb, q0 = setup_case(cases.advection_2d(N=63))
    nf, nd = cfg.num_fluids, cfg.num_dims
    al = q0[nf + nd + 1:]
    assert np.allclose(al.sum(axis=0), 1.0, atol=1e-12) and (al >= 0).all()
    # patch 2 (smoothed circle at (0.1, 0.1)) replaces alpha_rho(1) = 1000 by 1 across a tanh layer
    ar1 = q0[0, 0]
    assert 1.0 <= ar1.min() < 10.0 and ar1.max() == 1000.0 and ((ar1 > 200) & (ar1 < 800)).any()
    assert 0.0 <= al[0].min() and al[0].max() <= 1e-12
    rho = q0[:nf].sum(axis=0)
    assert np.allclose(q0[nf] / rho, 100.0, rtol=1e-13) and np.allclose(q0[nf + 1] / rho, 100.0, rtol=1e-13)
    Gam = sum(al[i] * cfg.gamma[i] for i in range(nf))
    Pi = sum(al[i] * cfg.pi_inf[i] for i in range(nf))
    p = (q0[nf + nd] - 0.5 * (q0[nf] ** 2 + q0[nf + 1] ** 2) / rho - Pi) / Gam
    assert np.allclose(p, 1e5, rtol=1e-10)


def test_initial_condition_box_equals_slice_of_global():
    cfg, cb, q0 = setup_case(cases.shockbubble_3d(nc=32))
    box = (slice(5, 17), slice(0, 32), slice(8, 30))
    part = pre_process.generate_initial_condition(cfg, cb, box=box)
    assert np.array_equal(part, q0[(slice(None),) + box])


# ---- decomposition (m_mpi_proxy.fpp:134-328) -----------------------------------------------------
@pytest.mark.parametrize("shape,nprocs,expect", [
    ((4096, 4096, 1), 8, (4, 2, 1)),          # tie |Mx/px - Ny/py| -> larger px  
    ((8192, 4096, 1), 8, (4, 2, 1)),          # 2048 x 2048 per rank
    ((400, 1, 1), 4, (4, 1, 1)),
    ((300, 100, 1), 4, (4, 1, 1)),
    ((512, 512, 512), 8, (2, 2, 2)),
    ((1024, 1024, 1), 2, (2, 1, 1)),
])
def test_processor_topology(shape, nprocs, expect):
    cfg = dataclasses.replace(cases.config(cases.advection_2d(N=99)), m=shape[0] - 1, n=shape[1] - 1, p=shape[2] - 1)
    assert domain.processor_topology(nprocs, cfg) == expect
    nd = cfg.num_dims
    L = oracle_lib.load("strict")
    out = (C.c_int * 3)()
    Nglb = (C.c_int * 3)(cfg.m, cfg.n, cfg.p)
    assert L.orc_decompose(nprocs, nd, Nglb, 5, out) == 0
    assert tuple(out) == expect


def test_topology_rejects_too_many_ranks():
    cfg = cases.config(cases.advection_2d(N=49))
    with pytest.raises(ValueError, match="num_procs"):
        domain.processor_topology(16, cfg)


@pytest.mark.parametrize("name,nprocs", [("sod", 3), ("shockbubble_2d", 4), ("shear_periodic", 4), ("shockbubble_3d", 8)])
def test_rank_layout_and_metrics_equal_the_oracles(name, nprocs):
    d = {"sod": lambda: cases.sod_1d(Nx=100), "shockbubble_2d": lambda: cases.shockbubble_2d_cells(103, 57),
         "shear_periodic": lambda: cases.shearlayer_2d(Nx=63, Ny=55), "shockbubble_3d": lambda: cases.shockbubble_3d(nc=53)}[name]()
    cfg = cases.config(d)
    cb = pre_process.generate_grid(cfg)
    o = oracle_lib.Oracle(cfg, cb, num_procs=nprocs)
    lays = [domain.rank_layout(r, nprocs, cfg) for r in range(nprocs)]
    cells = 0
    for r, lay in enumerate(lays):
        info = o.rank_info(r)
        assert info["N"][:cfg.num_dims] == lay.N[:cfg.num_dims]
        assert info["start_idx"][:cfg.num_dims] == lay.start_idx[:cfg.num_dims]
        assert info["bc"][:cfg.num_dims] == lay.bc[:cfg.num_dims]
        assert tuple(info["coords"]) == tuple(lay.coords)
        cells += int(np.prod(lay.shape))
        met = domain.ghosted_metrics(lay, cfg, cb, lays)
        for dd in range(cfg.num_dims):
            ocb, occ, ods = o.rank_metrics(r, dd)
            assert np.array_equal(met.cb[dd], ocb) and np.array_equal(met.cc[dd], occ) and np.array_equal(met.ds[dd], ods)
    assert cells == int(np.prod(cfg.shape_glb))                            # the blocks tile the domain


def test_remainder_cells_go_to_lowest_coordinates():
    cfg = cases.config(cases.sod_1d(Nx=100))                               # 101 cells on 3 ranks
    lays = [domain.rank_layout(r, 3, cfg) for r in range(3)]
    assert [l.N[0] + 1 for l in lays] == [34, 34, 33]                      # m_mpi_proxy.fpp:229-239
    assert [l.start_idx[0] for l in lays] == [0, 34, 68]
    assert lays[0].bc[0] == [-3, 1] and lays[1].bc[0] == [0, 2] and lays[2].bc[0] == [1, -3]


def test_periodic_neighbours_wrap():
    cfg = cases.config(cases.shearlayer_2d(Nx=63, Ny=55))                  # periodic in x
    lays = [domain.rank_layout(r, 2, cfg) for r in range(2)]
    assert lays[0].np_dir == (2, 1, 1)
    assert lays[0].bc[0] == [1, 1] and lays[1].bc[0] == [0, 0]             # m_mpi_proxy.fpp:242-255
    assert lays[0].bc[1] == list(cfg.bc[1])


# ---- the p_main loop (p_main.fpp:196-318) --------------------------------------------------------
def test_p_main_loop_dt_tweak_and_last_iteration():
    class Probe:
        def __init__(self): self.calls = []
        def step(self, t, dt): self.calls.append((t, dt)); return [0.0, 0.0, 0.0]
    cfg = dataclasses.replace(cases.config(cases.sod_1d()), t_step_stop=10)
    pr = Probe()
    rows = oracle_lib.run_p_main(pr, cfg)
    assert [c[0] for c in pr.calls] == list(range(0, 11))                  # t_step_stop is visited (no update there)
    assert all(c[1] == cfg.dt for c in pr.calls[:9])
    # the host overwrites dt with finaltime - mytime once mytime + dt >= finaltime (:287)
    t = sum(c[1] for c in pr.calls[:10])
    assert abs(t - 10 * cfg.dt) < 1e-18 and abs(pr.calls[10][1]) < 1e-15
    assert len(rows) == 11


# ---- device-side initial condition: what the host hands to mfc_b200_generate_initial_condition ----
def test_patch_array_mirrors_the_case_patches():
    from microfc_b200 import abi
    from microfc_b200.simulation import patch_array
    cfg = cases.config(cases.shockbubble_3d(nc=32))
    arr = patch_array(cfg)
    assert len(arr) == cfg.num_patches == 3
    for c, pt in zip(arr, cfg.patches):
        assert c.geometry == pt.geometry and bool(c.smoothen) == pt.smoothen and c.smooth_patch_id == pt.smooth_patch_id
        assert [bool(c.alter_patch[k]) for k in range(cfg.num_patches + 1)] == \
               [bool(pt.alter_patch.get(k, False)) for k in range(cfg.num_patches + 1)]
        assert all(c.alter_patch[k] == 0 for k in range(cfg.num_patches + 1, abi.MAX_PATCHES + 1))
        assert (c.x_centroid, c.y_centroid, c.z_centroid, c.radius, c.pres) == \
               (pt.x_centroid, pt.y_centroid, pt.z_centroid, pt.radius, pt.pres)
        assert list(c.vel) == pt.vel and list(c.alpha_rho) == pt.alpha_rho and list(c.alpha) == pt.alpha
    assert bool(arr[0].alter_patch[0]) and bool(arr[1].alter_patch[1]) and arr[2].smooth_patch_id == 1


def test_rank_cell_centres_tile_the_global_grid():
    from microfc_b200 import pre_process
    from microfc_b200.domain import rank_layout
    from microfc_b200.simulation import rank_cell_centres
    cfg = cases.config(cases.shockbubble_3d(ncx=64, ncy=52, ncz=50))
    cb = pre_process.generate_grid(cfg)
    glb = [(c[1:] + c[:-1]) / 2.0 for c in cb]
    got = [np.full_like(g, np.nan) for g in glb]
    for r in range(8):
        lay = rank_layout(r, 8, cfg)
        cc, ds_min = rank_cell_centres(cfg, lay, cb)
        zs, ys, xs = lay.interior_slices()
        for d, sl in enumerate((xs, ys, zs)):
            assert len(cc[d]) == lay.N[d] + 1
            got[d][sl] = cc[d]
        assert ds_min == min(float(np.min(c[1:] - c[:-1])) for c in cb)      # GLOBAL minimum on every rank
    for d in range(3):
        assert np.array_equal(got[d], glb[d])


def test_analytical_and_vortex_patches_of_the_host_pre_process():
    """Geometries 6, 7, 15 (m_create_patches.fpp:379-534) in microfc_b200/pre_process.py: 6 is a hard
    circle; 7 / 15 are a rectangle / line segment whose pressure carries the factor
    1 + 0.2 exp(-((x_cb(i) - xc)^2 [+ (y_cb(j) - yc)^2]) / (2 * 0.005)), evaluated at the RIGHT cell
    boundaries (not the centres)."""
    import dataclasses
    import numpy as np
    from microfc_b200 import cases, pre_process
    # 1-D, geometry 15 against geometry 1 of the same extent
    base = cases.config(cases.sod_1d(Nx=99))
    ps = [dataclasses.replace(p) for p in base.patches]
    ps[0].geometry = 15
    cfg = dataclasses.replace(base, patches=ps)
    cb = pre_process.generate_grid(cfg)
    q15 = pre_process.generate_initial_condition(cfg, cb)
    q1 = pre_process.generate_initial_condition(base, cb)
    xc = ps[0].x_centroid
    x_cc = (cb[0][1:] + cb[0][:-1]) / 2
    inside = (x_cc >= xc - 0.5 * ps[0].length_x) & (x_cc <= xc + 0.5 * ps[0].length_x)
    factor = np.where(inside, 1.0 + 0.2 * np.exp(-((cb[0][1:] - xc) ** 2) / 0.01), 1.0)
    gam = cfg.gamma[0]
    p1 = (q1[2, 0, 0] - 0.5 * q1[1, 0, 0] ** 2 / q1[0, 0, 0]) / gam
    p15 = (q15[2, 0, 0] - 0.5 * q15[1, 0, 0] ** 2 / q15[0, 0, 0]) / gam
    assert np.allclose(p15, p1 * factor, rtol=1e-13)
    assert np.array_equal(q15[0], q1[0]) and np.array_equal(q15[3], q1[3])
    # 2-D: geometry 6 == geometry 2 without smoothing; geometry 7 == geometry 3 times the bump
    b2 = cases.config(cases.advection_2d(N=39))
    for geo_a, geo_b in ((6, 2), (7, 3)):
        pa = [dataclasses.replace(p) for p in b2.patches]
        pb = [dataclasses.replace(p) for p in b2.patches]
        for p_, geo in ((pa[-1], geo_a), (pb[-1], geo_b)):
            p_.geometry, p_.smoothen = geo, False
            p_.x_centroid, p_.y_centroid, p_.radius, p_.length_x, p_.length_y = 0.5, 0.5, 0.2, 0.4, 0.3
        ca, cb_ = dataclasses.replace(b2, patches=pa), dataclasses.replace(b2, patches=pb)
        grid = pre_process.generate_grid(ca)
        qa, qb = pre_process.generate_initial_condition(ca, grid), pre_process.generate_initial_condition(cb_, grid)
        if geo_a == 6:
            # (geometry 2 would also repaint the cells of its smooth_patch_id, m_create_patches.fpp:131-137;
            # the vortex patch has no such clause) inside the circle: this patch; outside: the background
            x = (grid[0][1:] + grid[0][:-1]) / 2
            ins = ((x[None, :] - 0.5) ** 2 + (x[:, None] - 0.5) ** 2) <= 0.2 ** 2
            assert ins.any() and not ins.all()
            assert np.all(qa[0, 0][ins] == pa[-1].alpha_rho[0]) and np.all(qa[0, 0][~ins] == pa[0].alpha_rho[0])
        else:
            assert np.array_equal(qa[:4], qb[:4]) and np.array_equal(qa[5:], qb[5:])
            assert (qa[4] >= qb[4]).all() and (qa[4] > qb[4]).any()
