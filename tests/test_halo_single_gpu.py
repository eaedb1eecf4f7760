"""The halo-exchange kernels on ONE GPU (the driver's GPU box has a single device, so the NCCL
tests of tests/test_multigpu.py are skipped there): on a periodic case a rank is its own
neighbour, and the processor-boundary path -- k_halo_pack of the first / last buff_size layers,
a device copy standing in for ncclSend/ncclRecv, k_halo_unpack into the opposite ghost layers
(m_mpi_proxy.fpp:490-601 for x, :733-969 for y; z is the 3-D extension) -- must rebuild the
ghost cells exactly like the periodic boundary kernel (m_rhs.fpp:725-735, :786-787) and like
numpy's wrap padding.  Covers directions 0, 1, 2, buff_size 4 and 6, corners included (the y / z
slabs span the ghosts of the earlier directions, m_mpi_proxy.fpp:736-739)."""
import numpy as np
import pytest

from microfc_b200 import abi, cases

from common import setup_case

pytestmark = pytest.mark.gpu

PERIODIC = {f'bc_{d}%{s}': -1 for d in "xyz" for s in ("beg", "end")}


def _periodic(d, nd):
    out = dict(d)
    for k, v in PERIODIC.items():
        if "xyz".index(k[3]) < nd:
            out[k] = v
    return out


CASES = {
    "1d": lambda: _periodic(cases.sod_1d(Nx=99), 1),
    "2d": lambda: _periodic(cases.shockbubble_2d_cells(70, 45), 2),
    "2d_buff6": lambda: _periodic(cases.viscous_2d(N=39, weno_Re_flux=False), 2),       # viscous: buff_size = 6
    "3d": lambda: _periodic(cases.shockbubble_3d(ncx=40, ncy=33, ncz=29), 3),
}


@pytest.mark.parametrize("name", list(CASES))
def test_pack_copy_unpack_equals_periodic_ghost_fill(name):
    from microfc_b200.simulation import Simulation
    cfg, cb, q0 = setup_case(CASES[name](), n_steps=1)
    nd, b = cfg.num_dims, cfg.buff_size
    rng = np.random.default_rng(3)
    q = q0 * (1.0 + 0.01 * rng.random(q0.shape))                 # every cell distinct
    sim = Simulation(cfg, cb, strict=True)
    try:
        host = np.full((cfg.sys_size,) + sim.ghost_shape, np.nan)
        host[sim._interior()] = q
        out = []
        for mode in (0, 1):
            sim.upload_ghosted(host)                                 # NaN ghosts on the device
            abi.check(sim.L.mfc_b200_debug_fill_ghosts(mode))
            got = np.empty_like(host)
            sim.download_ghosted(got)
            out.append(got)
    finally:
        sim.close()
    pad = [(0, 0)] + [(b, b) if d < nd else (0, 0) for d in (2, 1, 0)]
    want = np.pad(q, pad, mode="wrap")
    assert np.array_equal(out[0], want), "periodic boundary kernel differs from wrap padding"
    assert np.array_equal(out[1], want), "pack -> copy -> unpack differs from the periodic ghost fill"
