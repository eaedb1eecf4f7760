"""Worker of tests/test_multigpu.py: launched with torch.distributed.run, one rank per GPU.
Every rank drives libmfc_b200.so on its block of the reference's domain decomposition; the
halo exchange is the library's NCCL send/recv path.  Rank 0 gathers the blocks and compares
them with the CPU oracle on the same decomposition -- which for inviscid cases is bit-identical
to the single-rank oracle (strict mode: bitwise; fast mode: <= 1e-10)."""
import dataclasses
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

from microfc_b200 import cases, pre_process  # noqa: E402
from microfc_b200.simulation import Simulation  # noqa: E402
from common import norm_linf, oracle_run, setup_case  # noqa: E402

def _advection_64x256():
    d = cases.advection_2d(N=63)
    d['n'] = 255
    d['dt'] = d['dt']*(100.0/256.0)
    return d


CASES = {
    "sod_1d": (lambda: cases.sod_1d(Nx=199), 40),
    "shockbubble_2d": (lambda: cases.shockbubble_2d_cells(120, 64), 25),
    "shearlayer_2d": (lambda: cases.shearlayer_2d(Nx=79, Ny=63), 25),          # periodic in x
    "shockdroplet_2d": (lambda: cases.shockdroplet_2d(Nx=199, Ny=59), 20),      # reflective y
    "shockbubble_3d": (lambda: cases.shockbubble_3d(nc=52), 6),
    # >= 78 cells per rank along x: the x sweep is launched in two parts (the cells whose stencils
    # read no x ghost column while the x halo is in flight, then the two boundary strips)
    "shockbubble_2d_wide": (lambda: cases.shockbubble_2d_cells(512, 64), 20),
    "shockbubble_3d_wide": (lambda: cases.shockbubble_3d(ncx=512, ncy=26, ncz=26), 4),
    # viscous: buff_size 6 and corner ghosts (y messages span the x ghosts, m_mpi_proxy.fpp:736-739)
    "viscous_2d": (lambda: cases.viscous_2d(N=63, weno_Re_flux=True), 10),
    "shockdroplet_2d_viscous": (lambda: cases.shockdroplet_2d(Nx=199, Ny=59, viscous=True), 20),
    # smooth field, every viscous stress term active (see tests/test_gpu_parity.py)
    "viscous_wave_2d_weno": (lambda: (cases.viscous_wave_2d(N=64, Nx=64, weno_Re_flux=True), cases.viscous_wave_state), 20),
    "viscous_wave_2d_fd": (lambda: (cases.viscous_wave_2d(N=64, Nx=64, weno_Re_flux=False, bc_y=-6), cases.viscous_wave_state), 20),
    # decompositions the reference's rule splits along y / z (m_mpi_proxy.fpp:163-203): the y and z
    # pack / unpack index maps and neighbour tables
    "advection_2d_ysplit": (lambda: _advection_64x256(), 20),                                 # 2 ranks: 1 x 2
    "viscous_wave_2d_ysplit": (lambda: (cases.viscous_wave_2d(N=128, Nx=64, weno_Re_flux=True), cases.viscous_wave_state), 10),   # 1 x 2, periodic y, corners
    # in-sweep viscous path with y neighbours: the y halo flies under the x sweep of the inner rows
    "viscous_wave_2d_fd_ysplit": (lambda: (cases.viscous_wave_2d(N=128, Nx=64, weno_Re_flux=False), cases.viscous_wave_state), 10),
    "shockbubble_3d_ysplit": (lambda: cases.shockbubble_3d(ncx=26, ncy=52, ncz=26), 6),      # 2 ranks: 1 x 2 x 1
    "shockbubble_3d_zsplit": (lambda: cases.shockbubble_3d(ncx=26, ncy=26, ncz=52), 6),      # 2 ranks: 1 x 1 x 2
    "shockbubble_3d_zsplit_periodic": (lambda: cases.shockbubble_3d(ncx=26, ncy=26, ncz=52, periodic_z=True), 6),
    "shockbubble_3d_yzsplit": (lambda: cases.shockbubble_3d(ncx=26, ncy=52, ncz=52), 6),     # 4 ranks: 1 x 2 x 2
    "shockbubble_3d_xysplit": (lambda: cases.shockbubble_3d(ncx=64, ncy=52, ncz=26), 6),     # 4 ranks: 2 x 2 x 1
    "shockbubble_2d_4x2": (lambda: cases.shockbubble_2d_cells(256, 128), 20),                 # 8 ranks: 4 x 2 (configs[2]'s layout)
    "viscous_wave_2d_4x2": (lambda: (cases.viscous_wave_2d(N=128, Nx=128, weno_Re_flux=True), cases.viscous_wave_state), 10),   # 8 ranks: 4 x 2 (configs[3]'s layout)
}
# fixed bounds of the badly conditioned cases, see FAST_TOL in tests/test_gpu_parity.py
FAST_TOL = {"shockdroplet_2d": 1.5e-9, "shockdroplet_2d_viscous": 2.0e-9}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def bcast(mine):
        obj = [mine]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]

    ok = True
    for name in sys.argv[1:]:
        mk, n = CASES[name]
        cfg, cb, q0 = setup_case(mk(), n_steps=n)
        ref = None
        if rank == 0:
            # the oracle on the SAME decomposition: the inviscid scheme is decomposition-invariant
            # bit for bit, the viscous finite-difference gradients are not (every rank rebuilds the
            # ghost cell centres x_cc by accumulation, m_start_up.fpp:517-655, so the reference
            # itself moves by ~1e-27 with the number of ranks)
            ref, rows_ref = oracle_run(cfg, cb, q0, num_procs=world)
            if not cfg.viscous:
                ref1, _ = oracle_run(cfg, cb, q0)
                assert np.array_equal(ref, ref1), "oracle: inviscid result depends on the decomposition"

        # the split x launch (interior cells while the x halo is in flight, then the boundary strips)
        # is off by default; the wide cases switch it on
        os.environ["MFC_B200_XSPLIT"] = "1" if "wide" in name else "0"
        # the x halo travels in pieces only above 16 MB per face: force three pieces here so that the
        # piecewise exchange + piecewise x sweep are what these small cases run
        os.environ["MFC_B200_XPIECES"] = "3"
        for strict in (True, False):
            sim = Simulation(cfg, cb, rank=rank, num_procs=world, strict=strict, device=local, broadcast_id=bcast)
            sim.upload(sim.scatter(q0))
            rows = sim.run()
            mine = sim.download()
            lay = sim.layout
            sim.close()
            parts = [None] * world if rank == 0 else None
            dist.gather_object((lay.interior_slices(), mine), parts, dst=0)
            if rank == 0:
                out = np.full_like(q0, np.nan)
                for sl, blk in parts:
                    out[(slice(None),) + sl] = blk
                if strict:
                    good = np.array_equal(out, ref)
                    if cfg.run_time_info:
                        good = good and all(a[2][0] == b[2][0] for a, b in zip(rows, rows_ref))
                else:
                    err = norm_linf(out, ref, cfg)
                    good = bool((err <= FAST_TOL.get(name, 1e-10)).all())
                layout = "x".join(str(n) for n in lay.np_dir)
                print(f"{name} world={world} layout={layout} strict={strict}: {'OK' if good else 'MISMATCH ' + str(norm_linf(out, ref, cfg))}", flush=True)
                ok = ok and good
    dist.barrier()
    if rank == 0:
        print("NCCL_WORKER_OK" if ok else "NCCL_WORKER_FAIL", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
