"""Regenerates tests/golden/*.npz.

    python tests/golden/make_golden.py

The reference cannot be built or imported in this image (Fortran 2008 + fypp + MPI; SURVEY.md
8c), so these vectors come from the strict build of the CPU oracle (oracle/liborc_strict.so,
-O2 -ffp-contract=off), whose own pinning is tests/test_oracle_known_answers.py.  They freeze
the oracle's output bit-for-bit so that (a) a change to the oracle cannot silently move the
parity target and (b) the CUDA strict path can be checked on the GPU box against committed
bytes.  PARITY UNPINNED with respect to the Fortran reference itself.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

from microfc_b200 import cases  # noqa: E402
from common import oracle_run, setup_case  # noqa: E402

GOLDEN = {
    # name: (case dictionary builder, steps)
    "sod_1d_100": (lambda: cases.sod_1d(), 100),
    "kapila_1d_60": (lambda: cases.kapila_1d(Nx=199), 60),
    "advection_2d_48_30": (lambda: cases.advection_2d(N=47), 30),
    "shockbubble_2d_25": (lambda: cases.shockbubble_2d(Ny=26), 25),
    "shockdroplet_2d_20": (lambda: cases.shockdroplet_2d(Nx=149, Ny=44), 20),
    "shearlayer_2d_20": (lambda: cases.shearlayer_2d(Nx=39, Ny=29), 20),
    "viscous_2d_fd_10": (lambda: cases.viscous_2d(N=29, Nt=10, weno_Re_flux=False), 10),
    "viscous_2d_weno_10": (lambda: cases.viscous_2d(N=29, Nt=10, weno_Re_flux=True), 10),
    "shockbubble_3d_26_6": (lambda: cases.shockbubble_3d(nc=26), 6),
    # smooth field: all viscous stress terms active (the viscous_2d vectors above never move,
    # their velocity is piecewise constant)
    "viscous_wave_2d_weno_20": (lambda: (cases.viscous_wave_2d(N=32, Nx=26, weno_Re_flux=True), cases.viscous_wave_state), 20),
    "viscous_wave_2d_fd_20": (lambda: (cases.viscous_wave_2d(N=32, Nx=26, weno_Re_flux=False, bc_y=-6), cases.viscous_wave_state), 20),
}


def main():
    only = sys.argv[1:]
    for name, (mk, n) in GOLDEN.items():
        if only and name not in only:
            continue
        cfg, cb, q0 = setup_case(mk(), n_steps=n)
        q, rows = oracle_run(cfg, cb, q0)
        stab = np.array([[r[1]] + [x if x == x else -1.0 for x in r[2]] for r in rows])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), q=q, stab=stab)
        print(name, q.shape, os.path.getsize(os.path.join(HERE, name + ".npz")))


if __name__ == "__main__":
    main()
