"""Command-line driver: what ``./mfc.sh run <case.py> -t pre_process simulation`` does in the
reference (toolchain/mfc/run/input.py:100-112, src/pre_process/p_main.f90,
src/simulation/p_main.fpp), with the CUDA library as the ``simulation`` target.

    python -m microfc_b200 pre_process <case.py> [--case-dir DIR]
    python -m microfc_b200 simulation  <case.py> [--case-dir DIR] [--strict]
    python -m microfc_b200 run         <case.py> [--case-dir DIR] [--strict]      (both)

``pre_process`` writes the grid and initial condition in the reference's formats
(microfc_b200/data_io.py); ``simulation`` starts from those files -- they may equally have been
written by the Fortran pre_process -- and saves every ``t_step_save`` steps in the same
formats, plus ``run_time.inf`` and ``time_data.dat``.  Under ``torchrun`` every rank drives one
GPU on its block of the reference's domain decomposition.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

from . import data_io, pre_process
from .case import load_case_file
from .domain import rank_layout


def _local_cb(cb_glb, lay, nd):
    # x_cb(-1:m) of this rank = x_cb_glb(start-1 : start+m), m_start_up.fpp:432
    sl = lay.interior_slices()[::-1]
    return [cb_glb[d][sl[d].start:sl[d].stop + 1] for d in range(nd)]


def do_pre_process(cfg, case_dir, rank=0, world=1):
    cb = pre_process.generate_grid(cfg)
    lay = rank_layout(rank, world, cfg)
    q = pre_process.generate_initial_condition(cfg, cb, box=lay.interior_slices())
    if cfg.parallel_io:
        if rank == 0:
            data_io.write_grid_parallel(case_dir, cb)
        data_io.write_restart_parallel(case_dir, 0, q, cfg, lay.interior_slices())
    else:
        data_io.write_serial(case_dir, rank, 0, _local_cb(cb, lay, cfg.num_dims), q)


def do_simulation(cfg, case_dir, strict=False, rank=0, world=1, local_rank=0, broadcast_id=None):
    from .simulation import Simulation
    lay = rank_layout(rank, world, cfg)
    t0 = cfg.t_step_start
    if cfg.parallel_io:
        cb = data_io.read_grid_parallel(case_dir, cfg)
        q = data_io.read_restart_parallel(case_dir, t0, cfg, lay.interior_slices())
    else:
        if world > 1:
            raise SystemExit("serial I/O with several ranks needs the global grid; use parallel_io = T")
        cb, q = data_io.read_serial(case_dir, rank, t0, cfg, lay.shape)
    sim = Simulation(cfg, cb, rank=rank, num_procs=world, strict=strict, device=local_rank, broadcast_id=broadcast_id)
    sim.upload(q)
    rti = data_io.RunTimeInfo(case_dir, cfg.viscous) if (cfg.run_time_info and rank == 0) else None
    secs = []

    def save(t_step):
        qs = sim.download()
        if cfg.parallel_io:
            data_io.write_restart_parallel(case_dir, t_step, qs, cfg, lay.interior_slices())
        else:
            cbl = _local_cb(cb, lay, cfg.num_dims)
            data_io.write_serial(case_dir, rank, t_step, cbl, qs)
            data_io.write_ascii(case_dir, rank, t_step, cbl, qs, cfg)

    def after_step(s, t_step):                     # p_main.fpp:290-317
        secs.append(s.last_step_seconds)
        if cfg.t_step_save > 0 and ((t_step - cfg.t_step_start) % cfg.t_step_save == 0 or t_step == cfg.t_step_stop):
            save(t_step)

    rows = sim.run(callback=after_step)
    if rti is not None:
        for t_step, dt, stab in rows:
            rti.row(t_step, dt, stab)
        rti.close()
    if rank == 0 and len(secs) > 4:                # time_avg, m_time_steppers.fpp:352-358
        data_io.append_time_data(case_dir, world, float(np.mean(secs[4:])))
    sim.close()
    return rows


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m microfc_b200")
    ap.add_argument("target", choices=["pre_process", "simulation", "run"])
    ap.add_argument("case")
    ap.add_argument("--case-dir", default=None, help="default: the directory of the case file")
    ap.add_argument("--strict", action="store_true", help="no FMA contraction, reference operation order")
    args = ap.parse_args(argv)
    cfg = load_case_file(args.case)
    cfg.check()
    case_dir = args.case_dir or os.path.dirname(os.path.abspath(args.case))
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    bcast = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl" if args.target != "pre_process" else "gloo")

        def bcast(mine):
            obj = [mine]
            dist.broadcast_object_list(obj, src=0)
            return obj[0]
    if args.target in ("pre_process", "run"):
        do_pre_process(cfg, case_dir, rank, world)
        if world > 1:
            dist.barrier()
    if args.target in ("simulation", "run"):
        rows = do_simulation(cfg, case_dir, args.strict, rank, world, local, bcast)
        if rank == 0 and rows:
            print(f"{len(rows)} time steps; last ICFL = {rows[-1][2][0]}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1:])
