#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

    python bench.py --gpus N --steps K --warmup W            (this repo's CUDA path)
    python bench.py --impl reference --gpus N --steps K ...   (CPU restatement of the reference)

One "step" is one SSP-RK3 time step (3 RHS evaluations: ghost fill, cons->prim, x/y/z WENO5 +
HLLC sweeps with the update fused into the last sweep, plus the ICFL diagnostic) on the
workload BASELINE.json's target names: the synthetic 3-D two-fluid shock-bubble at 512^3 cells
PER GPU (configs[4]), weak-scaled over N GPUs with the reference's block decomposition and an
NCCL halo exchange.  Other BASELINE configs: --workload advection_2d_1024 | shockbubble_2d_4096 |
shockdroplet_2d_viscous_2048 | sod_1d_400.

Printed JSON (rank 0, one line):
  value      Mcell-steps/s of the whole job, state resident in HBM, timed with CUDA events on
             the library's launching stream, max over ranks
  e2e        the same metric through the C ABI with HOST buffers: mfc_b200_upload of the state
             from pinned host memory + K x mfc_b200_step (ICFL read back every step) +
             mfc_b200_download, all inside the timed region
  roofline   dominant kernel against the FP64-pipe roof (measured DFMA peak) -- this path is
             FP64-bound once fused; roofline_hbm has the HBM view of the step
  cpu_baseline  the CPU oracle (timing build, all host threads) on a bounded sample
"""
from __future__ import annotations

import argparse
import dataclasses
import json
import os
import subprocess
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm (`--impl reference`, cpu_baseline)
# is meant to use every host core, and the OpenMP runtime reads the variable when it is first
# loaded -- so decide it here, before anything that links libgomp is imported
if os.environ.get("TORCHELASTIC_RUN_ID") or os.environ.get("OMP_NUM_THREADS") == "1":
    os.environ["OMP_NUM_THREADS"] = os.environ.get("MFC_B200_CPU_THREADS", str(os.cpu_count() or 1))

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line.  Libraries write there too (NCCL prints its "NCCL version
# ..." banner to stdout when NCCL_DEBUG is set), so file descriptor 1 is pointed at stderr for
# the whole run and the line goes to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict) -> None:
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

from microfc_b200 import cases, pre_process  # noqa: E402
from microfc_b200.case import CaseConfig  # noqa: E402

# measured on this pool's B200 with tools/fp64_peak.cu (profiles/r01_fp64_peak.json); the
# driver-written MEASURED_PEAKS.json has no FP64 entry
FP64_PEAK_TFLOPS_FALLBACK = 34.07
HBM_FALLBACK_GBS = 6650.0

# algorithmic work per cell, SURVEY.md 8(d) / BASELINE.md 3 (source count of the reference)
WENO_FLOPS = 84            # per variable per direction per cell, m_weno.fpp:476-531
HLLC_FLOPS = {1: 170, 2: 206, 3: 250}   # per face, m_riemann_solvers.fpp:136-325
DIV_FLOPS = 35             # flux divergence + source per direction, m_rhs.fpp:565-653
PRIM_FLOPS = 22
TOPOLOGY = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
# 2-D workloads: the reference's rule gives 4 x 2 on 8 ranks (ties -> larger px, m_mpi_proxy.fpp:177-203)
TOPOLOGY_2D = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (4, 2, 1)}


def sweep_flops_per_cell(E: int, nd: int, with_rk: bool) -> int:
    return WENO_FLOPS * E + HLLC_FLOPS[nd] + DIV_FLOPS + (4 * E if with_rk else 0)


def step_flops_per_cell(E: int, nd: int) -> int:
    per_rhs = sum(sweep_flops_per_cell(E, nd, d == nd - 1) for d in range(nd)) + PRIM_FLOPS
    return 3 * per_rhs


def stretch_grid(d: dict, nd: int) -> dict:
    """Cluster the cells towards the middle of the domain with the reference's tanh-type stretching
    (src/pre_process/m_grid.f90:171-187, a = 2: widths vary by ~2-3x): the sweeps then read per-cell
    WENO coefficient tables (m_weno.fpp:168-363) instead of the uniform-grid constants."""
    out = dict(d)
    for ax in "xyz"[:nd]:
        lo, hi = d[f'{ax}_domain%beg'], d[f'{ax}_domain%end']
        out.update({f'stretch_{ax}': 'T', f'a_{ax}': 2.0, f'{ax}_a': lo + 0.3 * (hi - lo), f'{ax}_b': lo + 0.7 * (hi - lo), f'loops_{ax}': 1})
    out['dt'] = d['dt'] * 0.25
    return out


def workload_case(name: str, n_gpus: int, cells: int | None, strong: bool = False, stretched: bool = False):
    """weak scaling: `cells` per GPU and direction, the global grid grows with the topology;
    strong scaling: ONE global grid of `cells` per direction, split by the reference's rule
    (m_mpi_proxy.fpp:163-203; 4096^2 on 8 ranks -> 4 x 2 blocks of 1024 x 2048)."""
    px, py, pz = TOPOLOGY[n_gpus] if name == "shockbubble_3d_512" else TOPOLOGY_2D[n_gpus]
    topo = (px, py, pz)
    if strong:
        px = py = pz = 1
    if name == "shockbubble_3d_512":
        nc = cells or 512
        d = cases.shockbubble_3d(ncx=nc * px, ncy=nc * py, ncz=nc * pz, Nt=10 ** 6)
        desc = f"synthetic 3D two-fluid shock-bubble, {nc}^3 cells per GPU (BASELINE configs[4]; 3-D is an extension beyond the 1D/2D reference)"
    elif name == "shockbubble_2d_4096":
        nc = cells or 4096
        d = cases.shockbubble_2d_cells(nc * px, nc * py, Nt=10 ** 6)
        desc = f"examples/2D_shockbubble scaled to {nc}^2 cells per GPU (BASELINE configs[2])"
    elif name == "advection_2d_1024":
        nc = cells or 1024
        d = cases.advection_2d(N=nc * px - 1, Nt=10 ** 6)
        d['n'] = nc * py - 1
        desc = f"examples/2D_advection at {nc}^2 cells per GPU (BASELINE configs[1])"
    elif name == "sod_1d_400":
        nc = cells or 400
        d = cases.sod_1d(Nx=nc * n_gpus - 1, Nt=10 ** 6)
        d['dt'] = d['dt'] * 400.0 / nc
        desc = f"examples/1D_sodshocktube, {nc} cells per GPU (BASELINE configs[0], the reference's own CPU-runnable case)"
    elif name == "shockdroplet_2d_viscous_2048":
        nc = cells or 2048
        d = cases.shockdroplet_2d(Nx=nc * px - 1, Ny=nc * py - 1, Nt=10 ** 6, viscous=True)
        # the shipped case takes dt from dx alone (dx = dy there); weak scaling changes the cell
        # aspect of the fixed 0.25 x 0.037 domain, so use the smaller width: ICFL stays 0.1
        d['dt'] = d['dt'] * min(1.0, (0.037 / (nc * py)) / (0.25 / (nc * px)))
        desc = (f"examples/2D_shockdroplet with viscous fluxes (fluid_pp%Re as in examples/2D_viscous), {nc}^2 cells per GPU "
                "(BASELINE configs[3]: 8192x4096 on 8 GPUs)")
    else:
        raise SystemExit(f"unknown workload {name}")
    if stretched:
        if name != "shockbubble_3d_512":
            raise SystemExit("--stretched is wired for the default 3-D workload")
        d = stretch_grid(d, 3)
        desc += ", STRETCHED grid (per-cell WENO coefficient tables)"
    if strong:
        desc = desc.replace("cells per GPU", "cells in total (one grid, strong-scaled)")
    if name == "sod_1d_400":
        topo = (n_gpus, 1, 1)
    return cases.config(d), desc, topo


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i] == "Active"})
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def bind_to_gpu_numa_node(device_index: int):
    """Run this rank (and first-touch its pinned host buffers) on the CPUs NVML reports as local to
    its GPU: with 8 ranks each pushing 9 GB through host memory in the end-to-end leg, buffers on
    the far socket halve the PCIe rate.  Returns the previous affinity (restored before the CPU
    baseline, which uses every host core)."""
    try:
        import pynvml
        prev = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = (max(prev) // 64) + 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, m in enumerate(mask) for b in range(64) if (m >> b) & 1} & prev
        if cpus:
            os.sched_setaffinity(0, cpus)
        return prev
    except Exception:
        return None


def fp64_peak_in_run(device_index: int):
    """tools/fp64_peak (DFMA micro-benchmark, built by __graft_entry__.build()) on this rank's GPU,
    right before the workload: the FP64 roof of THIS box at THIS moment (MEASURED_PEAKS.json has
    no FP64 entry).  None if the binary is missing."""
    exe = os.path.join(ROOT, "tools", "fp64_peak")
    if not os.path.exists(exe):
        return None
    try:
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(device_index))
        r = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=env)
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        return None


def measured_peaks():
    hbm, src = HBM_FALLBACK_GBS, "fallback"
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            hbm, src = float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    fp64, fsrc = FP64_PEAK_TFLOPS_FALLBACK, "fallback (earlier measurement on this pool)"
    fp = os.path.join(ROOT, "profiles", "r01_fp64_peak.json")
    if os.path.exists(fp):
        try:
            fp64 = float(json.load(open(fp))["fp64_dfma_tflops_sustained"])
            fsrc = "measured (tools/fp64_peak.cu on this pool, profiles/r01_fp64_peak.json)"
        except Exception:
            pass
    return hbm, src, fp64, fsrc


def parity_note(workload: str):
    """Where the parity of this workload's kernels is established (tests/, on the GPU) and what the
    fast build's measured deviation from the oracle is: the north-star gate is 1e-10; the water/air
    cases exceed it in the fast build (the strict build is bitwise equal to the oracle on all)."""
    note = {"gate": 1e-10, "strict_build": "bitwise equal to the CPU oracle after 100 RK3 steps (tests/test_gpu_parity.py)",
            "evidence": "profiles/r02_fast_build_linf.txt"}
    rel = {"shockbubble_3d_512": ["shockbubble_3d"], "shockbubble_2d_4096": ["shockbubble_2d"], "advection_2d_1024": ["advection_2d"],
           "sod_1d_400": ["sod_1d"], "shockdroplet_2d_viscous_2048": ["shockdroplet_2d_viscous", "shockdroplet_2d_inviscid"]}[workload]
    fp = os.path.join(ROOT, "profiles", "r02_fast_build_linf.txt")
    if os.path.exists(fp):
        for line in open(fp):
            w = line.lstrip(".F").split()                # pytest's progress dots precede the printed line
            if len(w) >= 8 and w[0] == "ASTLINF":
                w[0] = "FASTLINF"
            if len(w) >= 8 and w[0] == "FASTLINF" and w[1] in rel:
                note.setdefault("fast_build", {})[w[1]] = {"linf": float(w[3]), "linf_per_variable": float(w[5]), "test_bound": float(w[7]),
                                                           "oracle_1ulp_sensitivity": float(w[9]) if len(w) >= 10 else None}
    return note


def cpu_reference_run(cfg_full: CaseConfig, steps: int, warmup: int, sample_cells: int):
    """The CPU restatement of the reference (oracle timing build, OpenMP over all host threads)
    on a bounded sample of the workload: same case, same step, fewer cells."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    nd = cfg_full.num_dims
    if nd == 3:
        d = cases.shockbubble_3d(nc=sample_cells, Nt=10 ** 6)
        sample = f"{sample_cells}^3 cells of the same 3-D shock-bubble case"
    elif nd == 1:
        d = cases.sod_1d(Nx=sample_cells - 1, Nt=10 ** 6)
        sample = f"{sample_cells} cells of the same 1-D case"
    else:
        if cfg_full.viscous:
            d = cases.shockdroplet_2d(Nx=sample_cells - 1, Ny=sample_cells - 1, Nt=10 ** 6, viscous=True)
            d['dt'] = d['dt'] * min(1.0, 0.037 / 0.25)       # same rule as the GPU workload: dt from the smaller width
        elif cfg_full.bc[0][0] == -6:
            d = cases.shockbubble_2d_cells(sample_cells, sample_cells, Nt=10 ** 6)
        else:
            d = cases.advection_2d(N=sample_cells - 1, Nt=10 ** 6)
        sample = f"{sample_cells}^2 cells of the same 2-D case"
    cfg = cases.config(d)
    cb = pre_process.generate_grid(cfg)
    q0 = pre_process.generate_initial_condition(cfg, cb)
    o = oracle_lib.Oracle(cfg, cb, num_procs=1, kind="timing")
    o.set_q(q0)
    o.run_steps(0, warmup, cfg.dt)
    secs = o.run_steps(warmup, steps, cfg.dt)
    ncell = int(np.prod(cfg.shape_glb))
    q = o.get_q()
    assert np.isfinite(q).all()
    return {"cells_run": ncell, "value": ncell * steps / secs / 1e6, "unit": "Mcell-steps/s", "cores": oracle_lib.load("timing").orc_num_threads(),
            "kind": "port", "sample": f"{sample}, {steps} steps after {warmup} warm-up, all host threads (OpenMP); "
            "the Fortran reference cannot be built in this image", "ms_per_step": secs / steps * 1e3,
            "grind_ns_per_cell_eq_rhs": secs / steps / (ncell * cfg.sys_size * 3) * 1e9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="shockbubble_3d_512")
    ap.add_argument("--cells", type=int, default=None, help="cells per GPU and direction (default: the BASELINE size)")
    ap.add_argument("--cpu-sample-cells", type=int, default=None)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --cells per GPU (default); strong: one global grid of --cells per direction over all GPUs")
    ap.add_argument("--stretched", action="store_true", help="stretch the grid (m_grid.f90:171-187): per-cell WENO coefficient tables")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus not in TOPOLOGY:
        raise SystemExit("--gpus must be 1, 2, 4 or 8")
    K, W = args.steps, max(args.warmup, 0)
    strong = args.scaling == "strong"
    cfg, desc, topo = workload_case(args.workload, args.gpus, args.cells, strong, args.stretched)
    nd, E = cfg.num_dims, cfg.sys_size
    ncell_total = int(np.prod(cfg.shape_glb))
    if strong:
        from microfc_b200.domain import processor_topology
        topo = processor_topology(args.gpus, cfg)
    state_mb = ncell_total // args.gpus * E * 8 / 1e6
    config = {"workload": desc, "cells_total": ncell_total, "cells_per_gpu": ncell_total // args.gpus,
              "sys_size": E, "num_dims": nd, "time_stepper": "SSP-RK3", "weno_order": 5, "riemann_solver": "HLLC",
              "run_time_info": bool(cfg.run_time_info), "decomposition": "x".join(map(str, topo)),
              "l2": (f"state of {state_mb:.0f} MB per GPU (x3 stage buffers + RHS) " +
                     ("is far larger than the 126 MB L2; no flush needed" if state_mb * 4 > 4 * 126 else
                      "is comparable to the 126 MB L2: a step touches 4 such buffers in turn, nothing is flushed explicitly"))}

    if args.impl == "reference":
        if rank != 0:
            return
        sample = args.cpu_sample_cells or {3: 128, 2: 1024, 1: 400}[nd]
        r = cpu_reference_run(cfg, K, W, sample)
        # the CPU arm runs a bounded SAMPLE of the workload (same case, fewer cells): its config says
        # what was actually run; throughput per cell is what carries over to the full size
        config = dict(config, cells_total=r["cells_run"], cells_per_gpu=None, decomposition="1 CPU process, OpenMP",
                      sample_of=f"{ncell_total} cells ({desc})", workload=f"{desc} -- CPU arm: {r['sample'].split(',')[0]}")
        line = {"impl": "reference", "metric": "Mcell-steps/s", "value": r["value"], "unit": "Mcell-steps/s",
                "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "grind_ns_per_cell_eq_rhs": r["grind_ns_per_cell_eq_rhs"],
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "Mcell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # a second, gloo group: ranks that wait for rank 0's CPU baseline block in a socket there instead of
    # spinning in an NCCL barrier on cores the OpenMP threads of the baseline want
    cpu_group = None
    if dist is not None:
        try:
            os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")          # one node: the loopback interface always exists
            cpu_group = dist.new_group(backend="gloo")
        except Exception as e:                                         # fall back to the NCCL barrier
            print(f"[bench] no gloo group ({e}); waiting ranks will spin", file=sys.stderr)
            cpu_group = None

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def bcast_id(mine):
        obj = [mine]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]

    peak_run = fp64_peak_in_run(local_rank) if rank == 0 else None
    prev_affinity = bind_to_gpu_numa_node(local_rank)
    from microfc_b200.simulation import Simulation
    cb = pre_process.generate_grid(cfg)
    sim = Simulation(cfg, cb, rank=rank, num_procs=world, device=local_rank, broadcast_id=bcast_id if world > 1 else None)
    lay = sim.layout
    # initial condition of this rank: laid out on the device from the case's patches
    # (mfc_b200_generate_initial_condition) and brought back once into the (pinned)
    # host buffer the Fortran host would own, sf(-b:m+b, -b:n+b, -b:p+b) per variable, so that
    # the end-to-end leg below starts from HOST data
    host = torch.empty((E,) + sim.ghost_shape, dtype=torch.float64, pin_memory=True).numpy()
    sim.generate_initial_condition(cb)
    sim.download_ghosted(host)
    sim.upload_ghosted(host)
    sim.snapshot()
    dt = cfg.dt
    state_bytes = host.nbytes

    # ---- device-resident throughput ------------------------------------------------------------
    t_step = 0
    for _ in range(W):
        sim.step(t_step, dt); t_step += 1
    sampler = ClockSampler(local_rank)
    sampler.start()
    # per-kernel launch durations (the roofline's numerator) come from a CUDA-event pair around every
    # launch on the launching stream, inside the timed region -- except on grids small enough for the
    # library to replay a step as a CUDA graph (< 4 M cells, single rank): there the event records
    # would both cost as much as the kernels and disable the graph, so they get their own pass below
    prof_in_timed = world > 1 or ncell_total >= (4 << 20) or os.environ.get("MFC_B200_GRAPH") == "0"
    if prof_in_timed:
        sim.profile(True)
    launches0 = sim.kernel_launches()
    barrier(); sim.sync()
    sim.timer_start()
    icfl = None
    for _ in range(K):
        icfl = sim.step(t_step, dt)[0]; t_step += 1
    secs = sim.timer_stop()
    barrier()
    secs = max_over_ranks(secs)
    launches = sim.kernel_launches() - launches0
    if not prof_in_timed:
        sim.profile(True)
        for _ in range(min(K, 20)):
            sim.step(t_step, dt); t_step += 1
        sim.sync()
    prof = sim.profile_report()
    sim.profile(False)
    clocks = sampler.stop()
    if cfg.run_time_info:
        if not (icfl == icfl) or icfl > 1.0:
            raise SystemExit(f"unstable run: ICFL = {icfl}")
    elif not np.isfinite(sim.download()).all():      # run_time_info = F (2D_shockdroplet): no ICFL row to watch
        raise SystemExit("unstable run: non-finite state")
    value = ncell_total * K / secs / 1e6

    # ---- end to end through the C ABI with host buffers --------------------------------------
    e2e = None
    if not args.no_e2e:
        sim.restore()
        barrier(); sim.sync()
        t0 = time.perf_counter()
        sim.upload_ghosted(host)                 # H2D from pinned host memory
        ts = 0
        for _ in range(K):
            sim.step(ts, dt); ts += 1            # 24 B of stability data D2H every step
        sim.download_ghosted(host)               # D2H into the host's (pinned) ghosted fields
        sim.sync(); barrier()
        e2e_secs = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": ncell_total * K / e2e_secs / 1e6, "unit": "Mcell-steps/s",
               "h2d_bytes_per_step": state_bytes * world / K, "d2h_bytes_per_step": state_bytes * world / K + 24 * world,
               "timed_region": f"mfc_b200_upload (pinned host -> HBM) + {K} x mfc_b200_step + mfc_b200_download, wall clock, max over ranks"}

    # ---- roofline of the dominant kernel -------------------------------------------------------
    hbm_peak, hbm_src, fp64_peak, fp64_src = measured_peaks()
    if peak_run and peak_run.get("fp64_dfma_tflops_sustained"):
        fp64_peak = float(peak_run["fp64_dfma_tflops_sustained"])
        fp64_src = "measured in this run, right before the workload (tools/fp64_peak: independent DFMA chains, all SMs)"
    if world > 1:                                                 # every rank needs rank 0's figure
        t = torch.tensor([fp64_peak], dtype=torch.float64, device="cuda")
        dist.broadcast(t, src=0)
        fp64_peak = float(t.item())
    sweeps = {k: v for k, v in prof.items() if k in ("k_xstream", "k_march3<y>", "k_march3<z>") and v[1] > 0}
    dom = max(sweeps, key=lambda k: sweeps[k][0])
    dom_t = sweeps[dom][0] / sweeps[dom][1]                      # average launch duration (CUDA events)
    last_dir = {1: "k_xstream", 2: "k_march3<y>", 3: "k_march3<z>"}[nd]
    cells_gpu = ncell_total // args.gpus
    dom_flops = sweep_flops_per_cell(E, nd, dom == last_dir) * cells_gpu
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r02_v7_traffic.json")
    if not os.path.exists(tp):
        tp = os.path.join(ROOT, "profiles", "r01_v4_traffic.json")
    if os.path.exists(tp) and nd == 3 and E == 8:
        # DRAM bytes per cell of this kernel from the committed ncu --set full capture (256^3, same
        # kernels): one stage-1 launch and two stage-2/3 launches per step
        tj = json.load(open(tp))
        k1, k2 = tj["kernels"].get(dom + "/stage1"), tj["kernels"].get(dom + "/stage2")
        if k1 and k2:
            traffic = (k1["dram_bytes_per_cell"] + 2 * k2["dram_bytes_per_cell"]) / 3 * cells_gpu
            traffic_src = tj["source"]
    per_kernel = {}
    for k, (sec, n) in sweeps.items():
        fl = sweep_flops_per_cell(E, nd, k == last_dir) * cells_gpu
        per_kernel[k] = {"avg_launch_ms": sec / n * 1e3, "fp64_tflops_algorithmic": fl / (sec / n) / 1e12,
                         "frac_fp64": fl / (sec / n) / 1e12 / fp64_peak}
    roof = {"bound": "fp64", "kernel": dom, "achieved": dom_flops / dom_t / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": dom_flops / dom_t / 1e12 / fp64_peak, "traffic": traffic, "traffic_source": traffic_src,
            "per_kernel": per_kernel,
            "avg_launch_ms": dom_t * 1e3, "algorithmic_flops_per_cell": sweep_flops_per_cell(E, nd, dom == last_dir),
            "peak_source": fp64_src,
            "note": "FP64-vector-pipe roof (no tensor cores on this path); algorithmic flops = source count of the reference "
                    "(SURVEY.md 8d); see profiles/ for ncu pipe utilisation and DRAM traffic"}
    step_bytes = 8 * E * 8 * cells_gpu
    roof_hbm = {"bound": "hbm", "scope": "whole RK3 step", "achieved": step_bytes * K / secs / 1e9, "peak": hbm_peak,
                "unit": "GB/s", "frac": step_bytes * K / secs / 1e9 / hbm_peak, "peak_source": f"MEASURED_PEAKS.json ({hbm_src})",
                "algorithmic_bytes_per_cell_step": 8 * E * 8}
    kernel_share = {k: {"seconds": v[0], "launches": v[1]} for k, v in prof.items() if v[1] > 0}

    cpu = None
    if prev_affinity:
        os.sched_setaffinity(0, prev_affinity)
    if rank == 0 and not args.no_cpu_baseline:
        sample = args.cpu_sample_cells or {3: 128, 2: 1024, 1: 400}[nd]
        r = cpu_reference_run(cfg, 3, 1, sample)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if cpu_group is not None:
        dist.barrier(group=cpu_group)
    else:
        barrier()

    if rank == 0:
        line = {"metric": "Mcell-steps/s", "value": value, "unit": "Mcell-steps/s", "n_gpus": args.gpus, "steps": K,
                "warmup": W, "ms_per_step": secs / K * 1e3, "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "grind_ns_per_cell_eq_rhs": secs / K / (ncell_total * E * 3) * 1e9,
                "gflops_algorithmic": step_flops_per_cell(E, nd) * ncell_total * K / secs / 1e9,
                "e2e": e2e, "gpu_launches": launches * world, "clocks": clocks,
                "roofline": roof, "roofline_hbm": roof_hbm, "kernel_time": kernel_share,
                "kernel_time_source": "CUDA events around every launch, " + ("inside the timed region" if prof_in_timed else
                                      "separate pass right after the timed region (the timed steps replay a CUDA graph)"),
                "cpu_baseline": cpu,
                "icfl_last": icfl, "fp64_peak_run": peak_run, "parity": parity_note(args.workload)}
        emit(line)
    sim.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
