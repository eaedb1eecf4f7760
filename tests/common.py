"""Shared helpers for the test-suite."""
from __future__ import annotations

import dataclasses

import numpy as np

from microfc_b200 import cases, pre_process
from microfc_b200.case import CaseConfig

import oracle_lib


def setup_case(case, n_steps=None):
    """case dictionary, or (case dictionary, state function) for cases whose initial state is not
    laid down by patches -> (cfg, global cell boundaries, initial conservative state)."""
    case_dict, state = case if isinstance(case, tuple) else (case, None)
    cfg = cases.config(case_dict)
    if n_steps is not None:
        cfg = dataclasses.replace(cfg, t_step_stop=cfg.t_step_start + n_steps)
    cb = pre_process.generate_grid(cfg)
    q0 = state(cfg, cb) if state is not None else pre_process.generate_initial_condition(cfg, cb)
    return cfg, cb, q0


def norm_linf(a: np.ndarray, b: np.ndarray, cfg: CaseConfig | None = None) -> np.ndarray:
    """The parity metric of BASELINE.json / SURVEY.md 8d: per conservative variable,
    max|a-b| / max|b| (pointwise relative error is meaningless where e.g. a momentum is ~0).
    With cfg, the momentum components are scaled as one vector (by the largest component):
    in a flow along x the y-momentum is pure round-off noise and has no scale of its own."""
    E = a.shape[0]
    num = np.abs(a - b).reshape(E, -1).max(axis=1)
    den = np.abs(b).reshape(E, -1).max(axis=1)
    if cfg is not None:
        nf, nd = cfg.num_fluids, cfg.num_dims
        den[nf:nf + nd] = den[nf:nf + nd].max()
    den = np.where(den == 0.0, 1.0, den)
    return num / den


def roundoff_sensitivity(cfg: CaseConfig, cb, q0, q_ref) -> np.ndarray:
    """How far the ORACLE itself moves when every cell's energy is nudged by +-1 ulp: the
    conditioning of the case.  For stiffened-gas liquids (pi_inf ~ 1e9) p = (E - dyn - Pi)/Gamma
    cancels ~4 digits, so any two non-bit-identical evaluations of the same algorithm (e.g.
    gcc with and without FMA contraction) differ by more than 1e-10 after 100 steps."""
    rng = np.random.default_rng(0)
    Ei = cfg.num_fluids + cfg.num_dims
    q1 = q0.copy()
    q1[Ei] = np.where(rng.random(q0[Ei].shape) < 0.5, np.nextafter(q0[Ei], np.inf), np.nextafter(q0[Ei], -np.inf))
    q_p, _ = oracle_run(cfg, cb, q1)
    return norm_linf(q_p, q_ref, cfg)


def oracle_run(cfg: CaseConfig, cb, q0, num_procs=1, kind="strict"):
    o = oracle_lib.Oracle(cfg, cb, num_procs=num_procs, kind=kind)
    o.set_q(q0)
    rows = oracle_lib.run_p_main(o, cfg)
    return o.get_q(), rows


def gpu_run(cfg: CaseConfig, cb, q0, strict=False):
    from microfc_b200.simulation import Simulation
    sim = Simulation(cfg, cb, strict=strict)
    try:
        sim.upload(sim.scatter(q0))
        rows = sim.run()
        q = sim.download()
    finally:
        sim.close()
    return q, rows
