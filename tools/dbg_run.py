#!/usr/bin/env python
"""Run one named parity case on the GPU and compare with the oracle (debug helper):
    python tools/dbg_run.py <case> <strict 0|1> [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402

from common import gpu_run, norm_linf, oracle_run, setup_case  # noqa: E402
from test_gpu_parity import CASES  # noqa: E402

name, strict = sys.argv[1], bool(int(sys.argv[2]))
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
cfg, cb, q0 = setup_case(CASES[name](), n_steps=steps)
q_gpu, _ = gpu_run(cfg, cb, q0, strict=strict)
q_ref, _ = oracle_run(cfg, cb, q0)
print(name, "strict" if strict else "fast", "equal" if np.array_equal(q_gpu, q_ref) else "differs", norm_linf(q_gpu, q_ref, cfg))
