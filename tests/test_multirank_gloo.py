"""world_size > 1 on CPU (gloo): the host-side multi-rank logic, and bench.py's reference arm
under torchrun (rank 0 alone works and prints, the others exit 0)."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torchrun(nproc, script, *args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), script, *args]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)


@pytest.mark.parametrize("nproc,names", [(2, ["sod_1d", "shockbubble_2d", "shearlayer_2d", "shockdroplet_2d"]),
                                         (4, ["shockbubble_2d", "shearlayer_2d"]),
                                         (8, ["shockbubble_3d"])])
def test_halo_schedule_and_layouts_over_gloo(nproc, names):
    r = _torchrun(nproc, os.path.join(ROOT, "tests", "gloo_worker.py"), *names)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "GLOO_WORKER_OK" in r.stdout


def test_bench_reference_arm_under_torchrun_only_rank0_prints():
    r = _torchrun(2, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                  "--cpu-sample-cells", "32")
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["n_gpus"] == 2 and j["unit"] == "Mcell-steps/s" and j["value"] > 0
    assert j["cpu_baseline"]["kind"] == "port" and j["e2e"]["h2d_bytes_per_step"] == 0
