#!/usr/bin/env python
"""Build libmfc_b200 variants with different occupancy / ring knobs (in parallel, here) and
print the gpurun command line that times them all on one box.

    python tools/tune_variants.py build        # -> microfc_b200/libmfc_tune_<name>.so
    python tools/tune_variants.py bench        # on the GPU box: runs bench.py per variant
"""
import json
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = {
    # name: (RING_Y, WARPS_Y, CTAS_Y, RING_X, WARPS_X, CTAS_X, extra -D flags)
    # (v4: k_march3 adds 2-3 operand / output slots per warp to RING_Y / RING_Z; 72 KB per CTA at 3 CTAs)
    "x_r3c6": (7, 4, 3, 3, 4, 6, ""),
    "x_r3c7": (7, 4, 3, 3, 4, 7, ""),
}


def defs(v):
    ry, wy, cy, rx, wx, cx, extra = v
    return f"-DMFC_RING_Y={ry} -DMFC_WARPS_Y={wy} -DMFC_CTAS_Y={cy} -DMFC_RING_X={rx} -DMFC_WARPS_X={wx} -DMFC_CTAS_X={cx} {extra}"


def build_one(name):
    env = dict(os.environ, MFC_B200_DEFS=defs(VARIANTS[name]), MFC_B200_LIBNAME=f"libmfc_tune_{name}.so")
    r = subprocess.run([sys.executable, "-m", "microfc_b200.build", "--force"], cwd=ROOT, env=env, capture_output=True, text=True)
    return name, r.returncode, r.stdout[-300:] + r.stderr[-2000:]


def main():
    if sys.argv[1] == "build":
        with ThreadPoolExecutor(max_workers=2) as ex:
            for name, rc, out in ex.map(build_one, VARIANTS):
                print(name, "ok" if rc == 0 else "FAILED\n" + out)
    else:
        cells = sys.argv[2] if len(sys.argv) > 2 else "512"
        for name in VARIANTS:
            lib = os.path.join(ROOT, "microfc_b200", f"libmfc_tune_{name}.so")
            if not os.path.exists(lib):
                continue
            env = dict(os.environ, MFC_B200_LIB=lib)
            r = subprocess.run([sys.executable, "bench.py", "--steps", "4", "--warmup", "3", "--no-cpu-baseline", "--no-e2e", "--cells", cells],
                               cwd=ROOT, env=env, capture_output=True, text=True)
            try:
                j = json.loads(r.stdout.strip().splitlines()[-1])
                kt = {k: round(v["seconds"] / v["launches"] * 1e3, 3) for k, v in j["kernel_time"].items()}
                print(name, round(j["value"], 1), "Mcell-steps/s", round(j["ms_per_step"], 2), "ms", kt, flush=True)
            except Exception as e:
                print(name, "FAILED", e, r.stderr[-500:], flush=True)


if __name__ == "__main__":
    main()
