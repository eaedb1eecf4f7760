python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -s -k "visc or droplet" > gpurun_out/r2_c9_tests.log 2>&1; grep -h "FASTLINF\|passed\|failed\|^E  " gpurun_out/r2_c9_tests.log | cut -c1-250 | tail -30
python bench.py --workload shockdroplet_2d_viscous_2048 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c9_visc2048.json 2> gpurun_out/r2_c9_visc2048.err
MFC_B200_VISC_FUSED=0 python bench.py --workload shockdroplet_2d_viscous_2048 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c9_visc2048_unfused.json 2> gpurun_out/r2_c9_visc2048_unfused.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_c9_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("r2_c9_")[1], round(d["value"], 1), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1) if d.get("e2e") else None,
              {k: round(v["seconds"]/v["launches"]*1e3, 4) for k, v in d["kernel_time"].items()}, d["gpu_launches"])
    except Exception as e:
        print(f, "ERR", e)
PY
