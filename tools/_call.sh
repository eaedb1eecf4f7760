python -m pytest tests -m gpu -q -s > gpurun_out/r2_c11_tests_full.log 2>&1; tail -3 gpurun_out/r2_c11_tests_full.log | cut -c1-300; grep "FASTLINF\|SIZELINF" gpurun_out/r2_c11_tests_full.log > gpurun_out/r2_c11_linf.txt; grep SIZELINF gpurun_out/r2_c11_linf.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_c11_bench.json 2> gpurun_out/r2_c11_bench.err
python bench.py --stretched --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_c11_stretched.json 2> gpurun_out/r2_c11_stretched.err
python bench.py --workload shockdroplet_2d_viscous_2048 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c11_visc2048.json 2> gpurun_out/r2_c11_visc2048.err
python bench.py --workload shockbubble_2d_4096 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c11_sb4096.json 2> gpurun_out/r2_c11_sb4096.err
python bench.py --workload advection_2d_1024 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2_c11_adv1024.json 2> gpurun_out/r2_c11_adv1024.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_c11_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("r2_c11_")[1], round(d["value"], 1), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1) if d.get("e2e") else None,
              {k: round(v["seconds"]/v["launches"]*1e3, 4) for k, v in d["kernel_time"].items()}, d["roofline"]["frac"], d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "ERR", e)
PY
