python -m pytest tests -m gpu -q -x > gpurun_out/r2_c15_tests.log 2>&1; tail -3 gpurun_out/r2_c15_tests.log | cut -c1-300
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_c15_bench.json 2> gpurun_out/r2_c15_bench.err
MFC_B200_BCMAP=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_c15_bench_nomap.json 2> gpurun_out/r2_c15_bench_nomap.err
python bench.py --workload advection_2d_1024 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2_c15_adv1024.json 2> gpurun_out/r2_c15_adv1024.err
python bench.py --workload shockbubble_2d_4096 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_c15_sb4096.json 2> gpurun_out/r2_c15_sb4096.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_c15_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("r2_c15_")[1], d["n_gpus"], round(d["value"], 1), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1) if d.get("e2e") else None,
              {k: round(v["seconds"]/v["launches"]*1e3, 4) for k, v in d["kernel_time"].items()}, d["gpu_launches"])
    except Exception as e:
        print(f, "ERR", e)
PY
