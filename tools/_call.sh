python -m pytest tests -m gpu -q -s > gpurun_out/r2_c5_tests_full.log 2>&1; tail -5 gpurun_out/r2_c5_tests_full.log; grep FASTLINF gpurun_out/r2_c5_tests_full.log > gpurun_out/r2_c5_fastlinf.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_c5_bench.json 2> gpurun_out/r2_c5_bench.err
for w in advection_2d_1024 sod_1d_400; do
  python bench.py --workload $w --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2_c5_${w}_graph.json 2> gpurun_out/r2_c5_${w}_graph.err
  MFC_B200_GRAPH=0 python bench.py --workload $w --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2_c5_${w}_nograph.json 2> gpurun_out/r2_c5_${w}_nograph.err
done
python bench.py --workload shockbubble_2d_4096 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c5_sb4096.json 2> gpurun_out/r2_c5_sb4096.err
python bench.py --workload shockdroplet_2d_viscous_2048 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c5_visc2048.json 2> gpurun_out/r2_c5_visc2048.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_c5_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("r2_c5_")[1], round(d["value"], 1), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1) if d.get("e2e") else None,
              {k: round(v["avg_launch_ms"], 4) for k, v in d["roofline"]["per_kernel"].items()}, d["roofline"]["peak"], d["gpu_launches"])
    except Exception as e:
        print(f, "ERR", e)
PY
