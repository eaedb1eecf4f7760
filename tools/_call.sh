python bench.py --steps 10 --warmup 3 > gpurun_out/r02_v7_bench_512cube.json 2> gpurun_out/r2_c16_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_v7_bench_reference_arm.json 2> gpurun_out/r2_c16_ref.err
python bench.py --workload shockdroplet_2d_viscous_2048 --steps 20 --warmup 3 > gpurun_out/r02_v7_bench_viscous_2048sq.json 2> gpurun_out/r2_c16_visc.err
python bench.py --workload shockbubble_2d_4096 --steps 20 --warmup 3 > gpurun_out/r02_v7_bench_shockbubble_4096sq.json 2> gpurun_out/r2_c16_sb.err
python bench.py --workload advection_2d_1024 --steps 200 --warmup 20 > gpurun_out/r02_v7_bench_advection_1024sq.json 2> gpurun_out/r2_c16_adv.err
python bench.py --workload sod_1d_400 --steps 500 --warmup 20 > gpurun_out/r02_v7_bench_sod_400.json 2> gpurun_out/r2_c16_sod.err
python bench.py --stretched --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_v7_bench_512cube_stretched.json 2> gpurun_out/r2_c16_str.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_v7_launches_256cube.csv python bench.py --cells 256 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_c16_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_xstream|k_march3" -s 9 -c 6 -o gpurun_out/r02_v7_prof python bench.py --cells 256 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_c16_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_xstream|k_march3|k_vgrad" -s 9 -c 3 -o gpurun_out/r02_v7_prof_viscous python bench.py --workload shockdroplet_2d_viscous_2048 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_c16_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_v7_bench_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("r02_v7_bench_")[1], d["n_gpus"], round(d["value"], 1), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1) if d.get("e2e") else None,
              d.get("roofline", {}).get("frac"), d.get("cpu_baseline", {}) and d["cpu_baseline"].get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
