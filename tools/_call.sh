python -m pytest tests/test_multigpu.py -m gpu -q -x -k "2-" > gpurun_out/r2_c8_mgpu2.log 2>&1; grep -h "world=\|passed\|failed\|Error\|MISMATCH" gpurun_out/r2_c8_mgpu2.log | cut -c1-200 | tail -50
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c8_bench2.json 2> gpurun_out/r2_c8_bench2.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_c8_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("r2_c8_")[1], round(d["value"], 1), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1) if d.get("e2e") else None,
              {k: round(v["avg_launch_ms"], 4) for k, v in d["roofline"]["per_kernel"].items()}, d["roofline"]["peak"], d["gpu_launches"])
    except Exception as e:
        print(f, "ERR", e)
PY
