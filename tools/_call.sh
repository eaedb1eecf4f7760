run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
run 2 29701 tests/nccl_worker.py shockbubble_2d shockbubble_3d shockbubble_3d_zsplit advection_2d_ysplit > gpurun_out/r2_c14_worker.log 2>&1
grep -h "world=\|NCCL_WORKER\|Error\|error" gpurun_out/r2_c14_worker.log | head -50
run 2 29702 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c14_bench2.json 2> gpurun_out/r2_c14_bench2.err
MFC_B200_XPIECES=2 run 2 29703 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_c14_bench2_p2.json 2> gpurun_out/r2_c14_bench2_p2.err
MFC_B200_XPIECES=8 run 2 29704 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_c14_bench2_p8.json 2> gpurun_out/r2_c14_bench2_p8.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_c14_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("r2_c14_")[1], d["n_gpus"], round(d["value"], 1), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1) if d.get("e2e") else None,
              {k: round(v["seconds"]/v["launches"]*1e3, 4) for k, v in d["kernel_time"].items()}, d["config"]["decomposition"])
    except Exception as e:
        print(f, "ERR", e)
PY
