python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r2_c19_bench2.json 2> gpurun_out/r2_c19_bench2.err
tail -3 gpurun_out/r2_c19_bench2.err | cut -c1-300
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_c19_bench2.json"))
print(d["n_gpus"], round(d["value"], 1), round(d["ms_per_step"], 3), d["cpu_baseline"], d["e2e"]["value"])
PY
