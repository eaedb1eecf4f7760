set -x
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
# multi-GPU parity: 4 and 8 ranks (per-case lines)
run 4 29601 tests/nccl_worker.py shockbubble_2d shearlayer_2d shockbubble_3d viscous_2d shockdroplet_2d_viscous viscous_wave_2d_weno viscous_wave_2d_fd shockbubble_3d_yzsplit shockbubble_3d_xysplit > gpurun_out/r02_multigpu_4rank.log 2>&1
run 8 29602 tests/nccl_worker.py shockbubble_3d shockbubble_2d_4x2 viscous_wave_2d_4x2 > gpurun_out/r02_multigpu_8rank.log 2>&1
grep -h "world=\|NCCL_WORKER" gpurun_out/r02_multigpu_4rank.log gpurun_out/r02_multigpu_8rank.log
# weak scaling, default workload, 8 GPUs
run 8 29603 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_8gpu_512cube.json 2> gpurun_out/r2_c12_b8.err
# configs[3]: viscous shock-droplet 8192 x 4096 on 8 GPUs
run 8 29604 bench.py --gpus 8 --workload shockdroplet_2d_viscous_2048 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_8gpu_viscous_8192x4096.json 2> gpurun_out/r2_c12_v8.err
# configs[2]: one 4096^2 shock-bubble grid, strong-scaled
for n in 8 4 2; do
  run $n 2961$n bench.py --gpus $n --workload shockbubble_2d_4096 --scaling strong --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_strong_${n}gpu_shockbubble_4096sq.json 2> gpurun_out/r2_c12_s$n.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_*gpu*.json")):
    try:
        d = json.load(open(f))
        print(f.split("r02_bench_")[1], d["n_gpus"], d["scaling"], round(d["value"], 1), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1) if d.get("e2e") else None,
              {k: round(v["seconds"]/v["launches"]*1e3, 4) for k, v in d["kernel_time"].items()}, d["config"]["decomposition"])
    except Exception as e:
        print(f, "ERR", e)
PY
