python -m pytest tests -m gpu -q > gpurun_out/r2_c20_tests.log 2>&1; tail -3 gpurun_out/r2_c20_tests.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/r02_v8_bench_512cube.json 2> gpurun_out/r2_c20_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02_v8_bench_512cube.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel'], d['roofline']['traffic'], d['e2e']['value'], d['cpu_baseline']['value'], d['gpu_launches'])"
