python -m pytest tests -m gpu -q -s > gpurun_out/r2_c21_tests.log 2>&1; tail -3 gpurun_out/r2_c21_tests.log | cut -c1-200; grep "FASTLINF three" gpurun_out/r2_c21_tests.log; grep -n "FAILED\|Error" gpurun_out/r2_c21_tests.log | head
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_c21_bench.json 2> gpurun_out/r2_c21_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2_c21_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"
