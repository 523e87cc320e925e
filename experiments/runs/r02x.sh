set -x
timeout 900 python -m pytest tests/test_gpu_extend.py tests/test_gpu_map.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/gpu_tests_r02x.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --full-reads 0 --pagraph-reads 0 --no-cpu-baseline > gpurun_out/bench_r02x_value.json 2> gpurun_out/bench_r02x_value.err
timeout 300 python experiments/seed_bench.py --reads 250000 --steps 3 > gpurun_out/seed_r02x.log 2>&1
tail -3 gpurun_out/gpu_tests_r02x.log; tail -1 gpurun_out/seed_r02x.log | cut -c1-220
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02x_value.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'])
PY
