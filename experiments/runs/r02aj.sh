#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_extend.py tests/test_gpu_map.py -m gpu -x -q > gpurun_out/gpu_tests_r02aj.log 2>&1
tail -3 gpurun_out/gpu_tests_r02aj.log
timeout 900 python bench.py --steps 3 --warmup 3 --full-reads 0 --pagraph-reads 0 --no-cpu-baseline --e2e-sweep "AG2_DUMMY=1;AG2_DUMMY=2;AG2_STREAM_CTAS_PER_SM=5;AG2_STREAM_CTAS_PER_SM=7;AG2_STREAM_CTAS_PER_SM=8;AG2_E2E_PATH=chunked" > gpurun_out/bench_r02aj.json 2> gpurun_out/bench_r02aj.err
grep "bench sweep" gpurun_out/bench_r02aj.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02aj.json'))
print(d['value'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['e2e_ascii']['ms_per_step'])
PY
