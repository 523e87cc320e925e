#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_extend.py -m gpu -x -q > gpurun_out/gpu_tests_r02ar.log 2>&1
tail -3 gpurun_out/gpu_tests_r02ar.log
timeout 900 python bench.py --steps 3 --warmup 3 --full-reads 0 --pagraph-reads 0 --no-cpu-baseline --e2e-sweep "AG2_DUMMY=1;AG2_STREAM_CTAS_PER_SM=7;AG2_STREAM_CTAS_PER_SM=6;AG2_STREAM_FAT_KERNEL=1,AG2_STREAM_CTAS_PER_SM=6;AG2_DUMMY=2" > gpurun_out/bench_r02ar.json 2> gpurun_out/bench_r02ar.err
grep "bench sweep" gpurun_out/bench_r02ar.err | cut -c1-200
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02ar.json'))
print(d['value'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['e2e_ascii']['ms_per_step'])
PY
