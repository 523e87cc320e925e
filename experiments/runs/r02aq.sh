#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_extend.py tests/test_gpu_map.py tests/test_gpu_index.py -m gpu -x -q > gpurun_out/gpu_tests_r02aq.log 2>&1
tail -15 gpurun_out/gpu_tests_r02aq.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --full-reads 0 --pagraph-reads 0 --no-cpu-baseline > gpurun_out/bench_r02aq_value.json 2> gpurun_out/bench_r02aq_value.err
cut -c1-220 gpurun_out/bench_r02aq_value.json
AG2_TRACE=1 timeout 300 python experiments/seed_bench.py --reads 250000 --steps 5 > gpurun_out/seed_r02aq.log 2> gpurun_out/seed_r02aq.err
grep -o '"total_ms": [0-9.]*\|"extend_ms": [0-9.]*' gpurun_out/seed_r02aq.log | tr '\n' ' '
grep "pair kernel + consumer" gpurun_out/seed_r02aq.err | awk '$(NF-1)>50' | tr '\n' ' '
