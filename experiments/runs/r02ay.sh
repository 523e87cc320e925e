#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_extend.py tests/test_gpu_map.py tests/test_gpu_index.py -m gpu -x -q > gpurun_out/gpu_tests_r02ay.log 2>&1
tail -3 gpurun_out/gpu_tests_r02ay.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --full-reads 0 --pagraph-reads 0 --no-cpu-baseline > gpurun_out/bench_r02ay_value.json 2> gpurun_out/bench_r02ay_value.err
cut -c1-220 gpurun_out/bench_r02ay_value.json
timeout 300 python experiments/seed_bench.py --reads 250000 --steps 3 > gpurun_out/seed_r02ay.log 2>&1
tail -1 gpurun_out/seed_r02ay.log | cut -c1-420
