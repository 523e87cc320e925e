#!/bin/bash
# full GPU suite + driver-equivalent bench lines (both arms)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_r02ai.log 2>&1
tail -3 gpurun_out/gpu_tests_r02ai.log
timeout 900 python bench.py > gpurun_out/bench_r02ai.json 2> gpurun_out/bench_r02ai.err
cut -c1-300 gpurun_out/bench_r02ai.json
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_r02ai.json 2> gpurun_out/bench_ref_r02ai.err
cut -c1-300 gpurun_out/bench_ref_r02ai.json
