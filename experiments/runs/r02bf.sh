#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_extend.py tests/test_gpu_map.py -m gpu -x -q > gpurun_out/gpu_tests_r02bf.log 2>&1
tail -3 gpurun_out/gpu_tests_r02bf.log
