#!/bin/bash
# final state of the round: the whole GPU suite
set -x
mkdir -p gpurun_out
timeout 280 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_r02bi.log 2>&1
tail -3 gpurun_out/gpu_tests_r02bi.log
