#!/bin/bash
# compute-sanitizer: memcheck over the parity tests that reach every kernel path, racecheck (shared memory) over the small end-to-end run
set -x
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_extend.py tests/test_gpu_map.py -m gpu -x -q -k "golden or repeats or late_hand or packed or tier or edge" > gpurun_out/sanitizer_memcheck_tests_r02bb.log 2>&1
echo "memcheck tests rc=$?"
tail -6 gpurun_out/sanitizer_memcheck_tests_r02bb.log | cut -c1-300
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck_r02bb.log 2>&1
echo "racecheck rc=$?"
tail -6 gpurun_out/sanitizer_racecheck_r02bb.log | cut -c1-300
