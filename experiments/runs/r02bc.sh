#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck_r02bc.log 2>&1
echo "racecheck smoke rc=$?"
grep "RACECHECK SUMMARY" gpurun_out/sanitizer_racecheck_r02bc.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_map.py -m gpu -x -q -k "tier" > gpurun_out/sanitizer_racecheck_map_r02bc.log 2>&1
echo "racecheck map rc=$?"
grep "RACECHECK SUMMARY\|passed\|failed" gpurun_out/sanitizer_racecheck_map_r02bc.log
grep "Race reported" -A1 gpurun_out/sanitizer_racecheck_map_r02bc.log | grep -o "[a-z_]*\.cuh:[0-9]*" | sort | uniq -c
