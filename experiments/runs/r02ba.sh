#!/bin/bash
# compute-sanitizer memcheck over the small end-to-end run (smoke: extension on the pair kernel + whole per-read path)
set -x
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_r02ba.log 2>&1
echo "memcheck rc=$?"
tail -12 gpurun_out/sanitizer_memcheck_r02ba.log | cut -c1-300
