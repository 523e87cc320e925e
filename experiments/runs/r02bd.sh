#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_pagraph.py tests/test_kmer_counter.py tests/test_gpu_index.py -m gpu -x -q > gpurun_out/sanitizer_memcheck_pg_r02bd.log 2>&1
echo "memcheck pagraph/kmer/index rc=$?"
grep "ERROR SUMMARY\|passed\|failed" gpurun_out/sanitizer_memcheck_pg_r02bd.log | tail -4
