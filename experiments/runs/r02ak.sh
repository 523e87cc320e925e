#!/bin/bash
set -x
mkdir -p gpurun_out
SW=""; for i in $(seq 1 24); do SW="$SW;AG2_DUMMY=$i"; done
timeout 900 python bench.py --steps 1 --warmup 3 --full-reads 0 --pagraph-reads 0 --no-cpu-baseline --no-e2e-ascii --e2e-sweep "${SW:1}" > gpurun_out/bench_r02ak.json 2> gpurun_out/bench_r02ak.err
grep "bench sweep" gpurun_out/bench_r02ak.err | grep -o '"ms_per_step": [0-9.]*' | tr '\n' ' '
