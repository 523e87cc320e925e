#!/bin/bash
set -x
mkdir -p gpurun_out
AG2_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-e2e --full-reads 0 --pagraph-reads 0 --no-cpu-baseline > gpurun_out/bench_r02al_value.json 2> gpurun_out/bench_r02al_value.err
grep "ag2 trace" gpurun_out/bench_r02al_value.err | tail -12
AG2_TRACE=1 timeout 300 python experiments/seed_bench.py --reads 250000 --steps 3 > gpurun_out/seed_r02al.log 2> gpurun_out/seed_r02al.err
tail -1 gpurun_out/seed_r02al.log | cut -c1-300
grep "ag2 trace" gpurun_out/seed_r02al.err | tail -24
