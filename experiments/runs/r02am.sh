#!/bin/bash
set -x
mkdir -p gpurun_out
for g in 16 32 48; do
AG2_STREAM_GRID=$g AG2_TRACE=1 timeout 300 python experiments/seed_bench.py --reads 250000 --steps 4 > gpurun_out/seed_r02am_$g.log 2> gpurun_out/seed_r02am_$g.err
echo "grid $g"; grep -o '"total_ms": [0-9.]*\|"extend_ms": [0-9.]*\|"pair_kernel_ms": [0-9.]*' gpurun_out/seed_r02am_$g.log | tr '\n' ' '; echo
grep "pair kernel + consumer" gpurun_out/seed_r02am_$g.err | awk '$NF=="ms" && $(NF-1)>50' | tr '\n' ' '; echo
done
