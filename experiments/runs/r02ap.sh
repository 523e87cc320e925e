#!/bin/bash
# round-end sequence as the driver runs it: GPU suite, smoke, both bench arms; then the ncu launch list of a short bench run
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_r02ap.log 2>&1
tail -3 gpurun_out/gpu_tests_r02ap.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02ap.log 2>&1
tail -2 gpurun_out/smoke_r02ap.log | cut -c1-300
( time timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref_r02ap.json 2> gpurun_out/bench_ref_r02ap.err ) 2>&1 | grep real
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02ap.json 2> gpurun_out/bench_r02ap.err ) 2>&1 | grep real
cut -c1-300 gpurun_out/bench_r02ap.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02ap.csv python bench.py --reads 20000 --steps 2 --warmup 1 --no-e2e --full-reads 20000 --full-steps 1 --pagraph-reads 0 --no-cpu-baseline > gpurun_out/ncu_launches_r02ap.log 2>&1
tail -2 gpurun_out/ncu_launches_r02ap.log | cut -c1-200
