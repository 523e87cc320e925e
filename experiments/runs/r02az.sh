#!/bin/bash
# round-end sequence as the driver runs it (GPU suite, smoke, both bench arms), then the ncu captures of the two dominant kernels
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_r02az.log 2>&1
tail -3 gpurun_out/gpu_tests_r02az.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02az.log 2>&1
tail -2 gpurun_out/smoke_r02az.log | cut -c1-200
( time timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref_r02az.json 2> gpurun_out/bench_ref_r02az.err ) 2>&1 | grep real
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02az.json 2> gpurun_out/bench_r02az.err ) 2>&1 | grep real
cut -c1-300 gpurun_out/bench_r02az.json
AG2_STREAM_GRID=0 timeout 900 ncu --set full --import-source on --clock-control none -k regex:xdrop_pair_kernel -c 1 -f -o gpurun_out/pair_r02az python bench.py --reads 150000 --steps 1 --warmup 0 --no-e2e --full-reads 0 --pagraph-reads 0 --no-cpu-baseline > gpurun_out/ncu_pair_r02az.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:seed_cta_kernel -c 1 -f -o gpurun_out/seed_r02az python experiments/seed_bench.py --reads 60000 --steps 1 > gpurun_out/ncu_seed_r02az.log 2>&1
ls -la gpurun_out/*r02az.ncu-rep
