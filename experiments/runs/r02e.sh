set -x
timeout 900 python -m pytest tests/test_gpu_extend.py tests/test_gpu_map.py tests/test_gpu_index.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/gpu_tests_r02e.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --full-reads 0 --pagraph-reads 0 --no-cpu-baseline > gpurun_out/bench_r02e_defer.json 2> gpurun_out/bench_r02e_defer.err
AG2_NO_DEFER=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --full-reads 0 --pagraph-reads 0 --no-cpu-baseline > gpurun_out/bench_r02e_nodefer.json 2> gpurun_out/bench_r02e_nodefer.err
timeout 300 python experiments/seed_bench.py --reads 250000 --steps 3 > gpurun_out/seed_r02e.log 2>&1
tail -4 gpurun_out/gpu_tests_r02e.log; tail -2 gpurun_out/seed_r02e.log
