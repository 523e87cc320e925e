set -x
timeout 900 python -m pytest tests/test_gpu_extend.py tests/test_gpu_map.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/gpu_tests_r02f.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --full-reads 0 --pagraph-reads 0 --no-cpu-baseline > gpurun_out/bench_r02f_defer.json 2> gpurun_out/bench_r02f_defer.err
timeout 600 python bench.py --steps 5 --warmup 3 --pagraph-reads 0 > gpurun_out/bench_r02f.json 2> gpurun_out/bench_r02f.err
tail -4 gpurun_out/gpu_tests_r02f.log; tail -3 gpurun_out/bench_r02f.err
