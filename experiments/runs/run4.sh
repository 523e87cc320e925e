set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_extend.py -x -q > gpurun_out/gpu_tests_r01k.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/gpu_tests_r01k.log
AG2_TRACE=1 timeout 150 python bench.py --no-cpu-baseline --pagraph-reads 0 > gpurun_out/bench_r01k_1g.json 2> gpurun_out/bench_r01k_1g.err; echo "rc=$?"
tail -5 gpurun_out/bench_r01k_1g.err
AG2_TRACE=1 AG2_STREAM_WAIT_HOST=1 timeout 150 python bench.py --no-cpu-baseline --pagraph-reads 0 > gpurun_out/bench_r01k_hostwait.json 2> gpurun_out/bench_r01k_hostwait.err; echo "rc=$?"
tail -5 gpurun_out/bench_r01k_hostwait.err
