set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_extend.py -x -q > gpurun_out/gpu_tests_r01j.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/gpu_tests_r01j.log
AG2_TRACE=1 timeout 150 python bench.py --no-cpu-baseline --pagraph-reads 0 > gpurun_out/bench_r01j_1g.json 2> gpurun_out/bench_r01j_1g.err; echo "rc=$?"
tail -5 gpurun_out/bench_r01j_1g.err
AG2_TRACE=1 AG2_WS_STREAMED=536870912 timeout 150 python bench.py --no-cpu-baseline --pagraph-reads 0 > gpurun_out/bench_r01j_512m.json 2> gpurun_out/bench_r01j_512m.err; echo "rc=$?"
AG2_TRACE=1 AG2_WS_STREAMED=2147483648 timeout 150 python bench.py --no-cpu-baseline --pagraph-reads 0 > gpurun_out/bench_r01j_2g.json 2> gpurun_out/bench_r01j_2g.err; echo "rc=$?"
