set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_r01i.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/gpu_tests_r01i.log
for ws in 1073741824 3221225472; do
  AG2_WS_STREAMED=$ws timeout 300 python bench.py --no-cpu-baseline --pagraph-reads 0 > gpurun_out/bench_r01i_ws$ws.json 2> gpurun_out/bench_r01i_ws$ws.err
done
AG2_WS_STREAMED=536870912 timeout 300 python bench.py --no-cpu-baseline --pagraph-reads 0 > gpurun_out/bench_r01i_ws512m.json 2> gpurun_out/bench_r01i_ws512m.err
tail -2 gpurun_out/bench_r01i_ws*.err
