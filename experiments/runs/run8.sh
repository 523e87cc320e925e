set -x
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_extend.py -x -q -m gpu > gpurun_out/gpu_tests_r01n.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/gpu_tests_r01n.log
# the official line: library defaults, CPU baseline and the graph stage beside it
AG2_TRACE=1 timeout 420 python bench.py > gpurun_out/bench_r01n_500k.json 2> gpurun_out/bench_r01n_500k.err; echo "default rc=$?"
tail -4 gpurun_out/bench_r01n_500k.err | cut -c1-400
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r01n.log 2>&1; echo "smoke rc=$?"
