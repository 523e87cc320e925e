set -x
mkdir -p gpurun_out
# 1. the official line: library default for the host-buffer run, CPU baseline and the graph stage beside it
timeout 420 python bench.py > gpurun_out/bench_r01k_500k.json 2> gpurun_out/bench_r01k_500k.err; echo "default rc=$?"
tail -3 gpurun_out/bench_r01k_500k.err
# 2. the streamed form of the host-buffer run at the full size
AG2_TRACE=1 timeout 200 python bench.py --e2e-path streamed --no-cpu-baseline --pagraph-reads 0 > gpurun_out/bench_r01k_streamed.json 2> gpurun_out/bench_r01k_streamed.err; echo "streamed rc=$?"
tail -5 gpurun_out/bench_r01k_streamed.err
# 3. parity of the extension (both forms)
timeout 240 python -m pytest tests/test_gpu_extend.py -x -q > gpurun_out/gpu_tests_r01k.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/gpu_tests_r01k.log
# 4. reference arm
timeout 150 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_r01k.json 2> gpurun_out/bench_ref_r01k.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r01k.log 2>&1; echo "smoke rc=$?"
