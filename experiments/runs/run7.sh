set -x
mkdir -p gpurun_out
# 1. host-buffer run at full size: streamed form at 6 pair CTAs/SM (reads waited for on the host / inside the kernel, chunk
#    sizes, and 7 CTAs/SM with in-kernel waits = the stall and the library's own rerun), chunked form with balanced chunks
SW="AG2_E2E_PATH=streamed;AG2_E2E_PATH=streamed,AG2_STREAM_WAIT_KERNEL=1;AG2_E2E_PATH=streamed,AG2_STREAM_WAIT_KERNEL=1,AG2_WS_STREAMED=536870912"
SW="$SW;AG2_E2E_PATH=chunked;AG2_E2E_PATH=chunked,AG2_WS_STREAMED=3221225472;AG2_E2E_PATH=streamed,AG2_STREAM_WAIT_KERNEL=1,AG2_STREAM_CTAS_PER_SM=7"
AG2_TRACE=1 timeout 360 python bench.py --no-cpu-baseline --pagraph-reads 0 --steps 2 --e2e-sweep "$SW" > gpurun_out/bench_r01m_sweep.json 2> gpurun_out/bench_r01m_sweep.err; echo "sweep rc=$?"
grep -v "^\[ag2 trace\]" gpurun_out/bench_r01m_sweep.err | tail -12
# 2. parity: extension + the GPU tests not run yet in this session
timeout 60 python -m pytest tests/test_gpu_extend.py -x -q -m gpu > gpurun_out/gpu_tests_r01m_a.log 2>&1; echo "tests a rc=$?"
tail -3 gpurun_out/gpu_tests_r01m_a.log
timeout 240 python -m pytest tests/test_gpu_index.py tests/test_gpu_map.py tests/test_kmer_counter.py tests/test_pagraph_travel.py tests/test_gpu_pagraph.py -x -q -m gpu --durations=8 > gpurun_out/gpu_tests_r01m_b.log 2>&1; echo "tests b rc=$?"
tail -14 gpurun_out/gpu_tests_r01m_b.log
