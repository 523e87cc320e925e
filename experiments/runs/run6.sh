set -x
mkdir -p gpurun_out
# 1. parity: extension (all forms of the host-buffer run), the drop-in executable incl. the device-sharded runs
timeout 300 python -m pytest tests/test_gpu_extend.py tests/test_host_binary.py -x -q -m gpu > gpurun_out/gpu_tests_r01l.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/gpu_tests_r01l.log
# 2. smoke under ncu (kernels serialised: the consumer kernel must leave by itself) = launch list of the small run
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_smoke_r01l.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ncu_smoke_r01l.log 2>&1; echo "ncu smoke rc=$?"
tail -2 gpurun_out/ncu_smoke_r01l.log
# 3. host-buffer run at full size: chunk sizes of the chunked form, the streamed form (reads waited for on the host), spare CTAs
SW="AG2_E2E_PATH=chunked,AG2_WS_STREAMED=2147483648;AG2_E2E_PATH=chunked,AG2_WS_STREAMED=4294967296;AG2_E2E_PATH=chunked,AG2_WS_STREAMED=7516192768"
SW="$SW;AG2_E2E_PATH=streamed;AG2_E2E_PATH=streamed,AG2_STREAM_SPARE_CTAS=148;AG2_E2E_PATH=streamed,AG2_STREAM_SPARE_CTAS=148,AG2_STREAM_WAIT_KERNEL=1"
AG2_TRACE=1 timeout 360 python bench.py --no-cpu-baseline --pagraph-reads 0 --steps 2 --e2e-sweep "$SW" > gpurun_out/bench_r01l_sweep.json 2> gpurun_out/bench_r01l_sweep.err; echo "sweep rc=$?"
grep -v "^\[ag2 trace\]" gpurun_out/bench_r01l_sweep.err | tail -12
