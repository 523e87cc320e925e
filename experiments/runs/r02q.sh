set -x
timeout 600 python experiments/two_ctx.py --reads 250000 > gpurun_out/two_ctx_r02q.log 2>&1
AG2_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 3 --full-reads 0 --pagraph-reads 0 --no-cpu-baseline --no-e2e-ascii > gpurun_out/bench_r02q_trace.json 2> gpurun_out/bench_r02q_trace.err
cat gpurun_out/two_ctx_r02q.log | tail -4; grep "ag2 trace" gpurun_out/bench_r02q_trace.err | tail -2 | cut -c1-600
