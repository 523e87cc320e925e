set -x
timeout 900 python bench.py --steps 3 --warmup 3 --full-reads 0 --pagraph-reads 0 --no-cpu-baseline --no-e2e-ascii \
  --e2e-sweep "AG2_STREAM_CTAS_PER_SM=7,AG2_STREAM_FREE_CTAS=37;AG2_STREAM_CTAS_PER_SM=7,AG2_STREAM_FREE_CTAS=74;AG2_STREAM_CTAS_PER_SM=7,AG2_STREAM_FREE_CTAS=148;AG2_STREAM_CTAS_PER_SM=7,AG2_STREAM_FREE_CTAS=296;AG2_STREAM_CTAS_PER_SM=7,AG2_STREAM_FREE_CTAS=0;AG2_E2E_PATH=chunked" \
  > gpurun_out/bench_r02g_sweep.json 2> gpurun_out/bench_r02g_sweep.err
grep "bench sweep" gpurun_out/bench_r02g_sweep.err
