set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_r01h.log 2>&1; echo "tests rc=$?"
for ws in 3221225472 7516192768 32212254720; do
  AG2_WS_STREAMED=$ws timeout 300 python bench.py --no-cpu-baseline --pagraph-reads 0 > gpurun_out/bench_r01h_ws$ws.json 2> gpurun_out/bench_r01h_ws$ws.err
done
AG2_STREAM_GRID=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01h.csv python bench.py --reads 50000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --pagraph-reads 2000 > gpurun_out/ncu_lh.log 2>&1
AG2_STREAM_GRID=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:xdrop_pair -c 1 -o gpurun_out/prof_r01h_pair python bench.py --reads 150000 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --pagraph-reads 0 > gpurun_out/ncu_ph.log 2>&1
tail -3 gpurun_out/gpu_tests_r01h.log
