set -x
python -m pytest tests/test_gpu_map.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/gpu_tests_r02c.log
python experiments/seed_bench.py --reads 250000 --steps 3 > gpurun_out/seed_r02c.log 2>&1
AG2_STREAM_GRID=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02c.csv python experiments/seed_bench.py --reads 250000 --steps 1 > gpurun_out/ncu_launches_r02c.log 2>&1
tail -3 gpurun_out/gpu_tests_r02c.log; cat gpurun_out/seed_r02c.log
