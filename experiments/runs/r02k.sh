set -x
timeout 900 python -m pytest tests/test_gpu_map.py tests/test_gpu_index.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/gpu_tests_r02k.log
timeout 300 python experiments/seed_bench.py --reads 250000 --steps 3 > gpurun_out/seed_r02k.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:seed_cta_kernel -c 1 -f -o gpurun_out/seed_r02k python experiments/seed_bench.py --reads 60000 --steps 1 > gpurun_out/ncu_seed_r02k.log 2>&1
AG2_STREAM_GRID=0 timeout 900 ncu --set full --import-source on --clock-control none -k regex:xdrop_pair_kernel -c 1 -f -o gpurun_out/pair_r02k python bench.py --reads 150000 --steps 1 --warmup 0 --no-e2e --full-reads 0 --pagraph-reads 0 --no-cpu-baseline > gpurun_out/ncu_pair_r02k.log 2>&1
timeout 900 python bench.py --exec 300000 --exec-no-reference > gpurun_out/exec_r02k_300k.json 2> gpurun_out/exec_r02k_300k.err
tail -3 gpurun_out/gpu_tests_r02k.log; tail -1 gpurun_out/seed_r02k.log | cut -c1-200; tail -c 900 gpurun_out/exec_r02k_300k.json
