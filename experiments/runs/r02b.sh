set -x
python -m pytest tests/test_gpu_map.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/gpu_tests_r02b.log
python experiments/seed_bench.py --reads 250000 --steps 3 > gpurun_out/seed_r02b.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:seed_cta_kernel -c 1 -f -o gpurun_out/seed_r02b python experiments/seed_bench.py --reads 60000 --steps 1 > gpurun_out/ncu_seed_r02b.log 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipes3 experiments/pipes/pipes3.cu && /tmp/pipes3 > gpurun_out/pipes3_r02b.txt 2>&1
tail -3 gpurun_out/gpu_tests_r02b.log; cat gpurun_out/seed_r02b.log; tail -3 gpurun_out/ncu_seed_r02b.log; head -20 gpurun_out/pipes3_r02b.txt
