set -x
nvidia-smi topo -m > gpurun_out/topo_r02h.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_pagraph.py -x -q -m gpu -k "group_exchange or two_gpus" 2>&1 | tail -8 > gpurun_out/gpu_tests_r02h_pagraph.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 --pagraph-reads 0 > gpurun_out/bench_r02h_2gpu.json 2> gpurun_out/bench_r02h_2gpu.err
tail -5 gpurun_out/gpu_tests_r02h_pagraph.log; tail -5 gpurun_out/bench_r02h_2gpu.err
