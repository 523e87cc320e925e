set -x
timeout 900 python -m pytest tests/test_gpu_pagraph.py tests/test_kmer_counter.py -x -q -m gpu 2>&1 | tail -6 > gpurun_out/gpu_tests_r02n_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --only-pagraph --pagraph-reads 40000 > gpurun_out/pagraph_r02n_40k_2gpu.json 2> gpurun_out/pagraph_r02n_40k_2gpu.err
tail -4 gpurun_out/gpu_tests_r02n_2gpu.log; tail -c 1200 gpurun_out/pagraph_r02n_40k_2gpu.json; tail -3 gpurun_out/pagraph_r02n_40k_2gpu.err
