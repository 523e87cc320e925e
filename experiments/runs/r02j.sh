set -x
timeout 1500 python -m pytest tests/test_host_binary.py tests/test_gpu_pagraph.py tests/test_kmer_counter.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/gpu_tests_r02j.log
timeout 1200 python bench.py --exec 100000 > gpurun_out/exec_r02j.json 2> gpurun_out/exec_r02j.err
tail -5 gpurun_out/gpu_tests_r02j.log; tail -c 1500 gpurun_out/exec_r02j.json
