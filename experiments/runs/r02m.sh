set -x
timeout 900 python -m pytest tests/test_gpu_pagraph.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/gpu_tests_r02m.log
timeout 1500 python bench.py --only-pagraph --pagraph-reads 40000 > gpurun_out/pagraph_r02m_40k.json 2> gpurun_out/pagraph_r02m_40k.err
tail -3 gpurun_out/gpu_tests_r02m.log; tail -c 1500 gpurun_out/pagraph_r02m_40k.json
