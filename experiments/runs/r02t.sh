set -x
timeout 900 python -m pytest tests/test_gpu_pagraph.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/gpu_tests_r02t.log
AG2_PG_TRACE=1 timeout 900 python bench.py --only-pagraph --pagraph-reads 40000 > gpurun_out/pagraph_r02t_40k.json 2> gpurun_out/pagraph_r02t_40k.err
tail -3 gpurun_out/gpu_tests_r02t.log; grep "ag2_pg trace" gpurun_out/pagraph_r02t_40k.err | tail -9; tail -c 900 gpurun_out/pagraph_r02t_40k.json
