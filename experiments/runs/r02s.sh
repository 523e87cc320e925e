set -x
AG2_PG_TRACE=1 timeout 900 python bench.py --only-pagraph --pagraph-reads 40000 > gpurun_out/pagraph_r02s_trace.json 2> gpurun_out/pagraph_r02s_trace.err
grep "ag2_pg trace" gpurun_out/pagraph_r02s_trace.err | tail -24
timeout 1200 python -m pytest tests/test_host_binary.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/gpu_tests_r02s.log
timeout 900 python bench.py --exec 300000 --exec-no-reference > gpurun_out/exec_r02s_300k.json 2> gpurun_out/exec_r02s_300k.err
tail -3 gpurun_out/gpu_tests_r02s.log; tail -c 1200 gpurun_out/exec_r02s_300k.json
