set -x
timeout 1700 python bench.py --only-pagraph --pagraph-reads 200000 > gpurun_out/pagraph_r02v_200k.json 2> gpurun_out/pagraph_r02v_200k.err
tail -c 1100 gpurun_out/pagraph_r02v_200k.json; tail -3 gpurun_out/pagraph_r02v_200k.err
