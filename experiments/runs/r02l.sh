set -x
timeout 1500 python bench.py --only-pagraph --pagraph-reads 40000 > gpurun_out/pagraph_r02l_40k.json 2> gpurun_out/pagraph_r02l_40k.err
tail -c 1800 gpurun_out/pagraph_r02l_40k.json; tail -5 gpurun_out/pagraph_r02l_40k.err
