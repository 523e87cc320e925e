set -x
free -g | head -2; df -h /tmp /dev/shm | tail -2
timeout 1700 python bench.py --only-pagraph --pagraph-reads 200000 > gpurun_out/pagraph_r02r_200k.json 2> gpurun_out/pagraph_r02r_200k.err
tail -c 1500 gpurun_out/pagraph_r02r_200k.json; tail -5 gpurun_out/pagraph_r02r_200k.err
