set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline --pagraph-reads 0 > gpurun_out/bench_r01n_2gpu.json 2> gpurun_out/bench_r01n_2gpu.err; echo "2gpu rc=$?"
tail -3 gpurun_out/bench_r01n_2gpu.err | cut -c1-300
cut -c1-600 gpurun_out/bench_r01n_2gpu.json
