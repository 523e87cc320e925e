set -x
date +%s > gpurun_out/t0_r02o.txt
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02o.json 2> gpurun_out/bench_r02o.err ) 2> gpurun_out/time_r02o_ours.txt
( time timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref_r02o.json 2> gpurun_out/bench_ref_r02o.err ) 2> gpurun_out/time_r02o_ref.txt
cat gpurun_out/time_r02o_ours.txt gpurun_out/time_r02o_ref.txt; tail -2 gpurun_out/bench_r02o.err; tail -2 gpurun_out/bench_ref_r02o.err
