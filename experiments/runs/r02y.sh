#!/bin/bash
# seed_cta_kernel occupancy sweep: threads per CTA x min CTAs per SM (register cap) x events cap
set -x
mkdir -p gpurun_out
for v in t320_m1 t384_m3 t512_m2 t512_m3; do
  AG2_B200_LIB=$PWD/experiments/variants/libag2_$v.so timeout 300 python experiments/seed_bench.py --reads 250000 --steps 3 > gpurun_out/seed_r02y_$v.log 2>&1
  echo $v; tail -1 gpurun_out/seed_r02y_$v.log | cut -c1-120
done
AG2_SEED_CAP=2880 AG2_B200_LIB=$PWD/experiments/variants/libag2_t256_m4.so timeout 300 python experiments/seed_bench.py --reads 250000 --steps 3 > gpurun_out/seed_r02y_t256_m4_cap2880.log 2>&1
tail -1 gpurun_out/seed_r02y_t256_m4_cap2880.log | cut -c1-400
AG2_SEED_CAP=2880 timeout 300 python experiments/seed_bench.py --reads 250000 --steps 3 > gpurun_out/seed_r02y_t256_m1_cap2880.log 2>&1
tail -1 gpurun_out/seed_r02y_t256_m1_cap2880.log | cut -c1-400
