set -x
mkdir -p gpurun_out
K='regex:^(xdrop|seed|assemble|extend_|pack_reads|exclusive_scan|scan_|chain_keys|stream_keys|set_slots|ascii_kmer|ref_kmer|mask_counts|sort_buckets|sum_kcount|vote_|plan_|finish_|rescue_|gather_|flags_|compact_|widen_|seeds_to|DeviceRadixSort)'
AG2_STREAM_GRID=0 timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --csv --log-file gpurun_out/launches_r01n.csv python bench.py --reads 40000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --pagraph-reads 0 > gpurun_out/ncu_r01n.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/ncu_r01n.log | cut -c1-300
wc -l gpurun_out/launches_r01n.csv
