#!/bin/bash
set -x
mkdir -p gpurun_out
SW=""; for i in $(seq 1 16); do SW="$SW;AG2_DUMMY=$i"; done
timeout 900 python bench.py --steps 2 --warmup 3 --full-reads 0 --pagraph-reads 0 --no-cpu-baseline --no-e2e-ascii --e2e-sweep "${SW:1}" > gpurun_out/bench_r02at.json 2> gpurun_out/bench_r02at.err
grep "bench sweep" gpurun_out/bench_r02at.err | grep -o '"ms_per_step": [0-9.]*' | cut -d' ' -f2 | cut -c1-6 | tr '\n' ' '
grep -c "gave up" gpurun_out/bench_r02at.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02at.json'))
print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'])
PY
