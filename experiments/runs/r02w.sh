set -x
( time timeout 2400 python -m pytest tests -x -q -m gpu ) > gpurun_out/gpu_tests_r02w.log 2>&1
tail -6 gpurun_out/gpu_tests_r02w.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02w.log 2>&1; tail -3 gpurun_out/smoke_r02w.log | cut -c1-400
