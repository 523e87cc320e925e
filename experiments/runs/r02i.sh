set -x
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/gpu_tests_r02i.log
tail -6 gpurun_out/gpu_tests_r02i.log
