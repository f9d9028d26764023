set -x
timeout 300 python -m pytest tests/test_swd_gpu.py tests/test_wgan_harness.py -x -q -m gpu > gpurun_out/pytest_r01h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r01h.log
timeout 200 python tools/swd_phases.py 16384 64 > gpurun_out/swd_phases_r01h.log 2>&1
tail -3 gpurun_out/pytest_r01h.log; head -2 gpurun_out/swd_phases_r01h.log
