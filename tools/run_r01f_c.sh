# Round-1 (fourth session), third GPU call: SWD parity after the fused feed / pyramid rewrite, then timing.
set -x
timeout 600 python -m pytest tests/test_swd_gpu.py -x -q -m gpu > gpurun_out/pytest_r01f_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r01f_c.log
timeout 300 python tools/swd_scale.py 16384 64 > gpurun_out/swd_scale_r01f.log 2>&1
timeout 300 python tools/swd_phases.py 16384 64 > gpurun_out/swd_phases_r01f.log 2>&1
tail -3 gpurun_out/pytest_r01f_c.log; cat gpurun_out/swd_scale_r01f.log; tail -20 gpurun_out/swd_phases_r01f.log
