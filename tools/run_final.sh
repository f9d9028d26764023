# Round-1 (fourth session) closing check on one B200: every GPU test, smoke(), SWD timing with the final build.
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_r01g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r01g.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r01g.log 2>&1; echo "rc=$?" >> gpurun_out/smoke_r01g.log
timeout 200 python tools/swd_phases.py 16384 64 > gpurun_out/swd_phases_r01g.log 2>&1
tail -3 gpurun_out/pytest_r01g.log; tail -2 gpurun_out/smoke_r01g.log; tail -12 gpurun_out/swd_phases_r01g.log
