timeout 200 python -m pytest tests/test_swd_gpu.py -x -q -m gpu > gpurun_out/pytest_hist.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_hist.log
tail -2 gpurun_out/pytest_hist.log
timeout 100 python tools/sort_bench.py > gpurun_out/sort_bench_hist.log 2>&1; cat gpurun_out/sort_bench_hist.log
