# Round-1 (fourth session), first GPU call: new parity tests, MMA rate micro-benchmark, multi-launch timing.
set -x
timeout 120 tools/mma_rate.bin > gpurun_out/mma_rate_r01f.log 2>&1; echo "rc=$?" >> gpurun_out/mma_rate_r01f.log
timeout 600 python -m pytest tests/test_blur_gpu.py tests/test_wgan_harness.py -x -q -m gpu -k "multi or lerp or fused or plane_counts or golden or celeba_shape_nchw" > gpurun_out/pytest_r01f_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r01f_a.log
timeout 200 python tools/prof_blur.py --stds 1,5,23 --iters 40 --sets 6 --multi 2 > gpurun_out/prof_multi_r01f.log 2>&1
timeout 200 python tools/prof_blur.py --stds 1,5 --iters 40 --sets 6 --multi 3 --shape 32,3,128,128 >> gpurun_out/prof_multi_r01f.log 2>&1
tail -3 gpurun_out/pytest_r01f_a.log; cat gpurun_out/mma_rate_r01f.log; cat gpurun_out/prof_multi_r01f.log
