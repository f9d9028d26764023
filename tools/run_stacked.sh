# stacked-GEMM2 variant of the tensor-core blur (BGAN_TC_STACKED=1): parity, then launch time against the default
set -x
BGAN_TC_STACKED=1 timeout 200 python -m pytest tests/test_blur_gpu.py -x -q -m gpu -k "tensor_core or plane_counts or celeba_shape or golden or multi_tensor_launch or full_size or randomised" > gpurun_out/pytest_stacked.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_stacked.log
tail -4 gpurun_out/pytest_stacked.log
BGAN_TC_STACKED=1 timeout 100 python tools/prof_blur.py --stds 5,23 --iters 60 --sets 6 > gpurun_out/prof_stacked.log 2>&1
timeout 100 python tools/prof_blur.py --stds 5,23 --iters 60 --sets 6 >> gpurun_out/prof_stacked.log 2>&1
BGAN_TC_STACKED=1 timeout 100 python tools/prof_blur.py --stds 5 --iters 60 --sets 6 --layout nhwc --rule tf >> gpurun_out/prof_stacked.log 2>&1
cat gpurun_out/prof_stacked.log
