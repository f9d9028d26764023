python tools/prof_blur.py --stds 0.5,1,5,23 --iters 300 --sets 6
timeout 300 python -m pytest tests/test_blur_gpu.py -x -q -m gpu -k "tensor_core or celeba or direct" 2>&1 | tail -2
python bench.py --no-swd --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1
