timeout 600 python -m pytest tests/test_blur_gpu.py tests/test_wgan_harness.py -x -q -m gpu 2>&1 | tail -5
python tools/prof_blur.py --stds 2,5,23 --iters 300 --sets 6
python tools/prof_blur.py --stds 5 --iters 300 --sets 6 --layout nhwc --rule tf
python tools/trace_tc.py > gpurun_out/trace_sw128.log 2>&1
