set -x
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_r01d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r01d.log
python bench.py > gpurun_out/bench_r01d.log 2> gpurun_out/bench_r01d.err; echo "bench rc=$?" >> gpurun_out/bench_r01d.err
python tools/trace_tc.py > gpurun_out/trace_r01d.log 2>&1
python tools/prof_blur.py --stds 0.5,1,2,3,5,23 --iters 300 --sets 6 > gpurun_out/prof_r01d.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 2 --warmup 1 --no-swd --no-cpu-baseline > gpurun_out/ncu_l_r01d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blur_tc128 -c 3 -o gpurun_out/blur_tc_r01d -f python tools/prof_blur.py --stds 5 --iters 3 > gpurun_out/ncu_f_r01d.log 2>&1
ncu -i gpurun_out/blur_tc_r01d.ncu-rep --page raw --csv > gpurun_out/blur_tc_r01d_raw.csv 2>/dev/null
tail -3 gpurun_out/pytest_r01d.log; cat gpurun_out/prof_r01d.log
python tools/prof_blur.py --stds 0.5,5,23 --iters 300 --sets 6 --layout nhwc --rule tf >> gpurun_out/prof_r01d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blur_direct128 -c 2 -o gpurun_out/blur_direct_r01d -f python tools/prof_blur.py --stds 0.5 --iters 2 > gpurun_out/ncu_fd_r01d.log 2>&1
ncu -i gpurun_out/blur_direct_r01d.ncu-rep --page raw --csv > gpurun_out/blur_direct_r01d_raw.csv 2>/dev/null
