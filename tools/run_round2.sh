# Round-2 measurement run on one B200: tests, smoke, bench (both arms), ncu launch lists, --set full captures of the three
# kernels the round worked on.  Every ncu command follows a plain run of the same command line that exited 0.
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_r02.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02.log 2>&1; echo "rc=$?" >> gpurun_out/smoke_r02.log
python bench.py > gpurun_out/bench_r02.log 2> gpurun_out/bench_r02.err; echo "bench rc=$?" >> gpurun_out/bench_r02.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02.log 2>&1
python bench.py --steps 2 --warmup 3 --no-swd --no-cpu-baseline --no-large --no-train-step > gpurun_out/plain_l.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bench_r02.csv python bench.py --steps 2 --warmup 3 --no-swd --no-cpu-baseline --no-large --no-train-step > gpurun_out/ncu_l_r02.log 2>&1
python tools/swd_scale.py 2048 64 > gpurun_out/plain_ls.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_swd_r02.csv python tools/swd_scale.py 2048 64 > gpurun_out/ncu_ls_r02.log 2>&1
python tools/prof_blur.py --stds 5 --iters 3 > gpurun_out/plain_f1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"blur_tc128" -s 1 -c 2 -o gpurun_out/blur_tc128_r02 -f python tools/prof_blur.py --stds 5 --iters 3 > gpurun_out/ncu_f1_r02.log 2>&1
python tools/prof_axis.py --n 8 --iters 2 --stds 8 > gpurun_out/plain_f2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"blur_axis_kernel" -s 4 -c 2 -o gpurun_out/axis_sweep_r02 -f python tools/prof_axis.py --n 8 --iters 2 --stds 8 > gpurun_out/ncu_f2_r02.log 2>&1
python tools/sort_bench.py 128 262144 > gpurun_out/plain_f3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"radix_scatter" -s 4 -c 4 -o gpurun_out/sort_scatter_r02 -f python tools/sort_bench.py 128 262144 > gpurun_out/ncu_f3_r02.log 2>&1
for r in blur_tc128_r02 axis_sweep_r02 sort_scatter_r02; do ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null; done
tail -3 gpurun_out/pytest_r02.log; tail -2 gpurun_out/smoke_r02.log; tail -c 400 gpurun_out/bench_r02.log; tail -2 gpurun_out/bench_r02.err
