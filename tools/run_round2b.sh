# Round-2 closing run on one B200: tests, smoke, bench (both arms), ncu launch lists of the bench step and of an SWD evaluation.
# Every ncu command follows a plain run of the same command line that exited 0.
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_r02z.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02z.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02z.log 2>&1; echo "rc=$?" >> gpurun_out/smoke_r02z.log
python bench.py > gpurun_out/bench_r02z.log 2> gpurun_out/bench_r02z.err; echo "bench rc=$?" >> gpurun_out/bench_r02z.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02z.log 2>&1
python bench.py --steps 2 --warmup 3 --no-swd --no-cpu-baseline --no-large --no-train-step > gpurun_out/plain_l.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bench_r02z.csv python bench.py --steps 2 --warmup 3 --no-swd --no-cpu-baseline --no-large --no-train-step > gpurun_out/ncu_l_r02z.log 2>&1
python tools/swd_scale.py 2048 64 > gpurun_out/plain_ls.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_swd_r02z.csv python tools/swd_scale.py 2048 64 > gpurun_out/ncu_ls_r02z.log 2>&1
python tools/prof_blur.py --stds 2 --iters 3 > gpurun_out/plain_f1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"blur_direct128" -s 1 -c 2 -o gpurun_out/blur_direct_r5_r02z -f python tools/prof_blur.py --stds 2 --iters 3 > gpurun_out/ncu_f1_r02z.log 2>&1
ncu -i gpurun_out/blur_direct_r5_r02z.ncu-rep --page raw --csv > gpurun_out/blur_direct_r5_r02z_raw.csv 2>/dev/null
tail -3 gpurun_out/pytest_r02z.log; tail -2 gpurun_out/smoke_r02z.log; tail -c 300 gpurun_out/bench_r02z.log; tail -2 gpurun_out/bench_r02z.err; tail -c 300 gpurun_out/bench_ref_r02z.log
