# Round-1 (fourth session) measurement run on one B200: tests, smoke, bench (both arms), ncu launch list, --set full capture.
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_r01f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r01f.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r01f.log 2>&1; echo "rc=$?" >> gpurun_out/smoke_r01f.log
python bench.py > gpurun_out/bench_r01f.log 2> gpurun_out/bench_r01f.err; echo "bench rc=$?" >> gpurun_out/bench_r01f.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01f.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r01f.csv python bench.py --steps 2 --warmup 1 --no-swd --no-cpu-baseline > gpurun_out/ncu_l_r01f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"blur_tc128" -c 3 -o gpurun_out/blur_multi_r01f -f python tools/prof_blur.py --stds 5 --iters 1 --multi 2 > gpurun_out/ncu_f_r01f.log 2>&1
ncu -i gpurun_out/blur_multi_r01f.ncu-rep --page raw --csv > gpurun_out/blur_multi_r01f_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_swd_r01f.csv python tools/swd_scale.py 2048 64 > gpurun_out/ncu_ls_r01f.log 2>&1
tail -3 gpurun_out/pytest_r01f.log; tail -2 gpurun_out/smoke_r01f.log; tail -c 600 gpurun_out/bench_r01f.log; tail -2 gpurun_out/bench_r01f.err
