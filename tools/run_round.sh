# Round-1 (third session) measurement run on one B200: tests, bench, traces, ncu launch lists and --set full captures.
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_r01e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r01e.log
python bench.py > gpurun_out/bench_r01e.log 2> gpurun_out/bench_r01e.err; echo "bench rc=$?" >> gpurun_out/bench_r01e.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01e.log 2>&1
python tools/swd_scale.py 16384 64 > gpurun_out/swd_scale_r01e.log 2>&1
python tools/swd_phases.py 16384 64 > gpurun_out/swd_phases_r01e.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_swd_r01e.csv python tools/swd_scale.py 2048 64 > gpurun_out/ncu_ls_r01e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"project_tc|radix_scatter|radix_hist" -c 6 -o gpurun_out/swd_r01e -f python tools/swd_scale.py 2048 64 > gpurun_out/ncu_fs_r01e.log 2>&1
ncu -i gpurun_out/swd_r01e.ncu-rep --page raw --csv > gpurun_out/swd_r01e_raw.csv 2>/dev/null
tail -3 gpurun_out/pytest_r01e.log; tail -4 gpurun_out/swd_scale_r01e.log
