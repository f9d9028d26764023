# Round-1 (fourth session), second GPU call: cta_group::2 dispatch cost, smoke(), a short bench run.
set -x
timeout 120 tools/mma_rate.bin 2 > gpurun_out/mma_rate2_r01f.log 2>&1; echo "rc=$?" >> gpurun_out/mma_rate2_r01f.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r01f.log 2>&1; echo "rc=$?" >> gpurun_out/smoke_r01f.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --swd-images 2048 > gpurun_out/bench_quick_r01f.log 2> gpurun_out/bench_quick_r01f.err; echo "rc=$?" >> gpurun_out/bench_quick_r01f.err
cat gpurun_out/mma_rate2_r01f.log; tail -3 gpurun_out/smoke_r01f.log; tail -c 1500 gpurun_out/bench_quick_r01f.log; tail -3 gpurun_out/bench_quick_r01f.err
