# two-GPU check: bench line at N = 2 and the distributed SWD evaluation (with the single-rank cross-check at 4096 images)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-swd --no-cpu-baseline > gpurun_out/bench_n2_r01f.log 2> gpurun_out/bench_n2_r01f.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/swd_dist.py 4096 64 > gpurun_out/swd_n2_r01f.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/swd_dist.py 16384 64 >> gpurun_out/swd_n2_r01f.log 2>&1
grep -v "^W\|NCCL" gpurun_out/swd_n2_r01f.log | tail -6; tail -c 400 gpurun_out/bench_n2_r01f.log | head -c 400; grep -o '"value": [0-9.]*' gpurun_out/bench_n2_r01f.log | head -2
