python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --no-swd --no-cpu-baseline > gpurun_out/bench_n8_r01f.log 2> gpurun_out/bench_n8_r01f.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/swd_dist.py 16384 64 > gpurun_out/swd_n8_r01f.log 2>&1
tail -2 gpurun_out/swd_n8_r01f.log; grep -o '"value": [0-9.]*' gpurun_out/bench_n8_r01f.log | head -2
