# bench.py at N GPUs as the driver launches it, plus the distributed SWD cross-check at N = 2: bash tools/run_nN.sh N
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/bench_n${N}_r02z.log 2> gpurun_out/bench_n${N}_r02z.err; echo "rc=$?"
if [ "$N" = "2" ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/swd_dist.py 4096 64 > gpurun_out/swd_n2_r02z.log 2>&1
  grep -v "^W\|NCCL" gpurun_out/swd_n2_r02z.log | tail -4
  python -m pytest tests -m gpu -x -q -k "dist or two_rank or nccl" > gpurun_out/pytest_n2_r02z.log 2>&1; tail -n 2 gpurun_out/pytest_n2_r02z.log
fi
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n${N}_r02z.log").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], d.get("ms_per_step_per_rank"))
print("e2e", d["e2e"]["value"], d["e2e"].get("frac_of_copy_ceiling"))
s=d.get("swd",{})
print("swd", s.get("eval_s"), s.get("feed_s"), s.get("end_s"), json.dumps(s.get("sort"))[:400], json.dumps(s.get("all_gather"))[:300])
print("train", json.dumps(d.get("train_step_cfg3"))[:200])
PY
