"""Multi-GPU SWD evaluation (BASELINE configs[4]): torchrun --nproc-per-node N tools/swd_dist.py [n_images] [minibatch]

Every rank builds the same synthetic image sets (same seed) and feeds the minibatches it owns (global minibatch index
g = k * world + rank); descriptors are all-gathered over NCCL; every rank sorts its block of the 512 directions.
Prints the evaluation time (max over ranks) and, for n <= 4096, checks the result against a single-rank evaluation."""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blurred_gan_b200 import sliced_wasserstein as sw

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 64
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
reals = torch.rand(n, 3, 128, 128, device=dev) * 2 - 1
fakes = (torch.rand(n, 3, 128, 128, device=dev) * 2 - 1) * 0.9


def evaluate(distributed):
    r, w = (rank, world) if distributed else (0, 1)
    np.random.seed(0)
    api = sw.API((3, 128, 128), distributed=distributed)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for mode, data in (("reals", reals), ("fakes", fakes)):
        api.begin(mode)
        for k in range(n // (mb * w)):
            g = k * w + r
            api.feed(mode, data[g * mb:(g + 1) * mb])
        res = api.end(mode)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return [float(v) for v in res], float(dt.item())


evaluate(True)  # warm-up (allocator, NCCL channels)
res, dt = evaluate(True)
if rank == 0:
    print("world=%d n=%d mb=%d  eval %.4f s  SWDx1e3 %s" % (world, n, mb, dt, np.round(res, 4).tolist()), flush=True)
if n <= 4096:
    single, _ = evaluate(False)
    if rank == 0:
        rel = np.max(np.abs(np.array(res) - np.array(single)) / np.array(single))
        print("single-rank %s  max rel diff %.2e" % (np.round(single, 4).tolist(), rel), flush=True)
        assert rel < 1e-4
if world > 1:
    dist.destroy_process_group()
