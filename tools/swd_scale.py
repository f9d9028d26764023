"""Time sliced_wasserstein.API at scale on one GPU (inputs resident in HBM): python tools/swd_scale.py [n_images] [minibatch]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blurred_gan_b200 import ops
from blurred_gan_b200 import sliced_wasserstein as sw

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda")
torch.manual_seed(0)
reals = torch.rand(n, 3, 128, 128, device=dev) * 2 - 1
fakes = (torch.rand(n, 3, 128, 128, device=dev) * 2 - 1) * 0.9

def sync():
    torch.cuda.synchronize()

def run():
    np.random.seed(0)
    api = sw.API((3, 128, 128))
    t = {}
    sync(); t0 = time.perf_counter()
    api.begin("reals")
    for i in range(0, n, mb):
        api.feed("reals", reals[i:i + mb])
    sync(); t["feed_reals"] = time.perf_counter() - t0
    t1 = time.perf_counter(); api.end("reals"); sync(); t["end_reals"] = time.perf_counter() - t1
    t1 = time.perf_counter(); api.begin("fakes")
    for i in range(0, n, mb):
        api.feed("fakes", fakes[i:i + mb])
    sync(); t["feed_fakes"] = time.perf_counter() - t1
    t1 = time.perf_counter(); res = api.end("fakes"); sync(); t["end_fakes"] = time.perf_counter() - t1
    t["total"] = time.perf_counter() - t0
    return res, t

run() if n <= 4096 else None
res, t = run()
print("n=%d mb=%d" % (n, mb), {k: round(v, 4) for k, v in t.items()}, [round(float(r), 4) for r in res])
print("max mem GB", torch.cuda.max_memory_allocated() / 1e9)
# sub-phase timing of one level-0 sized problem
rows = n * 128
A = torch.randn(rows, 147, device=dev); dirs = torch.randn(147, 512, device=dev)
for name, fn in [("project", lambda: ops.swd_project(A, dirs))]:
    fn(); sync(); t0 = time.perf_counter(); P = fn(); sync(); print(name, round(time.perf_counter() - t0, 4), "s")
P2 = P.clone()
sync(); t0 = time.perf_counter(); ops.swd_sort_(P); sync(); dt = time.perf_counter() - t0
print("sort %d x %d keys: %.4f s = %.2f Gkeys/s" % (512, rows, dt, 512 * rows / dt / 1e9))
sync(); t0 = time.perf_counter(); ops.swd_absdiff_sum(P, P2); sync(); print("absdiff", round(time.perf_counter() - t0, 4))
