import cProfile, pstats, sys, os, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from blurred_gan_b200 import sliced_wasserstein as sw
n, mb = 4096, 64
x = torch.rand(n, 3, 128, 128, device="cuda") * 2 - 1
def feeds():
    np.random.seed(0)
    api = sw.API((3, 128, 128))
    api.begin("reals")
    for i in range(0, n, mb):
        api.feed("reals", x[i:i + mb])
    torch.cuda.synchronize()
feeds()
t0 = time.perf_counter(); feeds(); print("feed wall", time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable(); feeds(); pr.disable()
pstats.Stats(pr).sort_stats("cumtime").print_stats(22)
