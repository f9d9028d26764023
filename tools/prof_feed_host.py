"""cProfile of the host side of API.feed (512 feeds of 64 images): python tools/prof_feed_host.py"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blurred_gan_b200 import sliced_wasserstein as sw

x = torch.rand(64, 3, 128, 128, device="cuda") * 2 - 1
api = sw.API((3, 128, 128))
np.random.seed(0)
api.begin("reals")
for _ in range(32):
    api.feed("reals", x)
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for _ in range(256):
    api.feed("reals", x)
pr.disable()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host %.1f us per feed (under cProfile), device drained after another %.1f ms" % ((t1 - t0) / 256 * 1e6, (t2 - t1) * 1e3))
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
t0 = time.perf_counter()
for _ in range(256):
    api.feed("reals", x)
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host %.1f us per feed (no profiler), + %.1f ms to drain" % ((t1 - t0) / 256 * 1e6, (time.perf_counter() - t1) * 1e3))
