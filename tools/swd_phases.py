"""Where the time of one SWD evaluation goes (synchronising wrappers around every op; BASELINE configs[4] by default).

    python tools/swd_phases.py [n_images] [minibatch]"""
import collections, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blurred_gan_b200 import ops, sliced_wasserstein as sw, dist as bd

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 64
torch.manual_seed(0)
reals = torch.rand(n, 3, 128, 128, device="cuda") * 2 - 1
fakes = (torch.rand(n, 3, 128, 128, device="cuda") * 2 - 1) * 0.9
acc = collections.defaultdict(float)


def timed(mod, name):
    fn = getattr(mod, name)

    def wrapper(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn(*a, **k)
        torch.cuda.synchronize()
        acc[name] += time.perf_counter() - t0
        return out

    setattr(mod, name, wrapper)


def run(sync_ops):
    np.random.seed(0)
    api = sw.API((3, 128, 128))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    phases = {}
    for mode, data in (("reals", reals), ("fakes", fakes)):
        api.begin(mode)
        t1 = time.perf_counter()
        for i in range(0, n, mb):
            api.feed(mode, data[i:i + mb])
        torch.cuda.synchronize()
        phases["feed " + mode] = time.perf_counter() - t1
        t1 = time.perf_counter()
        res = api.end(mode)
        torch.cuda.synchronize()
        phases["end " + mode] = time.perf_counter() - t1
    return time.perf_counter() - t0, phases, res


run(False)
total, phases, res = run(False)
print("eval %.4f s" % total, {k: round(v, 4) for k, v in phases.items()}, np.round(res, 4).tolist())
for name in ("swd_project", "swd_sort_", "swd_sort_diff_", "swd_absdiff_sum", "swd_normalise_", "swd_feed", "swd_extract", "pyr_down", "pyr_up"):
    timed(ops, name)
timed(bd, "positions_packed_for_round")
timed(bd, "all_gather_rows")
timed(sw, "_draw_directions")
total, phases, res = run(True)
print("with synchronising wrappers: eval %.4f s" % total, {k: round(v, 4) for k, v in phases.items()})
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
    print("  %-28s %.4f s" % (k, v))
