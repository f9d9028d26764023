"""Random shapes and distributions through the segmented sort and the fused sort + |A - B| (checked against torch.sort / fp64 sums):
python tools/sort_fuzz.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blurred_gan_b200 import _lib, ops

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rs = np.random.RandomState(seed)
lib = _lib.load()
t_end = time.time() + budget
n_sorts = n_diffs = 0
paths = {1: 0, 2: 0, 3: 0}


def draw(segments, rows):
    kind = rs.randint(0, 8)
    if kind == 0:
        k = rs.randn(segments, rows)
    elif kind == 1:
        k = rs.rand(segments, rows) * 10.0 ** rs.randint(-20, 20)
    elif kind == 2:
        k = rs.randn(segments, rows) * 10.0 ** rs.randint(-3, 4, size=(segments, 1)) + rs.randn(segments, 1) * 100
    elif kind == 3:
        k = np.round(rs.randn(segments, rows) * rs.choice([1, 10, 1000]))  # many ties
    elif kind == 4:
        k = np.exp(rs.randn(segments, rows) * rs.choice([1, 4]))
    elif kind == 5:
        k = np.sort(rs.randn(segments, rows), axis=1)[:, :: rs.choice([1, -1])]
    elif kind == 6:
        k = rs.randn(segments, rows)
        k[rs.rand(segments, rows) < rs.choice([0.001, 0.05, 0.6])] = rs.choice([0.0, 1.5, -7.0])
    else:
        k = rs.standard_t(2, size=(segments, rows))  # heavy tails
    return np.ascontiguousarray(k, dtype=np.float32)


while time.time() < t_end:
    rows = int(rs.choice([rs.randint(1, 64), rs.randint(64, 9000), rs.randint(8000, 8400), rs.randint(9000, 70000), rs.randint(70000, 400000),
                          16384 * rs.randint(1, 20) + rs.randint(-2, 3)]))
    segments = int(rs.randint(1, max(2, min(40, 3_000_000 // rows))))
    a = draw(segments, rows)
    ad = torch.from_numpy(a).cuda()
    want = torch.sort(ad, dim=1).values
    got = ops.swd_sort_(ad.clone())
    paths[lib.bgan_swd_sort_last_path()] += 1
    if not torch.equal(got, want):
        bad = (got != want).nonzero()
        raise SystemExit("sort mismatch: segments %d rows %d first at %s (seed %d, case %d)" % (segments, rows, bad[0].tolist(), seed, n_sorts))
    n_sorts += 1
    if rs.rand() < 0.5:
        b = draw(segments, rows)
        per_sum = int(rs.randint(1, segments + 1))
        n_sums = (segments + per_sum - 1) // per_sum
        sums = torch.zeros(n_sums, dtype=torch.float64, device="cuda")
        ops.swd_sort_diff_(ad.clone(), torch.from_numpy(b).cuda(), sums, 0, per_sum)
        d = np.abs(np.sort(a, axis=1) - np.sort(b, axis=1)).astype(np.float64).sum(axis=1)
        ref = np.array([d[i * per_sum : (i + 1) * per_sum].sum() for i in range(n_sums)])
        if not np.allclose(sums.cpu().numpy(), ref, rtol=1e-11, atol=0):
            raise SystemExit("sort_diff mismatch: segments %d rows %d per_sum %d: %s vs %s" % (segments, rows, per_sum, sums.cpu().numpy()[:4], ref[:4]))
        n_diffs += 1
print("ok: %d sorts, %d fused sort + distance calls; routes %s" % (n_sorts, n_diffs, paths))
