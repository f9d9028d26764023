"""Time the segmented sort (CUDA events): python tools/sort_bench.py [segments] [rows]   (BGAN_SORT_ALGO=lsd: radix passes only)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blurred_gan_b200 import ops
seg = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
torch.manual_seed(0)
src = torch.randn(seg, rows, device="cuda")
keys = src.clone()
ops.swd_sort_(keys)
ref = torch.sort(src, dim=1).values
print("exact:", bool(torch.equal(keys, ref)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(5):
    keys.copy_(src)
    e0.record(); ops.swd_sort_(keys); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = min(ts)
print("%d x %d keys: %.3f ms  %.1f Gkeys/s" % (seg, rows, ms, seg * rows / ms / 1e6))
