"""Launch a few blur kernels at chosen stds (for ncu / timing experiments on the GPU box).

    python tools/prof_blur.py --stds 0.5,5,23 --iters 3 [--layout nchw|nhwc] [--rule torch|tf] [--shape 256,3,128,128] [--multi K]
--multi K: additionally time ONE bgan_blur2d_multi_f32 launch over K tensors against K single launches.
Prints the mean device time per launch (CUDA events, rotating buffers) when not under a profiler.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blurred_gan_b200 import _lib  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--stds", default="0.5,5,23")
p.add_argument("--iters", type=int, default=3)
p.add_argument("--layout", default="nchw")
p.add_argument("--rule", default="torch")
p.add_argument("--shape", default="256,3,128,128")
p.add_argument("--sets", type=int, default=4)
p.add_argument("--multi", type=int, default=0)
a = p.parse_args()
n, c, h, w = [int(v) for v in a.shape.split(",")]
layout = 0 if a.layout == "nchw" else 1
rule = _lib.RULE_TORCH if a.rule == "torch" else _lib.RULE_TF
lib = _lib.load()
shape = (n, c, h, w) if layout == 0 else (n, h, w, c)
xs = [torch.rand(shape, device="cuda") * 2 - 1 for _ in range(a.sets)]
ys = [torch.empty_like(x) for x in xs]
sp = torch.cuda.current_stream().cuda_stream
for s in [float(v) for v in a.stds.split(",")]:
    k, r, se = _lib.blur_plan(rule, s, h, w)
    wsb = lib.bgan_blur2d_workspace_bytes(n, c, h, w, layout, r)
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device="cuda")

    def go(i):
        _lib.check(lib.bgan_blur2d_f32(xs[i % a.sets].data_ptr(), ys[i % a.sets].data_ptr(), n, c, h, w, layout, rule, se, r, None,
                                       ws.data_ptr(), wsb, sp), "blur")

    go(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.iters):
        go(i + 1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    gb = 2 * xs[0].numel() * 4 / 1e9
    print("std=%g taps=%d path=%d  %.2f us/launch  %.0f GB/s (%.1f%% of 6550.7)  TH=%s TW=%s" % (
        s, 2 * r + 1, lib.bgan_blur2d_last_path(), ms * 1e3, gb / (ms * 1e-3), 100 * gb / (ms * 1e-3) / 6550.7,
        os.environ.get("BGAN_BLUR_TH", "-"), os.environ.get("BGAN_BLUR_TW", "-")), flush=True)
    if a.multi > 1:
        import ctypes

        k_t = a.multi
        groups = [[(i * k_t + t) % a.sets for t in range(k_t)] for i in range(a.sets)]
        arrs = [((ctypes.c_void_p * k_t)(*[xs[j].data_ptr() for j in g]), (ctypes.c_void_p * k_t)(*[ys[j].data_ptr() for j in g])) for g in groups]

        def go_multi(i):
            xp, yp = arrs[i % len(arrs)]
            _lib.check(lib.bgan_blur2d_multi_f32(xp, yp, k_t, n, c, h, w, layout, rule, se, r, None, ws.data_ptr(), wsb, sp), "multi")

        def go_singles(i):
            for j in groups[i % len(groups)]:
                go(j)

        for name, fn in (("%d single launches" % k_t, go_singles), ("one launch over %d tensors" % k_t, go_multi)):
            fn(0)
            torch.cuda.synchronize()
            e0.record()
            for i in range(a.iters):
                fn(i + 1)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.iters
            print("    std=%g %-28s %.2f us  %.0f GB/s (%.1f%% of 6550.7)" % (s, name, ms * 1e3, k_t * gb / (ms * 1e-3), 100 * k_t * gb / (ms * 1e-3) / 6550.7),
                  flush=True)
