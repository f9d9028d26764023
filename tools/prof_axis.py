"""BASELINE configs[3] alone: [32, 3, 1024, 1024] fp32 NCHW, TF-rule sigma 8 / 32 / 64 through bgan_blur2d_f32 (two-pass tcgen05
axis kernel), CUDA events, two rotating buffer sets (805 MB each way per launch: nothing survives in the 126 MB L2).

    python tools/prof_axis.py [--iters 10] [--stds 8,32,64] [--check]"""
import argparse, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blurred_gan_b200 import _lib

p = argparse.ArgumentParser()
p.add_argument("--iters", type=int, default=10)
p.add_argument("--stds", default="8,32,64")
p.add_argument("--n", type=int, default=32)
p.add_argument("--check", action="store_true", help="compare one plane with the CPU oracle")
args = p.parse_args()
lib = _lib.load()
N, C, H = args.n, 3, 1024
peak = 6550.7
if os.path.exists("MEASURED_PEAKS.json"):
    peak = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
xs = [torch.rand(N, C, H, H, device="cuda") * 2 - 1 for _ in range(2)]
ys = [torch.empty_like(xs[0]) for _ in range(2)]
nbytes = 2 * xs[0].numel() * 4
for s in [float(v) for v in args.stds.split(",")]:
    _, radius, std_eff = _lib.blur_plan(_lib.RULE_TF, s, H, H)
    ws_bytes = lib.bgan_blur2d_workspace_bytes(N, C, H, H, 0, radius)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device="cuda")
    sp = torch.cuda.current_stream().cuda_stream

    def go(i):
        _lib.check(lib.bgan_blur2d_f32(xs[i & 1].data_ptr(), ys[i & 1].data_ptr(), N, C, H, H, 0, _lib.RULE_TF, std_eff, radius, None,
                                       ws.data_ptr(), ws_bytes, sp), "blur")

    go(0); go(1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.iters):
        go(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    line = "std %5.1f taps %3d path %d: %.4f ms per launch, %.1f images/s, %.1f GB/s algorithmic = %.3f of %.1f" % (
        s, 2 * radius + 1, lib.bgan_blur2d_last_path(), ms, N / (ms * 1e-3), nbytes / (ms * 1e-3) / 1e9, nbytes / (ms * 1e-3) / 1e9 / peak, peak)
    if args.check:  # first, middle and last plane against the CPU oracle
        from oracle import blur_oracle as bo
        go(0)
        errs = []
        for n, c in ((0, 0), (N // 2, 1), (N - 1, C - 1)):
            xh = np.ascontiguousarray(xs[0][n : n + 1, c : c + 1].permute(0, 2, 3, 1).cpu().numpy())
            want = bo.blur_images_tf_ref(xh, s)[0, :, :, 0]
            got = ys[0][n, c].cpu().numpy()
            errs.append(np.abs(got - want).max() / np.abs(want).max())
        line += "  err %s" % " ".join("%.2e" % e for e in errs)
    print(line, flush=True)
