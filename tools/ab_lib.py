"""A/B of two builds of the library on the same box: python tools/ab_lib.py tools/ab_old.so [stds]   (interleaved timing of
bgan_blur2d_f32 on BASELINE configs[1]; the ABI of that entry point is the same in both)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blurred_gan_b200 import _lib

new = _lib.load()
old = ctypes.CDLL(os.path.abspath(sys.argv[1]))
for name in ("bgan_blur2d_f32", "bgan_blur_plan"):
    getattr(old, name).restype, getattr(old, name).argtypes = _lib.SIGNATURES[name]
stds = [float(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "1,5,23").split(",")]
n, c, h, w = 256, 3, 128, 128
xs = [torch.rand(n, c, h, w, device="cuda") * 2 - 1 for _ in range(6)]
ys = [torch.empty_like(x) for x in xs]
sp = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for s in stds:
    _, r, se = _lib.blur_plan(_lib.RULE_TORCH, s, h, w)
    res = {"old": [], "new": []}
    for rep in range(5):
        for name, lib in (("old", old), ("new", new)):
            for i in range(3):
                lib.bgan_blur2d_f32(xs[i].data_ptr(), ys[i].data_ptr(), n, c, h, w, 0, 0, se, r, None, None, 0, sp)
            e0.record()
            for i in range(60):
                lib.bgan_blur2d_f32(xs[i % 6].data_ptr(), ys[i % 6].data_ptr(), n, c, h, w, 0, 0, se, r, None, None, 0, sp)
            e1.record()
            torch.cuda.synchronize()
            res[name].append(e0.elapsed_time(e1) / 60 * 1e3)
    print("std=%g  old %s  new %s  (us per launch, 5 interleaved repetitions)" % (s, [round(v, 2) for v in res["old"]], [round(v, 2) for v in res["new"]]), flush=True)
