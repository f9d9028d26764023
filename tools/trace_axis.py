"""clock64 timelines of CTA 0 of blur_axis_sweep_kernel (debug aid): compute thread 0, the MMA-issuing lane, lane 0 of the
first store warp.      python tools/trace_axis.py [std] [n_images]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blurred_gan_b200 import _lib
lib = _lib.load()
std = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
N = int(sys.argv[2]) if len(sys.argv) > 2 else 32
C, H = 3, 1024
x = torch.rand(N, C, H, H, device="cuda") * 2 - 1
y = torch.empty_like(x)
_, radius, std_eff = _lib.blur_plan(_lib.RULE_TF, std, H, H)
ws_bytes = lib.bgan_blur2d_workspace_bytes(N, C, H, H, 0, radius)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
trace = torch.zeros(768, dtype=torch.int64, device="cuda")
sp = torch.cuda.current_stream().cuda_stream
for rep in range(3):
    if rep == 2:
        os.environ["BGAN_AXIS_TRACE_PTR"] = str(trace.data_ptr())
    _lib.check(lib.bgan_blur2d_f32(x.data_ptr(), y.data_ptr(), N, C, H, H, 0, _lib.RULE_TF, std_eff, radius, None, ws.data_ptr(), ws_bytes, sp), "blur")
torch.cuda.synchronize()
t = trace.cpu().tolist()
# (both passes write the same buffer: what is left is the second pass)
names = {0: ["wait", "landed", "set-free", "converted"], 1: ["ready", "issued"], 2: ["full", "stored"]}
t0 = min(t[r * 256 + 1] for r in range(3) if t[r * 256] > 0)
for role, label in ((0, "compute thread 0"), (1, "MMA lane"), (2, "store warp 0")):
    n = t[role * 256]
    st = t[role * 256 + 1: role * 256 + 1 + n]
    per = len(names[role])
    print("%s: %d stamps" % (label, n))
    for i in range(0, min(n, per * 20), per):
        row = st[i:i + per]
        print("  unit %2d: " % (i // per) + "  ".join("%s %7d" % (nm, v - t0) for nm, v in zip(names[role], row)) +
              ("   (period %d)" % (row[0] - st[i - per]) if i >= per else ""))
