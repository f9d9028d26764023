"""Print the clock64 phase stamps of CTA 0 of the tensor-core blur kernel (debug aid)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blurred_gan_b200 import _lib
lib = _lib.load()
n, c, h, w = 256, 3, 128, 128
x = torch.rand(n, c, h, w, device="cuda") * 2 - 1
y = torch.empty_like(x)
trace = torch.zeros(128, dtype=torch.int64, device="cuda")
k, r, se = _lib.blur_plan(0, 5.0, h, w)
sp = torch.cuda.current_stream().cuda_stream
for rep in range(3):
    if rep == 2:
        os.environ["BGAN_TC_TRACE_PTR"] = str(trace.data_ptr())
    _lib.check(lib.bgan_blur2d_f32(x.data_ptr(), y.data_ptr(), n, c, h, w, 0, 0, se, r, None, None, 0, sp), "blur")
    torch.cuda.synchronize()
t = trace.cpu().tolist()
cnt = t[0]
st = t[1:1 + cnt]
print("stamps", cnt)
base = st[0]
names = ["start", "init+alloc", "taps", "chunk-table", "T-copied", "prologue", "first-convert"]
per = ["barX", "G1done", "split+barZ", "fetch-next", "convert-next"]
i = 0
prev = base
print("clock: %.0f cycles in %.0f ns -> %.3f GHz" % (t[125] - st[0], t[127] - t[126], (t[125] - st[0]) / max(t[127] - t[126], 1)))
for k in range(7):
    print("%-12s %8d  (+%d)" % (names[k], st[k] - base, st[k] - prev)); prev = st[k]
i = 7
it = 0
NP = len(per)
while cnt - i >= NP + 1:
    print("-- iteration", it)
    for k in range(NP):
        print("  %-10s %8d  (+%d)" % (per[k], st[i + k] - base, st[i + k] - prev)); prev = st[i + k]
    i += NP; it += 1
for k in range(i, cnt):
    print("%-12s %8d  (+%d)" % ("drain", st[k] - base, st[k] - prev)); prev = st[k]
