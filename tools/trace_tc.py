"""Print the clock64 phase stamps of CTA 0 of the tensor-core blur kernel (debug aid).

Three timelines: compute thread 0, lane 0 of the MMA warp, lane 0 of the first store warp.
    python tools/trace_tc.py [std]
"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blurred_gan_b200 import _lib
lib = _lib.load()
n, c, h, w = 256, 3, 128, 128
std = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
variant, early, pipe = 0, False, True
nhwc = len(sys.argv) > 2 and sys.argv[2] == "nhwc"
xs = [torch.rand(*((n, h, w, c) if nhwc else (n, c, h, w)), device="cuda") * 2 - 1 for _ in range(4)]
ys = [torch.empty_like(x) for x in xs]
trace = torch.zeros(512, dtype=torch.int64, device="cuda")
k, r, se = _lib.blur_plan(0, std, h, w)
sp = torch.cuda.current_stream().cuda_stream
for rep in range(4):
    if rep == 3:
        os.environ["BGAN_TC_TRACE_PTR"] = str(trace.data_ptr())
    _lib.check(lib.bgan_blur2d_f32(xs[rep].data_ptr(), ys[rep].data_ptr(), n, c, h, w, 1 if nhwc else 0, 0, se, r, None, None, 0, sp), "blur")
torch.cuda.synchronize()
t = trace.cpu().tolist()
cnt = t[0]
st = t[1:1 + cnt]
base = st[0]
print("%d stamps; %.0f cycles in %.0f ns -> %.3f GHz" % (
    cnt, t[125] - st[0], t[127] - t[126], (t[125] - st[0]) / max(t[127] - t[126], 1)))
labels = ["start", "init+alloc", "taps", "chunk-table", "T-copied", "prologue", "fetch0-wait", "fetch0-landed", "first-convert"]
iters = 6
for it in range(iters):
    has_next = it + 1 < iters
    labels.append("it%d bar->G1" % it)
    if early and has_next:
        labels += ["it%d   fetch-wait" % it, "it%d   fetch-landed" % it]
    labels += ["it%d fetched(early)" % it, "it%d G1done" % it, "it%d split" % it]
    if not early and has_next:
        labels += ["it%d   fetch-wait" % it, "it%d   fetch-landed" % it]
    labels.append("it%d converted" % it)
labels.append("drain")
if pipe:
    labels = ["start", "init+alloc", "taps", "chunk-table", "T-built", "prologue", "fetch0-wait", "fetch0-landed", "conv0",
              "fetch1-wait", "fetch1-landed", "conv1"]
    for it in range(iters):
        labels += ["it%d D1 valid" % it, "it%d split" % it]
        if it + 2 < iters:
            labels += ["it%d   fetch-wait" % it, "it%d   fetch-landed" % it, "it%d   G1(i+1) done" % it]
        labels.append("it%d converted(i+2)" % it)
prev = base
for i, v in enumerate(st):
    print("%-22s %8d  (+%d)" % (labels[i] if i < len(labels) else "?", v - base, v - prev))
    prev = v
for name, b, per in (("MMA warp", 128, ["G1 issue", "G1 issued", "G2 issue", "G2 issued"]), ("store warp", 256, ["store start", "store issued"])):
    c2 = t[b]
    s2 = t[b + 1:b + 1 + c2]
    print("--", name)
    prev = base
    for i, v in enumerate(s2):
        if pipe and b == 128:  # G1(0) | G1(1) G2(0) | G1(2) G2(1) | ...
            lab = ["G1(%d) issue", "G1(%d) issued"][i % 2] % (i // 4 + (1 if i >= 2 else 0)) if (i < 2 or (i - 2) % 4 < 2) and i < 2 + 4 * (iters - 1) \
                else ["G2(%d) issue", "G2(%d) issued"][i % 2] % ((i - 2) // 4 if i < 2 + 4 * (iters - 1) else iters - 1)
            print("  %-16s %8d  (+%d)" % (lab, v - base, v - prev))
        else:
            print("  it%d %-14s %8d  (+%d)" % (i // len(per), per[i % len(per)], v - base, v - prev))
        prev = v
