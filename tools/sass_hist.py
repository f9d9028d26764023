"""Summarise `ncu --page source --csv --print-source sass` output: opcode histogram + hottest lines."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
print("kernels in file:", len(blocks))
b = blocks[which]
hdr, data = b["hdr"], b["data"]
ia, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot = sum(int(r[ia]) for r in data)
tsamp = sum(int(r[isamp]) for r in data)
print(b["name"][:80], "total warp-inst", tot, "sass lines", len(data), "samples", tsamp)
op, samp = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
    o = m.group(2) if m else "?"
    o = ".".join(o.split(".")[:2]) if o.startswith(("LDS", "STS", "LDG", "STG")) else o.split(".")[0]
    op[o] += int(r[ia])
    samp[o] += int(r[isamp])
for o, c in op.most_common(22):
    print("%-12s %10d %5.1f%%  samples %5.1f%%" % (o, c, 100 * c / tot, 100 * samp[o] / max(tsamp, 1)))
