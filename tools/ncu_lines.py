"""Source lines of an `ncu --page source --csv --print-source cuda,sass` dump ranked by executed instructions:
python tools/ncu_lines.py dump.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, out = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1]
        continue
    if len(r) > 2 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)
        try:
            ie = int(d["Instructions Executed"])
        except ValueError:
            continue
        if ie > 0:
            out.append((ie, cur.split("/")[-1], int(r[0]), r[1][:100], d.get("L1 Wavefronts Shared", ""), d.get("# Samples", "")))
tot = sum(o[0] for o in out)
print("total warp instructions", tot)
for o in sorted(out, reverse=True)[:top]:
    print("%5.1f%% %11d %s:%d  smem wavefronts=%s samples=%s | %s" % (100 * o[0] / tot, o[0], o[1], o[2], o[4], o[5], o[3]))
