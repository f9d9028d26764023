"""Blackwell-native evidence without a GPU: per kernel of libbgan_b200.so, how many SASS instructions of the mnemonics that only
tcgen05 / TMEM / TMA code produces (B200_PROFILING.md, "What proves a Blackwell-native kernel").

    python tools/sass_summary.py > profiles/r02_sass_hist.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "blurred_gan_b200", "libbgan_b200.so")
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UTCATOMSWS", "SYNCS", "LDGSTS", "HMMA", "VOTE", "MATCH",
         "ATOMS", "REDG", "ATOMG", "DFMA", "DADD", "FFMA2", "FFMA"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
kernels = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"bgan::\(anonymous namespace\)::", "", name)
        name = re.sub(r"\((.*)\)$", "", name)  # drop the parameter list
        cur = kernels.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur is not None:
        cur["_total"] += 1
        op = m.group(1)
        for w in WATCH:
            if op.startswith(w):
                cur[w] += 1
                break
print("%s: %d kernels, sm_100a SASS (cuobjdump -sass); columns = instructions whose mnemonic starts with ..." % (os.path.basename(LIB), len(kernels)))
cols = [w for w in WATCH if any(k[w] for k in kernels.values())]
print("%-74s %7s " % ("kernel", "total") + " ".join("%7s" % c[:7] for c in cols))
tot = collections.Counter()
for name, c in kernels.items():
    print("%-74s %7d " % (name[:74], c["_total"]) + " ".join("%7s" % (c[w] or "") for w in cols))
    tot.update(c)
print("%-74s %7d " % ("ALL", tot["_total"]) + " ".join("%7d" % tot[w] for w in cols))
