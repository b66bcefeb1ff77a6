#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the shipped library (cuobjdump -sass): the evidence that the fused kernels are
sm_100a code that stages by TMA (UTMALDG), computes on packed bytes (VABSDIFF4, IDP.4A, PRMT), elects single lanes
(ELECT) and synchronises on mbarriers (SYNCS) -- and that there is no tensor-core instruction, as expected for an
8-bit compare/reduce path.   python tools/sass_extract.py > profiles/r02_sass_extract.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pislam_b200", "lib", "libpislam_b200.so")
WATCH = ["UTMALDG", "UTMAPF", "SYNCS", "ELECT", "VABSDIFF4", "IDP.4A", "PRMT", "IMAD", "LOP3", "SHF", "POPC", "REDUX",
         "CREDUX", "ATOMS", "ATOMG", "LDS", "STS", "LDG", "STG", "BAR.SYNC", "HMMA", "IMMA", "UTCHMMA", "UTCIMMA", "TCGEN"]


def main():
    elf = subprocess.run(["cuobjdump", "-lelf", LIB], stdout=subprocess.PIPE, text=True).stdout.strip()
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
    print("library:", os.path.relpath(LIB, ROOT))
    print("cubins :", elf.replace("\n", "; "))
    kernels = collections.OrderedDict()
    cur = None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and cur:
            op = m.group(1)
            kernels[cur]["total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + ".") or (w == "IDP.4A" and op.startswith("IDP.4A")):
                    kernels[cur][w] += 1
    cols = ["total"] + WATCH
    print("%-62s" % "kernel" + "".join("%10s" % c for c in cols))
    for k, c in kernels.items():
        print("%-62s" % k[:61] + "".join("%10d" % c[x] for x in cols))
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print("%-62s" % "ALL KERNELS" + "".join("%10d" % tot[x] for x in cols))
    assert tot["UTMALDG"] > 0 and tot["VABSDIFF4"] > 0 and tot["IDP.4A"] > 0
    assert tot["HMMA"] + tot["IMMA"] + tot["UTCHMMA"] + tot["UTCIMMA"] + tot["TCGEN"] == 0


if __name__ == "__main__":
    main()
