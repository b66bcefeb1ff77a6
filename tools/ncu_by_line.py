#!/usr/bin/env python3
"""Aggregate an ncu SASS source page by CUDA source line.

    python tools/ncu_by_line.py <report.ncu-rep> <kernel-substring> [top]

ncu's CSV source page is per SASS instruction; the line table comes from `nvdisasm -g` on the cubin
embedded in pislam_b200/lib/libpislam_b200.so (must be the build that was profiled).  The substring must select ONE
kernel instantiation of the cubin, e.g. front_tile_kernelILi0ELi64 (MODE_TMA, 128x64 tiles) or orb_tma_kernel.
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_table(kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "pislam_b200/lib/libpislam_b200.so")],
                          cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, text=True).stdout
    table, cur, active = {}, None, False
    for ln in dis.splitlines():
        if ln.startswith("\t.section\t.text."):
            active = kernel_sub in ln
            cur = None
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + ksub.split("ILi")[0][-24:]],
                         stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ia, iinst, ithr, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), \
        hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    iconf = hdr.index("L1 Wavefronts Shared Excessive") if "L1 Wavefronts Shared Excessive" in hdr else None  # no smem: absent
    table = line_table(ksub)
    base = None
    agg = {}
    tot = [0, 0, 0]
    for r in rows[hi + 1:]:
        if len(r) <= ithr or not r[ia]:
            continue
        try:
            addr = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
        except ValueError:
            continue
        if base is None:
            base = addr
        key, sass = table.get(addr - base, (None, "?"))
        a = agg.setdefault(key, [0, 0, 0, 0])
        vals = [int(float(r[iinst] or 0)), int(float(r[ithr] or 0)), int(float(r[isamp] or 0)), int(float(r[iconf] or 0)) if iconf is not None else 0]
        for k in range(4):
            a[k] += vals[k]
        for k in range(3):
            tot[k] += vals[k]
    src_cache = {}
    print("total warp-inst %d  thread-inst %d  samples %d" % tuple(tot))
    print("%-28s %12s %6s %14s %8s %8s  %s" % ("file:line", "warp-inst", "%", "thread-inst", "samples", "smem-xs", "source"))
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if key:
            f = src_cache.setdefault(key[0], open(os.path.join(ROOT, "pislam_b200/csrc", key[0])).read().splitlines()
                                     if os.path.exists(os.path.join(ROOT, "pislam_b200/csrc", key[0])) else [])
            text = f[key[1] - 1].strip()[:70] if 0 < key[1] <= len(f) else ""
        print("%-28s %12d %6.2f %14d %8d %8d  %s" % ("%s:%d" % key if key else "?", a[0], 100.0 * a[0] / max(tot[0], 1),
                                                     a[1], a[2], a[3], text))


if __name__ == "__main__":
    main()
