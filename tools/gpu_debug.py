import sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
import oracle_lib
from pislam_b200 import api
from pislam_b200.pyramid import VGA, pyramid_layout, synthetic_stack
tma = int(sys.argv[1]) if len(sys.argv) > 1 else 0
lay = pyramid_layout(320, 240, nlevels=4)
stack = synthetic_stack(lay, 7)
o = oracle_lib.Oracle()
okp, olv, odesc = o.extract_pyramid(stack, lay.widths, lay.heights, lay.row_offsets)
ctx = api.Context(0)
ctx.set_tma(tma)
try:
    kp, lv, desc = ctx.extract_pyramid(stack, api.make_desc(lay), cap=4096)
    print("tma", tma, "n", len(kp), len(okp), "kp equal", np.array_equal(kp, okp), "desc equal", np.array_equal(desc, odesc))
    if not np.array_equal(kp, okp):
        s1, s2 = set(kp.tolist()), set(okp.tolist())
        print("missing", [hex(v) for v in sorted(s2 - s1)][:10], "extra", [hex(v) for v in sorted(s1 - s2)][:10])
except Exception as e:
    print("tma", tma, "failed:", e)
