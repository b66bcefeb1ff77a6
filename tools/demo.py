#!/usr/bin/env python3
"""The reference demo program (reference demo/demo.cpp:51-130) on the B200 path: read a vertically stacked
VGA 8-level pyramid PNG, extract ORB features through pislam_b200 (the fused whole-pyramid call), paint the
reference's cross markers at every keypoint (stack coordinates), write out.png, print time and feature count.

    python tools/demo.py input.png [out.png] [--buckets]      # needs a B200

With --stages it runs the literal per-level calls in the order of the reference's loop instead.
"""
import sys
import time

import cv2
import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from pislam_b200 import api  # noqa: E402
from pislam_b200.pyramid import VGA  # noqa: E402


def paint_point(img, x, y):
    """demo.cpp:119-130: an 8-pixel cross, arms at distance 4 and 5"""
    for d in (-5, -4, 4, 5):
        img[y + d, x] = 0
        img[y, x + d] = 0


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    flags = {a for a in sys.argv[1:] if a.startswith("--")}
    if not args:
        raise SystemExit(__doc__)
    img = cv2.imread(args[0], cv2.IMREAD_GRAYSCALE)
    if img is None or img.shape != (VGA.rows, VGA.vstep):
        raise SystemExit("expected a %dx%d grayscale PNG (the stacked pyramid of demo/input.png)" % (VGA.vstep, VGA.rows))
    img = np.ascontiguousarray(img)
    lb, bl = (4, 3) if "--buckets" in flags else (0, 5)
    with api.Context(0) as ctx:
        ctx.extract_pyramid(img, api.make_desc(VGA, log_bucket=lb, bucket_limit=bl), cap=8192)  # warm-up
        t0 = time.perf_counter()
        if "--stages" in flags:
            out = np.zeros_like(img)
            pts = []
            for w, h, r in zip(VGA.widths, VGA.heights, VGA.row_offsets):
                ctx.fastDetect(w, h, img, out, 20, row0=r)
                ctx.fastScoreHarris(w, h, img, 1 << 15, out, row0=r)
                kp = ctx.fastExtract(w, h, out, row0=r, logBucketSize=lb, bucketLimit=bl)
                pts.append((kp & 0xFFFFF000) | ((kp & 0xFFF) + np.uint32(r)))  # demo.cpp:91-97
            kp = np.concatenate(pts)
            desc = ctx.orbCompute(img, kp)
        else:
            kp, lv, desc = ctx.extract_pyramid(img, api.make_desc(VGA, log_bucket=lb, bucket_limit=bl), cap=8192)
            kp = (kp & 0xFFFFF000) | ((kp & 0xFFF) + np.asarray(VGA.row_offsets, np.uint32)[lv])
        ms = (time.perf_counter() - t0) * 1e3
    for p in kp:
        paint_point(img, int((p >> 12) & 0xFFF), int(p & 0xFFF))
    cv2.imwrite(args[1] if len(args) > 1 else "out.png", img)
    print("GPU  Time: %.3f ms" % ms)
    print("%d features" % len(kp))


if __name__ == "__main__":
    main()
