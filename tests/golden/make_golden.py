"""Generates tests/golden/*.npz from the REFERENCE's own headers (oracle/_ref/libpislam_ref.so, built by
oracle/Makefile from /root/reference through the NEON shim).  Run in the authoring container only:

    make -C oracle ref && python tests/golden/make_golden.py

The reference ships no known-answer vectors for this path, so these are outputs of
the reference itself: demo parameters (border 16, FAST 20, Harris 1<<15, words 8, `out` zeroed) on its
only fixture demo/input.png, plus seeded small cases that exercise ragged widths, ties and bucketing.
"""
import hashlib
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as ol  # noqa: E402
from pislam_b200.pyramid import VGA  # noqa: E402

REF_PNG = "/root/reference/demo/input.png"


def run_levels(ref, stack, layout, border, fast_thr, harris_thr, log_bucket, bucket_limit):
    out = np.zeros_like(stack)
    kps, lvl = [], []
    for l, (w, h, r) in enumerate(zip(layout.widths, layout.heights, layout.row_offsets)):
        ref.fast_detect(stack, out, border, w, h, fast_thr, row0=r)
    mask = out.copy()
    for l, (w, h, r) in enumerate(zip(layout.widths, layout.heights, layout.row_offsets)):
        ref.fast_score_harris(stack, out, border, w, h, harris_thr, row0=r)
        kp = ref.fast_extract(out, border, w, h, log_bucket, bucket_limit, row0=r)
        kps.append(kp)
        lvl.append(np.full(len(kp), l, np.uint8))
    return mask, out, np.concatenate(kps), np.concatenate(lvl)


def main():
    ref = ol.Reference()
    g = {}
    stack = cv2.imread(REF_PNG, cv2.IMREAD_UNCHANGED)
    assert stack.shape == (2210, 640) and stack.dtype == np.uint8
    ok, png = cv2.imencode(".png", stack, [cv2.IMWRITE_PNG_COMPRESSION, 9])
    g["fixture_png"] = png.ravel()
    mask, out, kp, lv = run_levels(ref, stack, VGA, 16, 20, 1 << 15, 0, 5)
    g["fixture_mask_nz"] = np.flatnonzero(mask).astype(np.uint32)  # all 0xff
    nz = np.flatnonzero(out).astype(np.uint32)
    g["fixture_out_nz"] = nz
    g["fixture_out_val"] = out.ravel()[nz]
    g["fixture_kp"] = kp  # level-local y
    g["fixture_kp_level"] = lv
    offs = np.asarray(VGA.row_offsets, np.uint32)[lv]
    kp_st = (kp & 0xFFFFF000) | ((kp & 0xFFF) + offs)  # demo.cpp:91-97
    g["fixture_kp_stacked"] = kp_st
    g["fixture_centroids"] = ref.orb_centroids(stack, kp_st)
    g["fixture_angles"] = ref.atan2(g["fixture_centroids"])
    g["fixture_desc"] = ref.orb_compute(stack, kp_st, 8)
    for lb, bl in ((4, 3), (4, 5), (3, 2), (5, 4)):
        _, _, kpb, lvb = run_levels(ref, stack, VGA, 16, 20, 1 << 15, lb, bl)
        g["fixture_kp_b%d_%d" % (lb, bl)] = kpb
        g["fixture_kp_level_b%d_%d" % (lb, bl)] = lvb
    # rBRIEF table, probed pixel by pixel out of the unmodified Brief.h
    table = np.zeros((30, 256, 4), np.int8)
    img = np.zeros((64, 128), np.uint8)
    cx, cy = 40, 32
    # I[p0] < I[p1]: light exactly one pixel q; bits that become 1 have p1 == q. Then invert the image to get p0.
    for rot in range(30):
        p1 = {}
        p0 = {}
        for dy in range(-15, 16):
            for dx in range(-15, 16):
                img[:] = 0
                img[cy + dy, cx + dx] = 255
                d = ref.brief_describe(img, cx, cy, rot)
                img[:] = 255
                img[cy + dy, cx + dx] = 0
                e = ref.brief_describe(img, cx, cy, rot)
                for w in range(8):
                    for b in range(32):
                        if (int(d[w]) >> b) & 1:
                            p1.setdefault(32 * w + b, []).append((dx, dy))
                        if (int(e[w]) >> b) & 1:
                            p0.setdefault(32 * w + b, []).append((dx, dy))
        for k in range(256):
            # a pair whose two clamped points coincide never fires: recorded as (127,127,127,127)
            if k in p1 and k in p0:
                assert len(p1[k]) == 1 and len(p0[k]) == 1, (rot, k)
                table[rot, k] = (p0[k][0][0], p0[k][0][1], p1[k][0][0], p1[k][0][1])
            else:
                table[rot, k] = 127
    g["brief_table"] = table
    print("brief table sha256", hashlib.sha256(table.tobytes()).hexdigest())

    # small seeded cases: noise images, ragged widths, several borders -> full stage outputs
    rng = np.random.Generator(np.random.PCG64(1234))
    cases = []
    for (vstep, w, h, border, ft, ht) in ((128, 96, 72, 16, 20, 1 << 15), (128, 77, 61, 16, 12, 1000),
                                          (128, 83, 50, 9, 30, 1 << 12), (96, 64, 64, 4, 25, 1 << 15),
                                          (128, 125, 40, 4, 40, 0), (96, 96, 48, 9, 200, -5)):
        img = rng.integers(0, 256, size=(h + 8, vstep), dtype=np.uint8)
        # smooth half of the image a little so that Harris codes vary
        img[:, : vstep // 2] = cv2.blur(img[:, : vstep // 2], (3, 3))
        out = np.zeros_like(img)
        ref.fast_detect(img, out, border, w, h, ft)
        mask = out.copy()
        ref.fast_score_harris(img, out, border, w, h, ht)
        kp0 = ref.fast_extract(out, border, w, h, 0, 5)
        kp1 = ref.fast_extract(out, border, w, h, 3, 2)
        kp2 = ref.fast_extract(out, border, w, h, 4, 3)
        i = len(cases)
        g["small%d_img" % i] = img
        g["small%d_mask" % i] = mask
        g["small%d_out" % i] = out
        g["small%d_kp" % i] = kp0
        g["small%d_kp_b3_2" % i] = kp1
        g["small%d_kp_b4_3" % i] = kp2
        cases.append((vstep, w, h, border, ft, ht))
    g["small_params"] = np.asarray(cases, np.int64)
    # tie-heavy score maps for fastExtract alone (values 0..3, 25% density)
    ties = []
    for i, (vstep, w, h, border) in enumerate(((128, 96, 72, 16), (128, 77, 61, 16), (96, 64, 64, 9),
                                               (128, 83, 50, 4))):
        m = rng.integers(0, 4, size=(h + 4, vstep), dtype=np.uint8) * (rng.random((h + 4, vstep)) < 0.25)
        m = m.astype(np.uint8)
        g["tie%d_out" % i] = m
        g["tie%d_kp" % i] = ref.fast_extract(m, border, w, h, 0, 5)
        g["tie%d_kp_b3_2" % i] = ref.fast_extract(m, border, w, h, 3, 2)
        g["tie%d_kp_b5_4" % i] = ref.fast_extract(m, border, w, h, 5, 4)
        ties.append((vstep, w, h, border))
    g["tie_params"] = np.asarray(ties, np.int64)
    # orb on a small noise image, several word counts (the driver instantiates vstep 128: words 1,4,5,8)
    img = rng.integers(0, 256, size=(80, 128), dtype=np.uint8)
    pts = np.asarray([(s << 24) | (x << 12) | y for s, x, y in
                      [(200, 16 + int(rng.integers(0, 96)), 16 + int(rng.integers(0, 48))) for _ in range(37)]],
                     np.uint32)
    g["orb_img"] = img
    g["orb_kp"] = pts
    for wds in (1, 4, 5, 8):
        g["orb_desc_w%d" % wds] = ref.orb_compute(img, pts, wds)
    g["orb_centroids"] = ref.orb_centroids(img, pts)
    g["orb_angles"] = ref.atan2(g["orb_centroids"])
    # angle bins on adversarial moment pairs
    m = rng.integers(-60000, 60000, size=(4096, 2)).astype(np.int32)
    m[:64, 0] = 0
    m[32:96, 1] = 0
    m[96:160, 1] = m[96:160, 0]
    m[160:224, 1] = -m[160:224, 0]
    xy = np.zeros(4096 * 2, np.int32).reshape(-1, 8)
    xy[:, :4] = m[:, 0].reshape(-1, 4)
    xy[:, 4:] = m[:, 1].reshape(-1, 4)
    g["angle_xys"] = xy.ravel()
    g["angle_bins"] = ref.atan2(xy.ravel())
    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **g)
    print("fixture kp", len(kp), ol.fnv1a64(kp_st), "desc", ol.fnv1a64(g["fixture_desc"]))
    print("wrote", os.path.join(HERE, "reference_golden.npz"), os.path.getsize(os.path.join(HERE, "reference_golden.npz")))


if __name__ == "__main__":
    main()
