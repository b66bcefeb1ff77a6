"""Pins the CPU oracle (oracle/pislam_oracle.c): against the golden vectors generated from the
reference's own headers (tests/golden/make_golden.py) and, where oracle/_ref is present, against the
reference live on seeded inputs.  CPU only."""
import hashlib

import numpy as np
import pytest

import oracle_lib as ol
from pislam_b200.pyramid import VGA, noise_stack, pyramid_layout, synthetic_stack


def level_run(lib, stack, layout, border=16, ft=20, ht=1 << 15, lb=0, bl=5):
    out = np.zeros_like(stack)
    kps = []
    for w, h, r in zip(layout.widths, layout.heights, layout.row_offsets):
        lib.fast_detect(stack, out, border, w, h, ft, row0=r)
        lib.fast_score_harris(stack, out, border, w, h, ht, row0=r)
        kps.append(lib.fast_extract(out, border, w, h, lb, bl, row0=r))
    return out, kps


def test_survey_pins(golden, fixture_stack, oracle):
    """The regression pins quoted in SURVEY.md section 8c, re-derived by our own build."""
    kp, lv, desc = oracle.extract_pyramid(fixture_stack, VGA.widths, VGA.heights, VGA.row_offsets, stacked_y=True)
    assert len(kp) == 1754
    assert [int((lv == l).sum()) for l in range(8)] == [271, 296, 275, 255, 229, 184, 148, 96]
    assert ol.fnv1a64(kp) == "1ab0b92100c9c317"
    assert ol.fnv1a64(desc) == "af0ceaba8d844310"
    assert [hex(int(v)) for v in kp[:3]] == ["0x891ed010", "0x8b1f5010", "0x9821b010"]
    assert hashlib.sha256(oracle.brief_table().tobytes()).hexdigest() == \
        "15c108c2d4c62b9421bd3fde6e2f937e786cde7d2f1c3dcd228dd1dc9669ec9b"


def test_fixture_stage_by_stage(golden, fixture_stack, oracle):
    out = np.zeros_like(fixture_stack)
    for w, h, r in zip(VGA.widths, VGA.heights, VGA.row_offsets):
        oracle.fast_detect(fixture_stack, out, 16, w, h, 20, row0=r)
    nz = np.flatnonzero(out)
    assert np.array_equal(nz, golden["fixture_mask_nz"]) and (out.ravel()[nz] == 255).all()
    assert len(nz) == 18637
    kps = []
    for w, h, r in zip(VGA.widths, VGA.heights, VGA.row_offsets):
        oracle.fast_score_harris(fixture_stack, out, 16, w, h, 1 << 15, row0=r)
        kps.append(oracle.fast_extract(out, 16, w, h, row0=r))
    nz = np.flatnonzero(out)
    assert np.array_equal(nz, golden["fixture_out_nz"])
    assert np.array_equal(out.ravel()[nz], golden["fixture_out_val"])
    assert ol.fnv1a64(out) == "062f6df2346ca87a"
    assert np.array_equal(np.concatenate(kps), golden["fixture_kp"])


@pytest.mark.parametrize("lb,bl", [(4, 3), (4, 5), (3, 2), (5, 4)])
def test_fixture_bucketed(golden, fixture_stack, oracle, lb, bl):
    kp, lv, _ = oracle.extract_pyramid(fixture_stack, VGA.widths, VGA.heights, VGA.row_offsets, log_bucket=lb,
                                       bucket_limit=bl, words=8)
    assert np.array_equal(kp, golden["fixture_kp_b%d_%d" % (lb, bl)])
    assert np.array_equal(lv, golden["fixture_kp_level_b%d_%d" % (lb, bl)])
    if (lb, bl) == (4, 3):
        assert len(kp) == 1315


def test_fixture_orb(golden, fixture_stack, oracle):
    kp = golden["fixture_kp_stacked"]
    c = oracle.orb_centroids(fixture_stack, kp)
    assert np.array_equal(c, golden["fixture_centroids"])
    a = oracle.atan2(c)
    assert np.array_equal(a, golden["fixture_angles"])
    hist = np.bincount(a[:len(kp)], minlength=30).tolist()
    assert hist == [37, 46, 51, 51, 61, 48, 83, 82, 65, 47, 54, 79, 101, 119, 99, 128, 60, 39, 37, 42, 40, 42, 53,
                    38, 28, 48, 38, 31, 55, 52]
    assert np.array_equal(oracle.orb_compute(fixture_stack, kp, 8), golden["fixture_desc"])


def test_small_cases(golden, oracle):
    for i, (vstep, w, h, border, ft, ht) in enumerate(golden["small_params"].tolist()):
        img = golden["small%d_img" % i]
        out = np.zeros_like(img)
        oracle.fast_detect(img, out, border, w, h, ft)
        assert np.array_equal(out, golden["small%d_mask" % i]), i
        oracle.fast_score_harris(img, out, border, w, h, ht)
        assert np.array_equal(out, golden["small%d_out" % i]), i
        assert np.array_equal(oracle.fast_extract(out, border, w, h), golden["small%d_kp" % i]), i
        assert np.array_equal(oracle.fast_extract(out, border, w, h, 3, 2), golden["small%d_kp_b3_2" % i]), i
        assert np.array_equal(oracle.fast_extract(out, border, w, h, 4, 3), golden["small%d_kp_b4_3" % i]), i


def test_tie_heavy_extract(golden, oracle):
    for i, (vstep, w, h, border) in enumerate(golden["tie_params"].tolist()):
        m = golden["tie%d_out" % i]
        assert np.array_equal(oracle.fast_extract(m, border, w, h), golden["tie%d_kp" % i]), i
        assert np.array_equal(oracle.fast_extract(m, border, w, h, 3, 2), golden["tie%d_kp_b3_2" % i]), i
        assert np.array_equal(oracle.fast_extract(m, border, w, h, 5, 4), golden["tie%d_kp_b5_4" % i]), i


def test_orb_words_and_angles(golden, oracle):
    img, kp = golden["orb_img"], golden["orb_kp"]
    assert np.array_equal(oracle.orb_centroids(img, kp), golden["orb_centroids"])
    assert np.array_equal(oracle.atan2(golden["orb_centroids"]), golden["orb_angles"])
    for w in (1, 4, 5, 8):
        assert np.array_equal(oracle.orb_compute(img, kp, w), golden["orb_desc_w%d" % w])
    assert np.array_equal(oracle.atan2(golden["angle_xys"]), golden["angle_bins"])
    assert oracle.angle_bin(0, 0) == 7  # flat patch (SURVEY.md spec O2)


def test_brief_table_and_rot_range(golden, oracle):
    assert np.array_equal(oracle.brief_table(), golden["brief_table"])
    img = golden["orb_img"]
    for rot in (-1, 30, 255):  # Brief.h:641-732: no case -> descriptor untouched
        assert (oracle.brief_describe(img, 40, 40, rot) == 0xdeadbeef).all()


def test_util_packing(oracle):
    L = oracle.lib
    for s, x, y in ((0, 0, 0), (255, 4095, 4095), (137, 291, 16), (1, 2, 3)):
        e = L.po_encode_fast(s, x, y)
        assert e == (s << 24) | (x << 12) | y
        assert (L.po_decode_fast_score(e), L.po_decode_fast_x(e), L.po_decode_fast_y(e)) == (s, x, y)
        assert L.po_rencode_fast_score(7, e) == (7 << 24) | (e & 0xFFFFFF)


# ---- live against the reference's own headers (authoring container only) ----------------------------------
def test_live_reference_synthetic(reference, oracle):
    stack = synthetic_stack(VGA, 3)
    o_out, o_kps = level_run(oracle, stack, VGA, lb=4, bl=3)
    r_out, r_kps = level_run(reference, stack, VGA, lb=4, bl=3)
    assert np.array_equal(o_out, r_out)
    for a, b in zip(o_kps, r_kps):
        assert np.array_equal(a, b)
    offs = np.repeat(np.asarray(VGA.row_offsets, np.uint32), [len(k) for k in o_kps])
    kp = np.concatenate(o_kps)
    kp = (kp & 0xFFFFF000) | ((kp & 0xFFF) + offs)
    assert np.array_equal(oracle.orb_compute(stack, kp), reference.orb_compute(stack, kp))


def test_live_reference_noise(reference, oracle):
    lay = pyramid_layout(640, 480, nlevels=3)
    stack = noise_stack(lay, 7)
    o_out, o_kps = level_run(oracle, stack, lay, ft=40, ht=1 << 10)
    r_out, r_kps = level_run(reference, stack, lay, ft=40, ht=1 << 10)
    assert np.array_equal(o_out, r_out)
    for a, b in zip(o_kps, r_kps):
        assert np.array_equal(a, b)


def test_live_reference_harris_thresholds(reference, oracle):
    rng = np.random.Generator(np.random.PCG64(99))
    img = rng.integers(0, 256, size=(64, 128), dtype=np.uint8)
    for thr in (-(1 << 31), -1, 0, 1, 1 << 15, (1 << 31) - 1):
        for _ in range(200):
            x, y = int(rng.integers(3, 120)), int(rng.integers(3, 59))
            assert oracle.harris_score_sobel(img, x, y, thr) == reference.harris_score_sobel(img, x, y, thr)


def test_live_reference_random_ragged_levels(reference, oracle):
    """Seeded sweep over ragged level sizes (every residue of w-2B modulo 16 and 2), the borders / bucket settings
    the reference driver instantiates, random thresholds and two image statistics: the C restatement must equal
    the reference's own headers for fastDetect, fastScoreHarris, fastExtract (and orbCompute at border 16)."""
    rng = np.random.Generator(np.random.PCG64(2024))
    buckets = [(0, 5), (4, 3), (4, 5), (3, 2), (5, 4)]
    ncases = 0
    for vstep in (128, 96):
        for border in (16, 9, 4):
            for _ in range(10):
                w = int(rng.integers(2 * border + 3, vstep + 1))
                h = int(rng.integers(2 * border + 3, 90))
                rows = h + 8
                if rng.integers(0, 2):
                    img = rng.integers(0, 256, size=(rows, vstep), dtype=np.uint8)
                else:  # blocky image with mild noise: real corners, ties in the score map
                    coarse = rng.integers(0, 256, size=(rows // 6 + 2, vstep // 6 + 2), dtype=np.uint8)
                    img = np.kron(coarse, np.ones((6, 6), np.uint8))[:rows, :vstep].copy()
                    img = np.clip(img.astype(np.int32) + rng.integers(-3, 4, size=img.shape), 0, 255).astype(np.uint8)
                ft = int(rng.choice([0, 1, 7, 20, 35, 90, 127, 128, 200, 255, 300, -5]))
                ht = int(rng.choice([-(1 << 20), 0, 1 << 8, 1 << 13, 1 << 15, 1 << 22]))
                lb, bl = buckets[int(rng.integers(0, len(buckets)))]
                outs, kps = [], []
                for lib in (oracle, reference):
                    out = np.zeros_like(img)
                    lib.fast_detect(img, out, border, w, h, ft)
                    lib.fast_score_harris(img, out, border, w, h, ht)
                    outs.append(out)
                    kps.append(lib.fast_extract(out, border, w, h, lb, bl, cap=1 << 16))
                tag = "vstep %d border %d w %d h %d ft %d ht %d buckets %s" % (vstep, border, w, h, ft, ht, (lb, bl))
                assert np.array_equal(outs[0], outs[1]), tag
                assert np.array_equal(kps[0], kps[1]), tag
                if border == 16 and len(kps[0]):
                    for words in ((8, 5) if vstep == 128 else (8,)):  # the driver's orbCompute<vstep,words> menu
                        assert np.array_equal(oracle.orb_compute(img, kps[0], words),
                                              reference.orb_compute(img, kps[0], words)), tag
                ncases += 1
    assert ncases == 60


def test_pyramid_builder_matches_its_numpy_definition(oracle):
    """the reference ships no x1/1.2 resize (README.md:28-31), so the builder's definition is
    this repo's own numpy code (pislam_b200/pyramid.py) -- PARITY UNPINNED BY THE REFERENCE; the C oracle is pinned
    to that numpy code here, on every pyramid layout the tests and the bench use plus ragged ones."""
    from pislam_b200 import pyramid
    rng = np.random.Generator(np.random.PCG64(77))
    for lay, kind in ((pyramid.VGA, "synthetic"), (pyramid.pyramid_layout(320, 240, nlevels=4), "noise"),
                      (pyramid.pyramid_layout(203, 117, nlevels=6, scale=1.37), "noise"),
                      (pyramid.pyramid_layout(64, 48, nlevels=8, scale=1.2), "noise"),
                      (pyramid.FHD, "noise")):
        W, H = lay.widths[0], lay.heights[0]
        base = (pyramid.synthetic_base(W, H, 3) if kind == "synthetic"
                else rng.integers(0, 256, size=(H, W), dtype=np.uint8))
        want = pyramid.build_stack(base, lay)
        got = oracle.build_pyramid(base, lay)
        assert np.array_equal(got, want), "layout %dx%d" % (W, H)
    # single resizes incl. upscaling and 1-pixel outputs
    src = rng.integers(0, 256, size=(37, 53), dtype=np.uint8)
    for w, h in ((53, 37), (44, 31), (1, 1), (2, 37), (53, 1), (80, 50), (7, 5)):
        assert np.array_equal(oracle.resize_bilinear(src, w, h), pyramid._resize_bilinear(src, w, h)), (w, h)
    # with the blur in front: level 0 is the Gaussian.h:48-72 blur of the base, the others are resized from it
    lay = pyramid.pyramid_layout(160, 120, nlevels=3)
    base = rng.integers(0, 256, size=(120, 160), dtype=np.uint8)
    blurred = np.zeros_like(base)
    oracle.gaussian5x5(base, blurred, 160, 120)
    assert np.array_equal(oracle.build_pyramid(base, lay, blur=True), pyramid.build_stack(blurred, lay))


def test_reference_batch_driver_equals_the_oracle(oracle, reference):
    """oracle/ref_driver.cpp:ref_extract_batch (the reference's caller loop run natively over a batch, bench.py's
    CPU arm) returns, frame by frame, what the oracle's po_extract_pyramid returns."""
    from pislam_b200 import pyramid
    stacks = np.stack([pyramid.synthetic_stack(pyramid.VGA, 700 + i) for i in range(5)])
    lay = pyramid.VGA
    for lb, bl in ((4, 3), (0, 5)):
        for nthreads in (1, 3):
            counts, kp, desc = reference.extract_batch(stacks, lay.widths, lay.heights, lay.row_offsets, nthreads,
                                                       log_bucket=lb, bucket_limit=bl, cap=4096, keep=True)
            for f in range(len(stacks)):
                okp, _, odesc = oracle.extract_pyramid(stacks[f], lay.widths, lay.heights, lay.row_offsets,
                                                       log_bucket=lb, bucket_limit=bl)
                n = int(counts[f])
                assert n == len(okp)
                assert np.array_equal(kp[f, :n], okp)
                assert np.array_equal(desc[f, :n * 8].reshape(n, 8), odesc)
