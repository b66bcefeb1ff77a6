"""Parity of the CUDA path (through the C ABI, via pislam_b200.api) against the CPU oracle and the
golden vectors generated from the reference's own headers.  Needs a B200: `pytest -m gpu`.

Bar: bit-exact everywhere (masks, score bytes, ordered keypoint lists, centroids, angle bins, descriptors)."""
import numpy as np
import pytest

from pislam_b200 import api
from pislam_b200.pyramid import FHD, UHD, VGA, PyramidLayout, noise_stack, pyramid_layout, synthetic_stack

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def first_diff(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        return "shape %s vs %s" % (a.shape, b.shape)
    d = np.flatnonzero(a.ravel() != b.ravel())
    if len(d) == 0:
        return "equal"
    i = int(d[0])
    return "%d diffs, first at %d: got %s want %s" % (len(d), i, a.ravel()[i], b.ravel()[i])


def assert_same(got, want, what):
    assert np.array_equal(got, want), "%s: %s" % (what, first_diff(got, want))


# ------------------------------------------------------------------------------------------ literal stages
def test_literal_stages_small_golden(ctx, golden):
    """fastDetect / fastScoreHarris / fastExtract one by one, against reference outputs: ragged widths,
    borders 16/9/4, thresholds incl. > 127 and negative Harris thresholds."""
    for i, (vstep, w, h, border, ft, ht) in enumerate(golden["small_params"].tolist()):
        img = np.ascontiguousarray(golden["small%d_img" % i])
        out = np.zeros_like(img)
        ctx.fastDetect(w, h, img, out, ft, border=border)
        assert_same(out, golden["small%d_mask" % i], "fastDetect case %d" % i)
        ctx.fastScoreHarris(w, h, img, ht, out, border=border)
        assert_same(out, golden["small%d_out" % i], "fastScoreHarris case %d" % i)
        assert_same(ctx.fastExtract(w, h, out, border=border), golden["small%d_kp" % i], "fastExtract case %d" % i)
        assert_same(ctx.fastExtract(w, h, out, border=border, logBucketSize=3, bucketLimit=2),
                    golden["small%d_kp_b3_2" % i], "fastExtract<3,2> case %d" % i)
        assert_same(ctx.fastExtract(w, h, out, border=border, logBucketSize=4, bucketLimit=3),
                    golden["small%d_kp_b4_3" % i], "fastExtract<4,3> case %d" % i)


def test_literal_stages_random_ragged(ctx, oracle):
    """Seeded sweep over ragged level sizes, borders 16/9/4, thresholds (incl. 0, >= 128, > 255, negative), bucket
    settings and two image statistics -- the sweep tests/test_oracle.py runs oracle-vs-reference, here CUDA-vs-oracle."""
    rng = np.random.Generator(np.random.PCG64(2025))
    for vstep in (128, 96, 200):
        for border in (16, 9, 4):
            for _ in range(5):
                w = int(rng.integers(2 * border + 3, vstep + 1))
                h = int(rng.integers(2 * border + 3, 90))
                rows = h + 8
                if rng.integers(0, 2):
                    img = rng.integers(0, 256, size=(rows, vstep), dtype=np.uint8)
                else:
                    coarse = rng.integers(0, 256, size=(rows // 6 + 2, vstep // 6 + 2), dtype=np.uint8)
                    img = np.kron(coarse, np.ones((6, 6), np.uint8))[:rows, :vstep].copy()
                    img = np.clip(img.astype(np.int32) + rng.integers(-3, 4, size=img.shape), 0, 255).astype(np.uint8)
                ft = int(rng.choice([0, 1, 7, 20, 35, 90, 127, 128, 200, 255, 300, -5]))
                ht = int(rng.choice([-(1 << 20), 0, 1 << 8, 1 << 13, 1 << 15, 1 << 22]))
                lb, bl = int(rng.integers(0, 6)), int(rng.integers(1, 7))
                tag = "vstep %d border %d w %d h %d ft %d ht %d buckets %s" % (vstep, border, w, h, ft, ht, (lb, bl))
                want = np.zeros_like(img)
                oracle.fast_detect(img, want, border, w, h, ft)
                out = np.zeros_like(img)
                ctx.fastDetect(w, h, img, out, ft, border=border)
                assert_same(out, want, "fastDetect " + tag)
                oracle.fast_score_harris(img, want, border, w, h, ht)
                ctx.fastScoreHarris(w, h, img, ht, out, border=border)
                assert_same(out, want, "fastScoreHarris " + tag)
                okp = oracle.fast_extract(want, border, w, h, lb, bl, cap=1 << 16)
                kp = ctx.fastExtract(w, h, out, border=border, logBucketSize=lb, bucketLimit=bl)
                assert_same(kp, okp, "fastExtract " + tag)
                if border == 16 and len(okp):
                    for words in (8, 3):
                        assert_same(ctx.orbCompute(img, okp, words=words), oracle.orb_compute(img, okp, words),
                                    "orbCompute words %d %s" % (words, tag))


def test_fast_detect_preserves_untouched_bytes(ctx, golden, oracle):
    """The reference only writes rows [border,h-border) x the 16-px blocks; everything else in `out` stays."""
    vstep, w, h, border, ft, ht = golden["small_params"].tolist()[1]
    img = np.ascontiguousarray(golden["small1_img"])
    out = np.full_like(img, 0x5a)
    want = out.copy()
    oracle.fast_detect(img, want, border, w, h, ft)
    ctx.fastDetect(w, h, img, out, ft, border=border)
    assert_same(out, want, "fastDetect on a dirty out buffer")


def test_fast_extract_tie_heavy(ctx, golden):
    for i, (vstep, w, h, border) in enumerate(golden["tie_params"].tolist()):
        m = np.ascontiguousarray(golden["tie%d_out" % i])
        assert_same(ctx.fastExtract(w, h, m, border=border), golden["tie%d_kp" % i], "tie %d" % i)
        assert_same(ctx.fastExtract(w, h, m, border=border, logBucketSize=3, bucketLimit=2),
                    golden["tie%d_kp_b3_2" % i], "tie %d <3,2>" % i)
        assert_same(ctx.fastExtract(w, h, m, border=border, logBucketSize=5, bucketLimit=4),
                    golden["tie%d_kp_b5_4" % i], "tie %d <5,4>" % i)


def test_fast_extract_results_append(ctx, golden):
    m = np.ascontiguousarray(golden["tie0_out"])
    vstep, w, h, border = golden["tie_params"].tolist()[0]
    results = [123]
    kp = ctx.fastExtract(w, h, m, results, border=border)
    assert results[0] == 123 and results[1:] == [int(v) for v in kp]  # Fast.h:320 appends


def test_literal_stages_fixture_levels(ctx, golden, fixture_stack):
    out = np.zeros_like(fixture_stack)
    for w, h, r in zip(VGA.widths, VGA.heights, VGA.row_offsets):
        ctx.fastDetect(w, h, fixture_stack, out, 20, row0=r)
    nz = np.flatnonzero(out)
    assert_same(nz, golden["fixture_mask_nz"], "fixture FAST mask")
    kps = []
    for w, h, r in zip(VGA.widths, VGA.heights, VGA.row_offsets):
        ctx.fastScoreHarris(w, h, fixture_stack, 1 << 15, out, row0=r)
        kps.append(ctx.fastExtract(w, h, out, row0=r))
    nz = np.flatnonzero(out)
    assert_same(nz, golden["fixture_out_nz"], "fixture score map support")
    assert_same(out.ravel()[nz], golden["fixture_out_val"], "fixture score bytes")
    assert_same(np.concatenate(kps), golden["fixture_kp"], "fixture keypoints")


def test_harris_points(ctx, oracle):
    rng = np.random.Generator(np.random.PCG64(5))
    img = rng.integers(0, 256, size=(64, 128), dtype=np.uint8)
    xs = rng.integers(3, 120, size=500)
    ys = rng.integers(3, 59, size=500)
    for thr in (-(1 << 31), -1, 0, 1 << 15, (1 << 31) - 1):
        got = ctx.harrisScoreSobel(img, xs, ys, thr)
        want = np.asarray([oracle.harris_score_sobel(img, int(x), int(y), thr) for x, y in zip(xs, ys)], np.uint8)
        assert_same(got, want, "harrisScoreSobel thr %d" % thr)


def test_orb_literal(ctx, golden, fixture_stack):
    img, kp = np.ascontiguousarray(golden["orb_img"]), golden["orb_kp"]
    assert_same(ctx.orbCentroids(img, kp), golden["orb_centroids"], "orbCentroids")
    assert_same(ctx.atan2(golden["orb_centroids"]), golden["orb_angles"], "atan2")
    assert_same(ctx.atan2(golden["angle_xys"]), golden["angle_bins"], "atan2 adversarial")
    for w in (1, 4, 5, 8):
        assert_same(ctx.orbCompute(img, kp, words=w), golden["orb_desc_w%d" % w], "orbCompute words=%d" % w)
    # the whole fixture through the literal orbCompute with demo-style stacked y
    kps = golden["fixture_kp_stacked"]
    assert_same(ctx.orbCentroids(fixture_stack, kps), golden["fixture_centroids"], "fixture centroids")
    d, a = ctx.orbCompute(fixture_stack, kps, want_angles=True)
    assert_same(a, golden["fixture_angles"][:len(kps)], "fixture angles")
    assert_same(d, golden["fixture_desc"], "fixture descriptors")
    descriptors = []
    ctx.orbCompute(img, kp, descriptors, words=8)
    assert descriptors == [int(v) for v in golden["orb_desc_w8"].ravel()]  # Orb.h:423 appends


def test_brief_describe(ctx, golden, oracle):
    img = np.ascontiguousarray(golden["orb_img"])
    for rot in range(30):
        assert_same(ctx.briefDescribe(img, 40 + rot, 30, rot), oracle.brief_describe(img, 40 + rot, 30, rot),
                    "briefDescribe rot %d" % rot)
    for rot in (-1, 30, 255):  # Brief.h:641-732: untouched
        assert (ctx.briefDescribe(img, 40, 40, rot) == 0xdeadbeef).all()
    assert_same(api.brief_table(), golden["brief_table"], "brief table")


def test_zero_moment_patch_bin(ctx):
    img = np.full((64, 128), 77, np.uint8)
    d, a = ctx.orbCompute(img, np.asarray([api.encodeFast(200, 40, 30)], np.uint32), want_angles=True)
    # a flat patch has m01 == 0 but m10 == 0 too only if the disc is symmetric -- it is; bin 7 (SURVEY spec O2)
    assert a[0] == 7 and (d == 0).all()


# ------------------------------------------------------------------------------------------ fused path
def oracle_pyramid(oracle, stack, lay, **kw):
    return oracle.extract_pyramid(stack, lay.widths, lay.heights, lay.row_offsets, **kw)


@pytest.mark.parametrize("tma", [True, False])
def test_fused_fixture(ctx, golden, fixture_stack, oracle, tma):
    ctx.set_tma(tma)
    try:
        kp, lv, desc = ctx.extract_pyramid(fixture_stack, api.make_desc(VGA), cap=4096)
    finally:
        ctx.set_tma(True)
    assert len(kp) == 1754
    assert_same(kp, golden["fixture_kp"], "fused fixture keypoints (tma=%s)" % tma)
    assert_same(lv, golden["fixture_kp_level"], "fused fixture levels")
    assert_same(desc, golden["fixture_desc"], "fused fixture descriptors")


@pytest.mark.parametrize("lb,bl", [(4, 3), (4, 5), (3, 2), (5, 4), (1, 1), (2, 1)])
def test_fused_fixture_bucketed(ctx, golden, fixture_stack, oracle, lb, bl):
    kp, lv, desc = ctx.extract_pyramid(fixture_stack, api.make_desc(VGA, log_bucket=lb, bucket_limit=bl), cap=8192)
    okp, olv, odesc = oracle_pyramid(oracle, fixture_stack, VGA, log_bucket=lb, bucket_limit=bl)
    key = "fixture_kp_b%d_%d" % (lb, bl)
    if key in golden:
        assert_same(okp, golden[key], "oracle vs golden")
    assert_same(kp, okp, "fused bucketed <%d,%d> keypoints" % (lb, bl))
    assert_same(lv, olv, "levels")
    assert_same(desc, odesc, "descriptors")


@pytest.mark.parametrize("name,lay,seed", [("vga", VGA, 0), ("vga", VGA, 1), ("fhd", FHD, 0)])
def test_fused_synthetic(ctx, oracle, name, lay, seed):
    stack = synthetic_stack(lay, seed)
    for lb, bl in ((0, 5), (4, 3)):
        kp, lv, desc = ctx.extract_pyramid(stack, api.make_desc(lay, log_bucket=lb, bucket_limit=bl), cap=1 << 15)
        okp, olv, odesc = oracle_pyramid(oracle, stack, lay, log_bucket=lb, bucket_limit=bl)
        assert len(okp) > 300
        assert_same(kp, okp, "%s seed %d <%d,%d> keypoints" % (name, seed, lb, bl))
        assert_same(lv, olv, "levels")
        assert_same(desc, odesc, "descriptors")


def test_fused_uhd_level_local_y(ctx, oracle):
    """4K stack is 9946 rows: keypoints stay level-local (12-bit y, Util.h:27-29)."""
    stack = synthetic_stack(UHD, 0)
    kp, lv, desc = ctx.extract_pyramid(stack, api.make_desc(UHD, log_bucket=4, bucket_limit=3), cap=1 << 16)
    okp, olv, odesc = oracle_pyramid(oracle, stack, UHD, log_bucket=4, bucket_limit=3)
    assert_same(kp, okp, "4K keypoints")
    assert_same(lv, olv, "4K levels")
    assert_same(desc, odesc, "4K descriptors")


def test_fused_noise_and_thresholds(ctx, oracle):
    """Adversarial uniform noise: every tile is full of corners, ties and overrun-column flags."""
    lay = pyramid_layout(640, 480, nlevels=4)
    stack = noise_stack(lay, 11)
    for ft, ht, lb, bl in ((40, 1 << 10, 0, 5), (20, 1 << 15, 4, 3), (130, -1, 2, 2), (255, 0, 0, 5)):
        d = api.make_desc(lay, fast_threshold=ft, harris_threshold=ht, log_bucket=lb, bucket_limit=bl)
        kp, lv, desc = ctx.extract_pyramid(stack, d, cap=1 << 17)
        okp, olv, odesc = oracle_pyramid(oracle, stack, lay, fast_threshold=ft, harris_threshold=ht, log_bucket=lb,
                                         bucket_limit=bl, cap=1 << 17)
        assert_same(kp, okp, "noise ft=%d keypoints" % ft)
        assert_same(desc, odesc, "noise descriptors")


def test_fused_small_thresholds_on_noise(ctx, oracle):
    """FAST thresholds 0..5 and 126/127 on uniform noise: ties |p-c| == t next to large differences are everywhere,
    which is where the carry-tolerant pre-test of phase A admits false candidates (and must lose no corner)."""
    lay = pyramid_layout(256, 160, nlevels=2)
    stack = noise_stack(lay, 17)
    for ft in (0, 1, 2, 5, 126, 127):
        d = api.make_desc(lay, fast_threshold=ft, harris_threshold=1 << 14, log_bucket=0, bucket_limit=5)
        kp, lv, desc = ctx.extract_pyramid(stack, d, cap=1 << 16)
        okp, olv, odesc = oracle_pyramid(oracle, stack, lay, fast_threshold=ft, harris_threshold=1 << 14, cap=1 << 16)
        assert len(okp) > 0 or ft >= 126
        assert_same(kp, okp, "noise ft=%d keypoints" % ft)
        assert_same(desc, odesc, "noise ft=%d descriptors" % ft)


def test_fused_ragged_geometries(ctx, oracle):
    """Odd / tiny / non-multiple-of-16 levels, a vstep that is not a multiple of 16 (plain-load staging),
    levels with no interior, border 16 and larger."""
    rng = np.random.Generator(np.random.PCG64(3))
    for vstep, dims, border in ((200, [(200, 77), (167, 64), (33, 40), (139, 33), (50, 31)], 16),
                                (203, [(203, 90), (170, 75)], 16), (256, [(256, 100), (213, 83)], 20)):
        offs, r = [], 0
        for w, h in dims:
            offs.append(r)
            r += h
        lay = PyramidLayout(vstep, [w for w, h in dims], [h for w, h in dims], offs)
        stack = np.zeros((r, vstep), np.uint8)
        for (w, h), o in zip(dims, offs):
            a = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
            stack[o:o + h, :w] = a
        d = api.make_desc(lay, border=border, fast_threshold=35, harris_threshold=1 << 12)
        kp, lv, desc = ctx.extract_pyramid(stack, d, cap=1 << 15)
        okp, olv, odesc = oracle_pyramid(oracle, stack, lay, border=border, fast_threshold=35,
                                         harris_threshold=1 << 12, cap=1 << 15)
        assert_same(kp, okp, "ragged vstep %d keypoints" % vstep)
        assert_same(lv, olv, "levels")
        assert_same(desc, odesc, "descriptors")


def test_fused_batch_and_capacity(ctx, oracle):
    nf = 5
    stacks = np.stack([synthetic_stack(VGA, 10 + f) for f in range(nf)])
    # junk right of each level: host batches are uploaded level by level without the padding columns, and
    # nothing the reference computes with border 16 may depend on them
    rng = np.random.Generator(np.random.PCG64(8))
    for w, h, r in zip(VGA.widths, VGA.heights, VGA.row_offsets):
        if w < VGA.vstep:
            stacks[1:, r:r + h, w:] = rng.integers(0, 256, size=(nf - 1, h, VGA.vstep - w), dtype=np.uint8)
    d = api.make_desc(VGA, log_bucket=4, bucket_limit=3)
    res = ctx.extract_batch(stacks, d, cap=4096)
    for f in range(nf):
        okp, olv, odesc = oracle_pyramid(oracle, stacks[f], VGA, log_bucket=4, bucket_limit=3)
        kp, lv, desc, level_start, n_total = res[f]
        assert n_total == len(okp)
        assert_same(kp, okp, "batch frame %d keypoints" % f)
        assert_same(desc, odesc, "batch frame %d descriptors" % f)
        want_start = [int((olv < l).sum()) for l in range(9)]
        assert level_start[:9].tolist() == want_start
    # capacity overflow is reported, with the count to size for
    with pytest.raises(api.PislamError) as e:
        ctx.extract_pyramid(stacks[0], d, cap=100)
    assert e.value.status == api.PISLAM_ERR_CAPACITY
    # and the context keeps working afterwards
    kp, _, _ = ctx.extract_pyramid(stacks[0], d, cap=4096)
    assert_same(kp, res[0][0], "after overflow")


def test_fused_batch_beyond_one_descriptor_launch(ctx, oracle):
    """1030 small frames: orb_tma_kernel numbers the keypoints of at most 1024 frames per launch, so this batch
    takes two descriptor launches (second one with a frame offset); all-black frames in between have no keypoints
    and must be skipped by the flat numbering."""
    lay = pyramid_layout(160, 96, nlevels=2)
    rng = np.random.Generator(np.random.PCG64(21))
    src = [synthetic_stack(lay, 40), noise_stack(lay, 41), np.zeros((lay.rows, lay.vstep), np.uint8),
           synthetic_stack(lay, 42), np.zeros((lay.rows, lay.vstep), np.uint8), noise_stack(lay, 43)]
    want = [oracle_pyramid(oracle, s_, lay, fast_threshold=30, harris_threshold=1 << 13, cap=1 << 14) for s_ in src]
    assert any(len(w[0]) == 0 for w in want) and any(len(w[0]) > 50 for w in want)
    nf = 1030
    which = rng.integers(0, len(src), size=nf)
    which[1022:1028] = [1, 2, 3, 0, 5, 4]  # both sides of the launch boundary
    stacks = np.stack([src[k] for k in which])
    d = api.make_desc(lay, fast_threshold=30, harris_threshold=1 << 13)
    cap = max(len(w[0]) for w in want) + 8
    res = ctx.extract_batch(stacks, d, cap=cap)
    assert len(res) == nf
    for f in range(nf):
        okp, olv, odesc = want[which[f]]
        kp, lv, desc, level_start, n_total = res[f]
        assert n_total == len(okp), "frame %d" % f
        assert_same(kp, okp, "frame %d keypoints" % f)
        assert_same(desc, odesc, "frame %d descriptors" % f)


def test_fused_keypoints_only_and_words(ctx, oracle):
    stack = synthetic_stack(VGA, 2)
    for words in (0, 1, 4, 8):
        kp, lv, desc = ctx.extract_pyramid(stack, api.make_desc(VGA, words=words), cap=8192)
        okp, olv, odesc = oracle_pyramid(oracle, stack, VGA, words=max(words, 1))
        assert_same(kp, okp, "words=%d keypoints" % words)
        if words:
            assert_same(desc, odesc, "words=%d descriptors" % words)


@pytest.mark.parametrize("tma", [True, False])
@pytest.mark.parametrize("border", [5, 8, 9, 12])
def test_fused_keypoints_only_small_borders(ctx, oracle, border, tma):
    """Keypoints-only extraction (words = 0) admits borders 5..15: borders that are multiples of 4 stage the tiles
    by TMA, the others by plain loads (and `tma=False` forces plain loads for all of them).  Synthetic, ragged and
    noise levels, plain and bucketed."""
    rng = np.random.Generator(np.random.PCG64(100 + border))
    lay = pyramid_layout(320, 240, nlevels=3)
    noise = noise_stack(lay, 50 + border)
    dims = [(208, 61), (173, 64), (2 * border + 1, 40), (139, 2 * border + 2), (50, 31)]
    offs, r = [], 0
    for w, h in dims:
        offs.append(r)
        r += h
    rag = PyramidLayout(208, [w for w, h in dims], [h for w, h in dims], offs)
    rstack = np.zeros((r, 208), np.uint8)
    for (w, h), o in zip(dims, offs):
        coarse = rng.integers(0, 256, size=(h // 5 + 2, w // 5 + 2), dtype=np.uint8)
        rstack[o:o + h, :w] = np.kron(coarse, np.ones((5, 5), np.uint8))[:h, :w]
    ctx.set_tma(tma)
    try:
        for name, l, stack, ft, ht in (("synthetic", lay, synthetic_stack(lay, 60 + border), 20, 1 << 15),
                                       ("noise", lay, noise, 45, 1 << 12), ("ragged", rag, rstack, 25, 1 << 10)):
            for lb, bl in ((0, 5), (3, 2), (5, 1)):
                d = api.make_desc(l, border=border, fast_threshold=ft, harris_threshold=ht, log_bucket=lb,
                                  bucket_limit=bl, words=0)
                kp, lv, _ = ctx.extract_pyramid(stack, d, cap=1 << 16)
                okp, olv, _ = oracle_pyramid(oracle, stack, l, border=border, fast_threshold=ft, harris_threshold=ht,
                                             log_bucket=lb, bucket_limit=bl, words=0, cap=1 << 16)
                assert len(okp) > 0
                tag = "%s border %d tma %s buckets <%d,%d>" % (name, border, tma, lb, bl)
                assert_same(kp, okp, tag + " keypoints")
                assert_same(lv, olv, tag + " levels")
        # descriptors with a border below 16 are refused, not silently mis-computed
        with pytest.raises(api.PislamError) as e:
            ctx.extract_pyramid(noise, api.make_desc(lay, border=border, words=8), cap=1 << 16)
        assert e.value.status == api.PISLAM_ERR_UNSUPPORTED
    finally:
        ctx.set_tma(True)


def test_geometry_cache_alternating(ctx, oracle):
    """A caller alternating geometries (two cameras, or the literal fastExtract level by level) hits the per-context
    geometry cache: after the first round no tile table is uploaded again, and results stay bit-exact."""
    lays = [pyramid_layout(320, 240, nlevels=3), pyramid_layout(256, 160, nlevels=2), pyramid_layout(200, 120, nlevels=2)]
    stacks = [synthetic_stack(l, 90 + i) for i, l in enumerate(lays)]
    want = [oracle_pyramid(oracle, s_, l, log_bucket=4, bucket_limit=3) for s_, l in zip(stacks, lays)]
    descs = [api.make_desc(l, log_bucket=4, bucket_limit=3) for l in lays]
    uploads = []
    for rnd in range(4):
        for i in range(len(lays)):
            kp, lv, desc = ctx.extract_pyramid(stacks[i], descs[i], cap=4096)
            assert_same(kp, want[i][0], "round %d geometry %d keypoints" % (rnd, i))
            assert_same(desc, want[i][2], "round %d geometry %d descriptors" % (rnd, i))
        uploads.append(ctx.geometry_uploads)
    assert uploads[1] == uploads[0] and uploads[3] == uploads[0], uploads


def test_literal_keypoints_outside_image_rejected(ctx):
    """Keypoints whose 31x31 patch (or 8x8 Harris window) leaves the image are refused instead of read out of bounds."""
    img = np.full((64, 128), 9, np.uint8)
    for x, y in ((3, 3), (120, 60), (10, 62)):
        with pytest.raises(api.PislamError) as e:
            ctx.orbCompute(img, np.asarray([api.encodeFast(200, x, y)], np.uint32))
        assert e.value.status == api.PISLAM_ERR_INVALID_ARG
        with pytest.raises(api.PislamError) as e:
            ctx.orbCentroids(img, np.asarray([api.encodeFast(200, x, y)], np.uint32))
        assert e.value.status == api.PISLAM_ERR_INVALID_ARG
    with pytest.raises(api.PislamError) as e:
        ctx.harrisScoreSobel(img, [1], [1], 0)
    assert e.value.status == api.PISLAM_ERR_INVALID_ARG
    with pytest.raises(api.PislamError) as e:
        ctx.harrisScoreSobel(img, [126], [62], 0)
    assert e.value.status == api.PISLAM_ERR_INVALID_ARG
    with pytest.raises(ValueError):
        ctx.extract_pyramid(np.zeros((100, 640), np.uint8), api.make_desc(VGA), cap=64)


def test_async_host_pipeline_back_to_back(ctx, oracle):
    """pislam_b200_extract_batch_async / _extract_frames_async on host buffers: several calls issued back to back
    (double-buffered device staging: the copies of call k+1 overlap the kernels of call k) into separate record
    buffers, one synchronize at the end; every call's records equal the blocking call's and the oracle's."""
    import torch
    from pislam_b200.pyramid import build_stack, synthetic_base
    lay = pyramid_layout(320, 240, nlevels=4)
    d = api.make_desc(lay, log_bucket=4, bucket_limit=3)
    cap, nf, ncalls = 1024, 9, 5
    rec_bytes = api.record_layout(cap, 8)["bytes"]
    bases = [[synthetic_base(320, 240, 0x5EED0000 + 900 + 10 * k + f) for f in range(nf)] for k in range(ncalls)]
    h_bases = [torch.from_numpy(np.stack(b)).pin_memory() for b in bases]
    h_stacks = [torch.from_numpy(np.stack([build_stack(x, lay) for x in b])).pin_memory() for b in bases]
    for kind in ("stacks", "bases"):
        outs = [torch.zeros(nf * rec_bytes, dtype=torch.uint8).pin_memory() for _ in range(ncalls)]
        for k in range(ncalls):
            if kind == "stacks":
                st = ctx.extract_batch_ptr(h_stacks[k].data_ptr(), lay.stack_bytes, nf, d, cap, outs[k].data_ptr(),
                                           asynchronous=True)
            else:
                st = ctx.extract_frames_ptr(h_bases[k].data_ptr(), 320, 320 * 240, nf, d, cap, outs[k].data_ptr(),
                                            asynchronous=True)
            assert st == 0
        ctx.synchronize()
        for k in range(ncalls):
            recs = api.parse_records(outs[k].numpy(), nf, cap, 8)
            for f in (0, nf // 2, nf - 1):
                okp, olv, odesc = oracle_pyramid(oracle, h_stacks[k].numpy()[f], lay, log_bucket=4, bucket_limit=3)
                assert_same(recs[f][0], okp, "%s call %d frame %d keypoints" % (kind, k, f))
                assert_same(recs[f][2], odesc, "%s call %d frame %d descriptors" % (kind, k, f))
            blocking = ctx.extract_batch(h_stacks[k].numpy(), d, cap=cap)
            for f in range(nf):
                assert_same(recs[f][0], blocking[f][0], "%s call %d frame %d vs blocking" % (kind, k, f))
                assert_same(recs[f][2], blocking[f][2], "%s call %d frame %d descriptors vs blocking" % (kind, k, f))


def test_argument_errors(ctx):
    stack = np.zeros((VGA.rows, VGA.vstep), np.uint8)
    with pytest.raises(api.PislamError) as e:
        ctx.extract_pyramid(stack, api.make_desc(VGA, border=3), cap=64)
    assert e.value.status == api.PISLAM_ERR_UNSUPPORTED
    with pytest.raises(api.PislamError) as e:
        ctx.extract_pyramid(stack, api.make_desc(VGA, words=9), cap=64)
    assert e.value.status == api.PISLAM_ERR_INVALID_ARG
    kp, lv, desc = ctx.extract_pyramid(stack, api.make_desc(VGA), cap=64)  # empty image: no keypoints
    assert len(kp) == 0


# ------------------------------------------------------------------------------------------ full-size properties
def test_full_batch_properties(ctx, oracle):
    """BASELINE-size batch (256 VGA pyramids, buckets <4,3>): size-independent properties instead of 256 oracle runs.
    (1) a frame's record does not depend on its position in the batch or on its neighbours (every copy of a
    frame is identical, and equals the frame extracted alone); (2) keypoints are level-sorted and inside the
    cell-covered area; (3) 16 sampled frames are bit-exact against the oracle; (4) a second run is identical."""
    nf, ndistinct = 256, 16
    base = [synthetic_stack(VGA, 40 + i) for i in range(ndistinct)]
    stacks = np.stack([base[i % ndistinct] for i in range(nf)])
    d = api.make_desc(VGA, log_bucket=4, bucket_limit=3)
    res = ctx.extract_batch(stacks, d, cap=2048)
    res2 = ctx.extract_batch(stacks, d, cap=2048)
    for f in range(nf):
        kp, lv, desc, level_start, n_total = res[f]
        kp0, lv0, desc0, _, _ = res[f % ndistinct]
        assert n_total == len(kp)
        assert np.array_equal(kp, kp0) and np.array_equal(lv, lv0) and np.array_equal(desc, desc0), f
        assert np.array_equal(kp, res2[f][0]) and np.array_equal(desc, res2[f][2]), f
        assert (np.diff(lv.astype(np.int32)) >= 0).all()
        x, y = (kp >> 12) & 0xFFF, kp & 0xFFF
        w = np.asarray(VGA.widths)[lv]
        h = np.asarray(VGA.heights)[lv]
        assert (x >= 16).all() and (x <= w - 16).all() and (y >= 16).all() and (y <= h - 16).all()
        assert ((kp >> 24) > 0).all()
    for f in range(ndistinct):
        okp, olv, odesc = oracle_pyramid(oracle, base[f], VGA, log_bucket=4, bucket_limit=3)
        assert_same(res[f][0], okp, "frame %d keypoints" % f)
        assert_same(res[f][2], odesc, "frame %d descriptors" % f)
        alone = ctx.extract_pyramid(base[f], d, cap=2048)
        assert_same(alone[0], okp, "frame %d alone" % f)


def test_shifted_image_shifts_keypoints(ctx):
    """Translating the image by a multiple of the cell size translates the interior keypoints: detection,
    scoring and NMS are position-independent (SURVEY spec F3)."""
    lay = pyramid_layout(640, 480, nlevels=1)
    a = synthetic_stack(lay, 77)
    b = np.zeros_like(a)
    dx, dy = 6, 4
    b[dy:, dx:] = a[:-dy, :-dx]
    d = api.make_desc(lay, words=8)
    ka, _, da = ctx.extract_pyramid(a, d, cap=1 << 14)
    kb, _, db = ctx.extract_pyramid(b, d, cap=1 << 14)
    # compare keypoints well inside both images (3x3 NMS + patch radius away from the moved border)
    def inside(k, ox, oy):
        x, y = ((k >> 12) & 0xFFF).astype(int) - ox, (k & 0xFFF).astype(int) - oy
        return (x >= 40) & (x < 560) & (y >= 40) & (y < 400)
    sa = {(int(k >> 24), int((k >> 12) & 0xFFF), int(k & 0xFFF)): tuple(v) for k, v in zip(ka[inside(ka, 0, 0)], da[inside(ka, 0, 0)])}
    sb = {(int(k >> 24), int((k >> 12) & 0xFFF) - dx, int(k & 0xFFF) - dy): tuple(v)
          for k, v in zip(kb[inside(kb, dx, dy)], db[inside(kb, dx, dy)])}
    assert len(sa) > 50 and sa == sb


def test_concurrent_contexts(oracle):
    """The reference is re-entrant and safe to call concurrently on disjoint buffers:
    one pislam_ctx per host thread, four threads hammering fused and literal calls at once."""
    import threading
    lays = [pyramid_layout(320, 240, nlevels=4), VGA]
    errors = []

    def work(t):
        try:
            lay = lays[t % 2]
            stack = synthetic_stack(lay, 300 + t)
            okp, olv, odesc = oracle_pyramid(oracle, stack, lay, log_bucket=4, bucket_limit=3)
            with api.Context(0) as c:
                for _ in range(10):
                    kp, lv, desc = c.extract_pyramid(stack, api.make_desc(lay, log_bucket=4, bucket_limit=3), cap=4096)
                    assert np.array_equal(kp, okp) and np.array_equal(desc, odesc)
                    out = np.zeros_like(stack)
                    c.fastDetect(lay.widths[0], lay.heights[0], stack, out, 20)
                    c.fastScoreHarris(lay.widths[0], lay.heights[0], stack, 1 << 15, out)
                    k0 = c.fastExtract(lay.widths[0], lay.heights[0], out, logBucketSize=4, bucketLimit=3)
                    assert np.array_equal(k0, okp[olv == 0])
        except Exception as e:  # noqa: BLE001
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
