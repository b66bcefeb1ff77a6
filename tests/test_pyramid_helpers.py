"""Pyramid-preparation helpers: gaussian5x5, bilinear7_8, bilinear13_16.
CPU part: the oracle restatement against the reference's Bilinear.h compiled through the shim (when present)
and against an independent transcription of the reference's scalar Gaussian test model; GPU part: the CUDA
kernels against the oracle over the dimension sweeps the reference's own tests use
(test/GaussianTest.cpp:156 -> 16..63, test/BilinearTest.cpp:168 -> 1..47), exact equality per byte."""
import numpy as np
import pytest

RNG = np.random.Generator(np.random.PCG64(2024))


def spiral(h, w):
    """a binary log-spiral-ish pattern + noise: edges at every orientation (after test/TestUtil.cpp:27-52)"""
    yy, xx = np.mgrid[0:h, 0:w]
    r = np.hypot(yy - h / 2.0, xx - w / 2.0) + 1.0
    th = np.arctan2(yy - h / 2.0, xx - w / 2.0)
    return (((np.log(r) * 3.0 + th) % 1.5) < 0.75).astype(np.uint8) * 255


def padded(h, w, B, vstep=128, rows=80, kind="noise"):
    img = np.zeros((rows, vstep), np.uint8)
    H, W = (h + B - 1) // B * B, (w + B - 1) // B * B
    img[:H, :W] = RNG.integers(0, 256, size=(H, W), dtype=np.uint8) if kind == "noise" else spiral(H, W)
    return img


def rhadd(a, b):
    return (a >> 1) + (b >> 1) + ((a | b) & 1)


def gaussian_model(m, w, h):
    """transcription of the running-window model in test/GaussianTest.cpp:159-215"""
    m = m.astype(np.int64).copy()
    for axis in (0, 1):
        n_outer, n_inner = (w, h) if axis == 0 else (h, w)
        for o in range(n_outer):
            get = (lambda i: m[i, o]) if axis == 0 else (lambda i: m[o, i])
            a, b, c, d = get(2), get(1), get(0), get(1)
            for i in range(n_inner):
                e = c if i == n_inner - 2 else (a if i == n_inner - 1 else get(i + 2))
                x = rhadd(rhadd(rhadd(a, e), c), c)
                v = rhadd(x, rhadd(b, d))
                if axis == 0:
                    m[i, o] = v
                else:
                    m[o, i] = v
                a, b, c, d = b, c, d, e
    return m.astype(np.uint8)


def test_oracle_gaussian_matches_reference_test_model(oracle):
    for w, h in ((16, 16), (17, 23), (63, 40), (33, 16), (64, 64), (3, 3), (5, 4)):
        img = RNG.integers(0, 256, size=(70, 128), dtype=np.uint8)
        out = np.full_like(img, 0x5a)
        oracle.gaussian5x5(img, out, w, h)
        assert np.array_equal(out[:h, :w], gaussian_model(img, w, h)[:h, :w]), (w, h)
        assert (out[h:] == 0x5a).all() and (out[:h, w:] == 0x5a).all()


def test_oracle_bilinear_matches_reference_build(oracle, reference):
    for name, B in (("bilinear7_8", 8), ("bilinear13_16", 16)):
        for w in range(1, 48, 2):
            for h in range(1, 48, 3):
                img = padded(h, w, B)
                a, b = np.full_like(img, 0x5a), np.full_like(img, 0x5a)
                getattr(oracle, name)(img, a, w, h)
                getattr(reference, name)(img.copy(), b, w, h)
                assert np.array_equal(a, b), (name, w, h)
                c, d = img.copy(), img.copy()  # in place, as the reference allows
                getattr(oracle, name)(c, c, w, h)
                getattr(reference, name)(d, d, w, h)
                assert np.array_equal(c, d), (name, w, h, "in place")


def test_documented_output_sizes(oracle):
    """Bilinear.h:33-35 / :156-158: a 39x29 image scales to 34x25 (7/8) and 31x23 (13/16) -- rounded down."""
    assert (39 * 7 // 8, 29 * 7 // 8) == (34, 25) and (39 * 13 // 16, 29 * 13 // 16) == (31, 23)


@pytest.fixture(scope="module")
def ctx():
    from pislam_b200 import api
    c = api.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["noise", "spiral"])
def test_gpu_gaussian_dimension_sweep(ctx, oracle, kind):
    for w in list(range(16, 64, 3)) + [63]:
        for h in list(range(16, 64, 5)) + [63]:
            img = padded(h, w, 1, kind=kind)
            got, want = np.full_like(img, 0x5a), np.full_like(img, 0x5a)
            ctx.gaussian5x5(w, h, img, got)
            oracle.gaussian5x5(img, want, w, h)
            assert np.array_equal(got, want), (w, h)
    # VGA, and in place
    img = RNG.integers(0, 256, size=(480, 640), dtype=np.uint8)
    want = np.zeros_like(img)
    oracle.gaussian5x5(img, want, 640, 480)
    got = img.copy()
    ctx.gaussian5x5(640, 480, got, got)
    assert np.array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("name,B", [("bilinear7_8", 8), ("bilinear13_16", 16)])
def test_gpu_bilinear_dimension_sweep(ctx, oracle, name, B):
    for w in range(1, 48, 2):
        for h in range(1, 48, 3):
            img = padded(h, w, B)
            got, want = np.full_like(img, 0x5a), np.full_like(img, 0x5a)
            getattr(ctx, name)(w, h, img, got)
            getattr(oracle, name)(img, want, w, h)
            assert np.array_equal(got, want), (name, w, h)
    img = RNG.integers(0, 256, size=(480, 640), dtype=np.uint8)
    want = img.copy()
    getattr(oracle, name)(want, want, 640, 480)
    got = img.copy()
    getattr(ctx, name)(640, 480, got, got)  # in place
    assert np.array_equal(got, want)
