"""The device-side arithmetic of pislam_b200/csrc/device_ops.cuh compiled for the HOST (tests/device_arith_host.cpp:
g++ + scalar shims of the CUDA intrinsics) and checked against plain per-pixel models and the oracle -- no GPU.

* the two FAST-9 formulations of the kernels (serial masks; packed two-polarity with the PRMT sign gather) equal the
  ring definition (Fast.h:54-158) on every pixel, for thresholds across the range;
* the SWAR pre-tests of the cascade (N|S any, compass, diagonal; 5 and 6 rows per task) are supersets of their exact
  per-pixel definitions, contain every FAST-9 corner, have tie-only false positives below t = 128 and none above;
* the Harris code expression equals the oracle's harrisScoreSobel (Harris.h:37-248).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from pislam_b200.pyramid import noise_stack, pyramid_layout, synthetic_stack

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
u8p = C.POINTER(C.c_uint8)


def _cuda_include():
    for p in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if p and os.path.exists(os.path.join(p, "include", "cuda_runtime.h")):
            return os.path.join(p, "include")
    return None


@pytest.fixture(scope="module")
def dah(tmp_path_factory):
    inc = _cuda_include()
    if inc is None:
        pytest.skip("CUDA headers not found")
    out = str(tmp_path_factory.mktemp("dah") / "libdah.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-I" + inc, "-o", out,
                           os.path.join(HERE, "device_arith_host.cpp")])
    lib = C.CDLL(out)
    lib.dah_fast9.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.dah_swar.argtypes = [u8p, C.c_int, C.c_int]
    lib.dah_harris.argtypes = [u8p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int32, u8p]
    lib.dah_prmt.argtypes = [C.c_uint32] * 3
    lib.dah_prmt.restype = C.c_uint32
    return lib


def _ptr(a):
    return a.ctypes.data_as(u8p)


def _images():
    lay = pyramid_layout(160, 120, nlevels=1)
    rng = np.random.Generator(np.random.PCG64(11))
    yield "synthetic", np.ascontiguousarray(synthetic_stack(lay, 3)[:120])
    yield "noise", np.ascontiguousarray(noise_stack(lay, 4)[:120])
    # two grey levels d apart on noise-free blocks: exercises the ties |p - c| == t exactly
    blocks = (rng.integers(0, 2, size=(30, 40), dtype=np.uint8) * 20 + 100).repeat(4, 0).repeat(4, 1)
    yield "ties", np.ascontiguousarray(blocks)
    yield "extremes", np.ascontiguousarray((rng.integers(0, 2, size=(120, 160), dtype=np.uint8) * 255))


def test_prmt_model_matches_ptx_semantics(dah):
    assert dah.dah_prmt(0x03020100, 0x07060504, 0x7531) == 0x07050301
    # selector 0xfbd9: result bytes 0..3 = the sign of source bytes 1, 5, 3, 7 (a = bytes 0-3, b = bytes 4-7)
    assert dah.dah_prmt(0x80008000, 0x00000000, 0xfbd9) == 0x00ff00ff
    assert dah.dah_prmt(0x80000080, 0x00800000, 0xfbd9) == 0x00ff0000
    assert dah.dah_prmt(0x00008000, 0x80000000, 0xfbd9) == 0xff0000ff


def test_fast9_variants_equal_the_ring_definition(dah):
    total = 0
    for name, img in _images():
        h, w = img.shape
        for t in (0, 1, 5, 19, 20, 21, 60, 127, 128, 200, 254, 255):
            n = C.c_int(0)
            bad = dah.dah_fast9(_ptr(img), w, w, h, t, C.byref(n))
            assert bad == 0, (name, t, bad)
            total += n.value
    assert total > 1000


def test_swar_pretests_are_supersets_with_tie_only_false_positives(dah):
    rng = np.random.Generator(np.random.PCG64(12))
    for name, img in _images():
        tile = np.ascontiguousarray(img[:76, :160])  # the tile kernel's staged shape: 76 rows x 40 words
        for t in (0, 3, 20, 21, 100, 127, 128, 129, 250):
            assert dah.dah_swar(_ptr(tile), tile.shape[0], t) == 0, (name, t)
    # random tiles with a small dynamic range (dense ties)
    for t in (2, 20):
        tile = rng.integers(100, 100 + 2 * t + 3, size=(76, 160), dtype=np.uint8)
        assert dah.dah_swar(_ptr(tile), 76, t) == 0


def test_harris_expression_equals_the_oracle(dah, oracle):
    rng = np.random.Generator(np.random.PCG64(13))
    for name, img in _images():
        h, w = img.shape
        n = 400
        xs = rng.integers(4, w - 5, size=n).astype(np.int32)
        ys = rng.integers(4, h - 5, size=n).astype(np.int32)
        for thr in (-2 ** 31, 0, 1 << 15, 1 << 24, 2 ** 31 - 1):
            got = np.zeros(n, np.uint8)
            dah.dah_harris(_ptr(img), w, xs.ctypes.data_as(C.POINTER(C.c_int)), ys.ctypes.data_as(C.POINTER(C.c_int)), n,
                           thr, _ptr(got))
            want = np.array([oracle.harris_score_sobel(img, int(x), int(y), thr) for x, y in zip(xs, ys)], np.uint8)
            assert np.array_equal(got, want), (name, thr)
