"""ctypes bindings of the CPU oracle (oracle/libpislam_oracle.so) and, when it has been built,
of the reference's own headers compiled through the NEON shim (oracle/_ref/libpislam_ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
The product package (pislam_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libpislam_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libpislam_ref.so")

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
intp = C.POINTER(C.c_int)


def _p(a, t):
    return a.ctypes.data_as(t)


def build_oracle(force=False):
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
            os.path.join(ORACLE_DIR, "pislam_oracle.c")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


class _Lib:
    """Common numpy-level API over either library (`po_*` or `ref_*` symbols)."""

    def __init__(self, path, prefix):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        L, p = self.lib, prefix
        f = getattr(L, p + "fast_detect")
        f.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        f = getattr(L, p + "fast_score_harris")
        f.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int32]
        f = getattr(L, p + "harris_score_sobel")
        f.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int32]
        f.restype = C.c_int
        f = getattr(L, p + "fast_extract")
        f.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u32p, C.c_int]
        f.restype = C.c_int
        f = getattr(L, p + "orb_centroids")
        f.argtypes = [u8p, C.c_int, u32p, C.c_int, i32p]
        f.restype = C.c_int
        f = getattr(L, p + "atan2")
        f.argtypes = [i32p, C.c_int, u8p]
        f.restype = C.c_int
        f = getattr(L, p + "orb_compute")
        f.argtypes = [u8p, C.c_int, u32p, C.c_int, C.c_int, u32p]
        f = getattr(L, p + "brief_describe")
        f.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u32p]
        for name in ("bilinear7_8", "bilinear13_16"):
            f = getattr(L, p + name)
            f.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int]
        for name in ("encode_fast",):
            f = getattr(L, p + name)
            f.argtypes = [C.c_uint32] * 3
            f.restype = C.c_uint32
        f = getattr(L, p + "rencode_fast_score")
        f.argtypes = [C.c_uint32] * 2
        f.restype = C.c_uint32
        for name in ("decode_fast_x", "decode_fast_y", "decode_fast_score"):
            f = getattr(L, p + name)
            f.argtypes = [C.c_uint32]
            f.restype = C.c_uint32

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    # img / out: 2-D uint8 C-contiguous arrays whose row length is vstep; `row0` selects the level
    def fast_detect(self, img, out, border, width, height, threshold, row0=0):
        vstep = img.shape[1]
        r = self._f("fast_detect")(_p(img[row0:], u8p), _p(out[row0:], u8p), vstep, border, width, height,
                                   threshold)
        assert r in (0, None) or self.prefix == "po_", "reference driver has no such instantiation"

    def fast_score_harris(self, img, out, border, width, height, threshold, row0=0):
        vstep = img.shape[1]
        self._f("fast_score_harris")(_p(img[row0:], u8p), _p(out[row0:], u8p), vstep, border, width, height,
                                     threshold)

    def harris_score_sobel(self, img, x, y, threshold):
        return int(self._f("harris_score_sobel")(_p(img, u8p), img.shape[1], x, y, threshold)) & 0xff

    def fast_extract(self, out, border, width, height, log_bucket=0, bucket_limit=5, row0=0, cap=None):
        vstep = out.shape[1]
        cap = cap if cap is not None else max(16, (width * height) // 4 + 16)
        kp = np.zeros(cap, np.uint32)
        n = self._f("fast_extract")(_p(out[row0:], u8p), vstep, border, width, height, log_bucket,
                                    bucket_limit, _p(kp, u32p), cap)
        assert n >= 0, n
        return kp[:n].copy()

    def orb_centroids(self, img, kp):
        kp = np.ascontiguousarray(kp, np.uint32)
        c = np.zeros(((2 * len(kp) + 7) & ~7) + 8, np.int32)
        n = self._f("orb_centroids")(_p(img, u8p), img.shape[1], _p(kp, u32p), len(kp), _p(c, i32p))
        assert n >= 0
        return c[:n].copy()

    def atan2(self, xys):
        xys = np.ascontiguousarray(xys, np.int32)
        a = np.zeros(len(xys) // 2 + 8, np.uint8)
        n = self._f("atan2")(_p(xys, i32p), len(xys), _p(a, u8p))
        return a[:n].copy()

    def orb_compute(self, img, kp, words=8):
        kp = np.ascontiguousarray(kp, np.uint32)
        d = np.zeros(max(1, len(kp) * words), np.uint32)
        self._f("orb_compute")(_p(img, u8p), img.shape[1], _p(kp, u32p), len(kp), words, _p(d, u32p))
        return d[:len(kp) * words].reshape(len(kp), words).copy()

    def bilinear7_8(self, img, out, width, height):
        self._f("bilinear7_8")(_p(img, u8p), _p(out, u8p), img.shape[1], width, height)

    def bilinear13_16(self, img, out, width, height):
        self._f("bilinear13_16")(_p(img, u8p), _p(out, u8p), img.shape[1], width, height)

    def brief_describe(self, img, x, y, rot, words=8, fill=0xdeadbeef):
        d = np.full(words, fill, np.uint32)
        self._f("brief_describe")(_p(img, u8p), img.shape[1], x, y, rot, words, _p(d, u32p))
        return d


class Oracle(_Lib):
    def __init__(self):
        super().__init__(build_oracle(), "po_")
        L = self.lib
        L.po_angle_bin.argtypes = [C.c_int32, C.c_int32]
        L.po_angle_bin.restype = C.c_uint8
        L.po_recpe.argtypes = [C.c_float]
        L.po_recpe.restype = C.c_float
        L.po_brief_table.restype = C.POINTER(C.c_int8)
        L.po_extract_pyramid.argtypes = [u8p, C.c_int, C.c_int, intp, intp, intp, C.c_int, C.c_int, C.c_int32,
                                         C.c_int, C.c_int, C.c_int, C.c_int, u8p, u32p, u8p, u32p, C.c_int]
        L.po_extract_pyramid.restype = C.c_int
        L.po_extract_batch.argtypes = [u8p, C.c_long, C.c_int, C.c_int, C.c_int, intp, intp, intp, C.c_int,
                                       C.c_int, C.c_int32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, intp,
                                       u32p, u32p]
        L.po_extract_batch.restype = C.c_int

    def gaussian5x5(self, img, out, width, height):
        self.lib.po_gaussian5x5.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int]
        self.lib.po_gaussian5x5(_p(img, u8p), _p(out, u8p), img.shape[1], width, height)

    def resize_bilinear(self, src, w, h):
        """src: 2-D uint8 (any pitch) -> h x w uint8 (the pyramid builder's resize, oracle/pislam_oracle.h)."""
        self.lib.po_resize_bilinear.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int, C.c_int]
        src = np.ascontiguousarray(src)
        dst = np.zeros((h, w), np.uint8)
        self.lib.po_resize_bilinear(_p(src, u8p), src.shape[1], src.shape[1], src.shape[0], _p(dst, u8p), w, w, h)
        return dst

    def build_pyramid(self, base, layout, blur=False):
        """base: H x W uint8 -> stacked pyramid [layout.rows][layout.vstep] (levels resized from the base)."""
        self.lib.po_build_pyramid.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, intp, intp, intp, u8p]
        base = np.ascontiguousarray(base)
        assert base.shape == (layout.heights[0], layout.widths[0])
        w = np.asarray(layout.widths, np.int32)
        h = np.asarray(layout.heights, np.int32)
        r = np.asarray(layout.row_offsets, np.int32)
        stack = np.full((layout.rows, layout.vstep), 0xEE, np.uint8)
        self.lib.po_build_pyramid(_p(base, u8p), base.shape[1], int(bool(blur)), layout.vstep, layout.nlevels,
                                  _p(w, intp), _p(h, intp), _p(r, intp), _p(stack, u8p))
        return stack

    def angle_bin(self, m10, m01):
        return int(self.lib.po_angle_bin(int(m10), int(m01)))

    def recpe(self, x):
        return float(self.lib.po_recpe(float(x)))

    def brief_table(self):
        p = self.lib.po_brief_table()
        return np.ctypeslib.as_array(p, shape=(30, 256, 4)).copy()

    def extract_pyramid(self, stack, widths, heights, row_offsets, border=16, fast_threshold=20,
                        harris_threshold=1 << 15, log_bucket=0, bucket_limit=5, words=8, stacked_y=False,
                        cap=1 << 16, want_out=False):
        """Returns (kp, kp_level, desc[, out]) -- the demo loop (demo/demo.cpp:77-101)."""
        nl = len(widths)
        w = np.asarray(widths, np.int32)
        h = np.asarray(heights, np.int32)
        r = np.asarray(row_offsets, np.int32)
        scratch = np.zeros_like(stack)
        kp = np.zeros(cap, np.uint32)
        lv = np.zeros(cap, np.uint8)
        desc = np.zeros(cap * words, np.uint32)
        # words == 0: keypoints only (no descriptor pass, so borders < 16 are fine)
        n = self.lib.po_extract_pyramid(_p(stack, u8p), stack.shape[1], nl, _p(w, intp), _p(h, intp), _p(r, intp),
                                        border, fast_threshold, harris_threshold, log_bucket, bucket_limit,
                                        words, int(stacked_y), _p(scratch, u8p), _p(kp, u32p), _p(lv, u8p),
                                        _p(desc, u32p) if words else None, cap)
        assert n >= 0, n
        res = (kp[:n].copy(), lv[:n].copy(), desc[:n * words].reshape(n, words).copy())
        return res + (scratch,) if want_out else res

    def extract_batch(self, stacks, widths, heights, row_offsets, nthreads, border=16, fast_threshold=20,
                      harris_threshold=1 << 15, log_bucket=0, bucket_limit=5, words=8, cap=1 << 14,
                      keep=False):
        """stacks: [nframes, rows, vstep] uint8.  Returns counts (and kp, desc when keep)."""
        nf = stacks.shape[0]
        w = np.asarray(widths, np.int32)
        h = np.asarray(heights, np.int32)
        r = np.asarray(row_offsets, np.int32)
        counts = np.zeros(nf, np.int32)
        kp = np.zeros((nf, cap), np.uint32) if keep else None
        desc = np.zeros((nf, cap * words), np.uint32) if keep else None
        st = self.lib.po_extract_batch(_p(stacks, u8p), stacks.shape[1] * stacks.shape[2], nf, stacks.shape[2],
                                       len(w), _p(w, intp), _p(h, intp), _p(r, intp), border, fast_threshold,
                                       harris_threshold, log_bucket, bucket_limit, words, cap, nthreads,
                                       _p(counts, intp), _p(kp, u32p) if keep else None,
                                       _p(desc, u32p) if keep else None)
        assert st == 0, st
        return (counts, kp, desc) if keep else counts


class Reference(_Lib):
    """The reference's own headers (only present where oracle/_ref has been built)."""

    def __init__(self):
        super().__init__(REF_SO, "ref_")
        f = self.lib.ref_extract_batch
        f.argtypes = [u8p, C.c_long, C.c_int, C.c_int, C.c_int, intp, intp, intp, C.c_int, C.c_int32, C.c_int,
                      C.c_int, C.c_int, C.c_int, intp, u32p, u32p]
        f.restype = C.c_int

    def extract_batch(self, stacks, widths, heights, row_offsets, nthreads, fast_threshold=20,
                      harris_threshold=1 << 15, log_bucket=0, bucket_limit=5, cap=1 << 14, keep=False):
        """The reference's caller loop (demo/demo.cpp:77-101) over [nframes, rows, vstep] stacks on `nthreads`
        std::threads, natively (oracle/ref_driver.cpp:ref_extract_batch; border 16, 8 words).  Returns counts
        (and kp [nf][cap], desc [nf][cap*8] when keep)."""
        nf = stacks.shape[0]
        w = np.asarray(widths, np.int32)
        h = np.asarray(heights, np.int32)
        r = np.asarray(row_offsets, np.int32)
        counts = np.zeros(nf, np.int32)
        kp = np.zeros((nf, cap), np.uint32) if keep else None
        desc = np.zeros((nf, cap * 8), np.uint32) if keep else None
        st = self.lib.ref_extract_batch(_p(stacks, u8p), stacks.shape[1] * stacks.shape[2], nf, stacks.shape[2],
                                        len(w), _p(w, intp), _p(h, intp), _p(r, intp), fast_threshold,
                                        harris_threshold, log_bucket, bucket_limit, cap, nthreads,
                                        _p(counts, intp), _p(kp, u32p) if keep else None,
                                        _p(desc, u32p) if keep else None)
        assert st == 0, "ref_extract_batch: %d" % st
        return (counts, kp, desc) if keep else counts


def have_reference():
    return os.path.exists(REF_SO)


def fnv1a64(a):
    """64-bit FNV-1a over the raw little-endian bytes of `a` (the hash SURVEY.md section 8c quotes)."""
    h = 1469598103934665603
    data = np.ascontiguousarray(a).view(np.uint8).ravel()
    # vectorised in chunks would change nothing: FNV is inherently serial; fine for <= a few MB
    for b in data.tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h
