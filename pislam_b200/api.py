"""Host-side mirror of the reference's `pislam::` template API over the C ABI (include/pislam_b200.h).

The reference is a header-only C++ template library (reference include/Fast.h, Harris.h, Orb.h, Brief.h,
Util.h); its template arguments (vstep, border, logBucketSize, bucketLimit, words) become keyword
arguments here, `vstep` being the row length of the numpy image.  Every call goes through
libpislam_b200.so into the sm_100a kernels -- there is no CPU implementation behind this module, and
creating a Context without a usable B200 raises.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

PISLAM_OK = 0
PISLAM_ERR_INVALID_ARG = -10001
PISLAM_ERR_UNSUPPORTED = -10002
PISLAM_ERR_CAPACITY = -10003
PISLAM_ERR_NO_DEVICE = -10004
PISLAM_ERR_PEER_TIMEOUT = -10005
PISLAM_MAX_LEVELS = 16
REC_HEADER_WORDS = 32

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_intp = C.POINTER(C.c_int)


class PyramidDesc(C.Structure):
    """struct pislam_pyramid_desc"""
    _fields_ = [("vstep", C.c_int), ("nlevels", C.c_int), ("widths", C.c_int * PISLAM_MAX_LEVELS),
                ("heights", C.c_int * PISLAM_MAX_LEVELS), ("row_offsets", C.c_int * PISLAM_MAX_LEVELS),
                ("border", C.c_int), ("fast_threshold", C.c_int), ("harris_threshold", C.c_int32),
                ("log_bucket_size", C.c_int), ("bucket_limit", C.c_int), ("words", C.c_int)]


class PislamError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("pislam_b200: %s (status %d)" % (msg, status))
        self.status = status


_lib = None

# every symbol include/pislam_b200.h declares: (restype, argtypes)
_SIGNATURES = {
    "pislam_b200_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_size_t, C.c_int]),
    "pislam_b200_destroy": (C.c_int, [C.c_void_p]),
    "pislam_b200_error_string": (C.c_char_p, [C.c_int]),
    "pislam_b200_stream": (C.c_void_p, [C.c_void_p]),
    "pislam_b200_synchronize": (C.c_int, [C.c_void_p]),
    "pislam_b200_launch_count": (C.c_long, [C.c_void_p]),
    "pislam_b200_geometry_uploads": (C.c_long, [C.c_void_p]),
    "pislam_b200_set_tma": (C.c_int, [C.c_void_p, C.c_int]),
    "pislam_b200_set_sub_batch": (C.c_int, [C.c_void_p, C.c_int]),
    "pislam_b200_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "pislam_b200_kernel_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_long)]),
    "pislam_b200_build_pyramid": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_int, C.c_int,
                                            C.POINTER(PyramidDesc), C.c_void_p, C.c_size_t]),
    "pislam_b200_extract_frames": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_int, C.c_int,
                                             C.POINTER(PyramidDesc), C.c_int, C.c_void_p]),
    "pislam_b200_fast_detect": (C.c_int, [C.c_void_p, _u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "pislam_b200_fast_score_harris": (C.c_int, [C.c_void_p, _u8p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int,
                                                C.c_int32]),
    "pislam_b200_fast_extract": (C.c_int, [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           _u32p, C.c_int, _intp]),
    "pislam_b200_harris_score_sobel": (C.c_int, [C.c_void_p, _u8p, C.c_int, C.c_int, _intp, _intp, C.c_int,
                                                 C.c_int32, _u8p]),
    "pislam_b200_orb_centroids": (C.c_int, [C.c_void_p, _u8p, C.c_int, C.c_int, _u32p, C.c_int, _i32p]),
    "pislam_b200_atan2": (C.c_int, [C.c_void_p, _i32p, C.c_int, _u8p]),
    "pislam_b200_brief_describe": (C.c_int, [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_int, _u32p]),
    "pislam_b200_orb_compute": (C.c_int, [C.c_void_p, _u8p, C.c_int, C.c_int, _u32p, C.c_int, C.c_int, _u32p,
                                          _u8p]),
    "pislam_b200_gaussian5x5": (C.c_int, [C.c_void_p, _u8p, _u8p, C.c_int, C.c_int, C.c_int]),
    "pislam_b200_bilinear7_8": (C.c_int, [C.c_void_p, _u8p, _u8p, C.c_int, C.c_int, C.c_int]),
    "pislam_b200_bilinear13_16": (C.c_int, [C.c_void_p, _u8p, _u8p, C.c_int, C.c_int, C.c_int]),
    "pislam_b200_record_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "pislam_b200_record_kp_offset": (C.c_size_t, [C.c_int, C.c_int]),
    "pislam_b200_record_desc_offset": (C.c_size_t, [C.c_int, C.c_int]),
    "pislam_b200_record_level_offset": (C.c_size_t, [C.c_int, C.c_int]),
    "pislam_b200_extract_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(PyramidDesc),
                                            C.c_int, C.c_void_p]),
    "pislam_b200_extract_batch_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int,
                                                  C.POINTER(PyramidDesc), C.c_int, C.c_void_p]),
    "pislam_b200_extract_frames_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_int, C.c_int,
                                                   C.POINTER(PyramidDesc), C.c_int, C.c_void_p]),
    "pislam_b200_extract_pyramid": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(PyramidDesc), _u32p, _u8p, _u32p,
                                              C.c_int, _intp]),
    "pislam_b200_comm_unique_id": (C.c_int, [C.c_void_p]),
    "pislam_b200_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "pislam_b200_comm_destroy": (C.c_int, [C.c_void_p]),
    "pislam_b200_comm_register": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "pislam_b200_comm_p2p_enabled": (C.c_int, [C.c_void_p]),
    "pislam_b200_set_exchange_ctas": (C.c_int, [C.c_void_p, C.c_int]),
    "pislam_b200_extract_allgather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                                                C.POINTER(PyramidDesc), C.c_int, C.c_void_p]),
    "pislam_b200_extract_frames_allgather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_int, C.c_int,
                                                       C.c_int, C.POINTER(PyramidDesc), C.c_int, C.c_void_p]),
    "pislam_b200_allgather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "pislam_b200_join": (C.c_int, [C.c_void_p]),
    "pislam_b200_match_hamming": (C.c_int, [C.c_void_p, _u32p, _u32p, _u8p, C.c_int, _u32p, _u32p, _u8p, C.c_int,
                                            C.c_int, C.c_int, _i32p]),
    "pislam_b200_match_records": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, _i32p, C.c_int,
                                            C.c_int, C.c_void_p]),
    "pislam_b200_brief_table": (C.c_int, [C.POINTER(C.c_int8)]),
}


def load_library(rebuild_if_stale=True):
    """dlopen libpislam_b200.so (building it in-tree first when sources are newer)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    override = os.environ.get("PISLAM_B200_LIB")  # A/B runs of two builds of the library
    if override:
        rebuild_if_stale = False
        path = override
    if rebuild_if_stale and (not os.path.exists(path) or (_build.stale() and _have_nvcc())):
        path = _build.build_library()
    lib = C.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        if override and not hasattr(lib, name):
            continue  # an A/B build of an older revision (tools/gpu_front_ab.py) may lack the newest entry points
        f = getattr(lib, name)  # AttributeError here == a declared symbol is missing
        f.restype = res
        f.argtypes = args
    _lib = lib
    return lib


def _have_nvcc():
    try:
        _build._nvcc()
        return True
    except RuntimeError:
        return False


def _p(a, t):
    return a.ctypes.data_as(t)


def _check_img(a, name="img"):
    if not (isinstance(a, np.ndarray) and a.dtype == np.uint8 and a.ndim == 2 and a.flags.c_contiguous):
        raise TypeError("%s must be a C-contiguous 2-D uint8 array [rows][vstep]" % name)


def record_layout(cap, words):
    lib = load_library()
    return dict(bytes=lib.pislam_b200_record_bytes(cap, words), kp=lib.pislam_b200_record_kp_offset(cap, words),
                desc=lib.pislam_b200_record_desc_offset(cap, words),
                level=lib.pislam_b200_record_level_offset(cap, words))


def make_desc(layout, border=16, fast_threshold=20, harris_threshold=1 << 15, log_bucket=0, bucket_limit=5,
              words=8):
    """PyramidDesc from a pyramid.PyramidLayout and the demo's parameters (demo/demo.cpp:85-101)."""
    d = PyramidDesc()
    d.vstep, d.nlevels = layout.vstep, layout.nlevels
    for i in range(layout.nlevels):
        d.widths[i], d.heights[i], d.row_offsets[i] = layout.widths[i], layout.heights[i], layout.row_offsets[i]
    d.border, d.fast_threshold, d.harris_threshold = border, fast_threshold, harris_threshold
    d.log_bucket_size, d.bucket_limit, d.words = log_bucket, bucket_limit, words
    return d


def parse_records(buf, nframes, cap, words):
    """Splits a host copy of `nframes` frame records into [(kp, level, desc, level_start, n_total), ...]."""
    lay = record_layout(cap, words)
    buf = np.ascontiguousarray(buf).view(np.uint8).reshape(nframes, lay["bytes"])
    res = []
    for f in range(nframes):
        r = buf[f]
        hdr = r[:REC_HEADER_WORDS * 4].view(np.uint32)
        n = int(hdr[0])
        kp = r[lay["kp"]:lay["kp"] + 4 * n].view(np.uint32).copy()
        lv = r[lay["level"]:lay["level"] + n].copy()
        desc = (r[lay["desc"]:lay["desc"] + 4 * n * words].view(np.uint32).reshape(n, words).copy()
                if words else np.zeros((n, 0), np.uint32))
        res.append((kp, lv, desc, hdr[2:2 + PISLAM_MAX_LEVELS + 1].copy(), int(hdr[1])))
    return res


class Context:
    """One pislam_ctx: one CUDA stream on one GPU.  Method names follow the reference's functions."""

    def __init__(self, device=0, max_stack_bytes=0, max_keypoints=0):
        self.lib = load_library()
        h = C.c_void_p()
        st = self.lib.pislam_b200_create(C.byref(h), device, max_stack_bytes, max_keypoints)
        self._h = None
        self._ok(st)
        self._h = h

    def close(self):
        if self._h is not None:
            self.lib.pislam_b200_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ok(self, st):
        if st != PISLAM_OK:
            raise PislamError(st, self.lib.pislam_b200_error_string(st).decode())

    @property
    def stream(self):
        return self.lib.pislam_b200_stream(self._h)

    @property
    def launch_count(self):
        return self.lib.pislam_b200_launch_count(self._h)

    @property
    def geometry_uploads(self):
        return self.lib.pislam_b200_geometry_uploads(self._h)

    def synchronize(self):
        self._ok(self.lib.pislam_b200_synchronize(self._h))

    def set_profiling(self, enable):
        self._ok(self.lib.pislam_b200_set_profiling(self._h, int(bool(enable))))

    def kernel_times(self):
        """(ms[5], count[5]) for front_tile_kernel, order_kernel, orb_kernel, pyramid_build_kernel and the record
        exchange since the last call."""
        ms = (C.c_double * 5)()
        cnt = (C.c_long * 5)()
        self._ok(self.lib.pislam_b200_kernel_times(self._h, ms, cnt))
        return list(ms), list(cnt)

    def set_sub_batch(self, frames):
        self._ok(self.lib.pislam_b200_set_sub_batch(self._h, int(frames)))

    def set_tma(self, enable):
        self._ok(self.lib.pislam_b200_set_tma(self._h, int(bool(enable))))

    # ---- literal stages: same argument meaning as the reference templates ---------------------------
    def fastDetect(self, width, height, img, out, threshold, border=16, row0=0):
        """pislam::fastDetect<vstep,border>(width,height,img,out,threshold) -- Fast.h:54-158."""
        _check_img(img)
        _check_img(out, "out")
        self._ok(self.lib.pislam_b200_fast_detect(self._h, _p(img[row0:], _u8p), _p(out[row0:], _u8p), img.shape[1],
                                                  border, width, height, threshold))

    def fastScoreHarris(self, width, height, img, threshold, out, border=16, row0=0):
        """pislam::fastScoreHarris<vstep,border>(width,height,img,threshold,out) -- Fast.h:166-180."""
        _check_img(img)
        _check_img(out, "out")
        self._ok(self.lib.pislam_b200_fast_score_harris(self._h, _p(img[row0:], _u8p), _p(out[row0:], _u8p),
                                                        img.shape[1], border, width, height, threshold))

    def fastExtract(self, width, height, out, results=None, border=16, logBucketSize=0, bucketLimit=5, row0=0):
        """pislam::fastExtract<vstep,border,logBucketSize,bucketLimit>(width,height,out,results) --
        Fast.h:196-355.  Appends to `results` (a list) when given and returns the new keypoints."""
        _check_img(out, "out")
        cap = max(16, ((max(width - 2 * border, 0) + 1) // 2) * ((max(height - 2 * border, 0) + 1) // 2))
        kp = np.zeros(cap, np.uint32)
        n = C.c_int(0)
        self._ok(self.lib.pislam_b200_fast_extract(self._h, _p(out[row0:], _u8p), out.shape[1], border, width, height,
                                                   logBucketSize, bucketLimit, _p(kp, _u32p), cap, C.byref(n)))
        kp = kp[:n.value].copy()
        if results is not None:
            results.extend(int(v) for v in kp)
        return kp

    def harrisScoreSobel(self, img, xs, ys, threshold):
        """pislam::harrisScoreSobel<vstep>(img,x,y,threshold) for arrays of points -- Harris.h:80-248."""
        _check_img(img)
        xs = np.ascontiguousarray(xs, np.int32)
        ys = np.ascontiguousarray(ys, np.int32)
        s = np.zeros(len(xs), np.uint8)
        self._ok(self.lib.pislam_b200_harris_score_sobel(self._h, _p(img, _u8p), img.shape[1], img.shape[0],
                                                         _p(xs, _intp), _p(ys, _intp), len(xs), threshold,
                                                         _p(s, _u8p)))
        return s

    def orbCentroids(self, img, points):
        """pislam::orbCentroids<vstep>(img,points) -- Orb.h:80-308 (4+4 interleaved layout)."""
        _check_img(img)
        kp = np.ascontiguousarray(points, np.uint32)
        c = np.zeros(max(8, (2 * len(kp) + 7) & ~7), np.int32)
        self._ok(self.lib.pislam_b200_orb_centroids(self._h, _p(img, _u8p), img.shape[1], img.shape[0],
                                                    _p(kp, _u32p), len(kp), _p(c, _i32p)))
        return c[:(2 * len(kp) + 7) & ~7]

    def atan2(self, xys):
        """pislam::atan2(xys) -- Orb.h:310-387."""
        xys = np.ascontiguousarray(xys, np.int32)
        a = np.zeros(max(4, len(xys) // 2), np.uint8)
        self._ok(self.lib.pislam_b200_atan2(self._h, _p(xys, _i32p), len(xys), _p(a, _u8p)))
        return a[:len(xys) // 2]

    def briefDescribe(self, img, x, y, rot, words=8, fill=0xdeadbeef):
        """pislam::briefDescribe<vstep,words>(img,x,y,rot,descriptor) -- Brief.h:637-733."""
        _check_img(img)
        d = np.full(words, fill, np.uint32)
        self._ok(self.lib.pislam_b200_brief_describe(self._h, _p(img, _u8p), img.shape[1], img.shape[0], x, y, rot,
                                                     words, _p(d, _u32p)))
        return d

    def orbCompute(self, img, points, descriptors=None, words=8, want_angles=False):
        """pislam::orbCompute<vstep,words>(img,points,descriptors) -- Orb.h:396-441."""
        _check_img(img)
        kp = np.ascontiguousarray(points, np.uint32)
        d = np.zeros((max(1, len(kp)), words), np.uint32)
        a = np.zeros(max(1, len(kp)), np.uint8)
        self._ok(self.lib.pislam_b200_orb_compute(self._h, _p(img, _u8p), img.shape[1], img.shape[0], _p(kp, _u32p),
                                                  len(kp), words, _p(d, _u32p), _p(a, _u8p)))
        d = d[:len(kp)]
        if descriptors is not None:
            descriptors.extend(int(v) for v in d.ravel())
        return (d, a[:len(kp)]) if want_angles else d

    # ---- pyramid-preparation helpers (not on the hot path) ---------------------------------------------
    def gaussian5x5(self, width, height, img, out):
        """pislam::gaussian5x5<vstep>(width,height,img,out) -- Gaussian.h:48-72."""
        _check_img(img)
        _check_img(out, "out")
        self._ok(self.lib.pislam_b200_gaussian5x5(self._h, _p(img, _u8p), _p(out, _u8p), img.shape[1], width, height))

    def bilinear7_8(self, width, height, img, out):
        """pislam::bilinear7_8<vstep>(width,height,img,out) -- Bilinear.h:42-150."""
        _check_img(img)
        _check_img(out, "out")
        self._ok(self.lib.pislam_b200_bilinear7_8(self._h, _p(img, _u8p), _p(out, _u8p), img.shape[1], width, height))

    def bilinear13_16(self, width, height, img, out):
        """pislam::bilinear13_16<vstep>(width,height,img,out) -- Bilinear.h:165-321."""
        _check_img(img)
        _check_img(out, "out")
        self._ok(self.lib.pislam_b200_bilinear13_16(self._h, _p(img, _u8p), _p(out, _u8p), img.shape[1], width, height))

    # ---- descriptor matching (not in the reference, definition in include/pislam_b200.h) ----
    def match_hamming(self, q_desc, t_desc, q_kp=None, q_level=None, t_kp=None, t_level=None, radius=0):
        """Best / second-best Hamming match of every query descriptor -> int32 [nq][3] = (index, dist, second)."""
        q_desc, t_desc = np.ascontiguousarray(q_desc, np.uint32), np.ascontiguousarray(t_desc, np.uint32)
        words = next((int(a.size // len(a)) for a in (q_desc, t_desc) if len(a)), 8)
        q_desc, t_desc = q_desc.reshape(len(q_desc), words), t_desc.reshape(len(t_desc), words)
        out = np.zeros((len(q_desc), 3), np.int32)
        arrs = []
        for a, t in ((q_kp, np.uint32), (q_level, np.uint8), (t_kp, np.uint32), (t_level, np.uint8)):
            arrs.append(None if a is None else np.ascontiguousarray(a, t))
        ptr = lambda a, t: None if a is None or a.size == 0 else _p(a, t)  # noqa: E731
        self._ok(self.lib.pislam_b200_match_hamming(
            self._h, ptr(q_desc, _u32p), ptr(arrs[0], _u32p), ptr(arrs[1], _u8p), len(q_desc),
            ptr(t_desc, _u32p), ptr(arrs[2], _u32p), ptr(arrs[3], _u8p), len(t_desc), words, radius,
            _p(out, _i32p) if out.size else None))
        return out

    def match_records(self, records, nframes, cap, words, pairs, radius=0):
        """Matches for frame pairs of a host batch of records -> int32 [npairs][cap][3] (rows >= n stay -2)."""
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        out = np.full((len(pairs), cap, 3), -2, np.int32)
        self._ok(self.lib.pislam_b200_match_records(self._h, records.ctypes.data, nframes, cap, words,
                                                    _p(pairs, _i32p) if pairs.size else None, len(pairs), radius,
                                                    out.ctypes.data))
        return out

    # ---- fused whole-pyramid path (the caller loop of demo/demo.cpp:77-101) -------------------------
    def extract_pyramid(self, stack, desc, cap=4096):
        """One stacked pyramid (host array) -> (kp, level, descriptors)."""
        _check_img(stack, "stack")
        rows = max(desc.row_offsets[i] + desc.heights[i] for i in range(desc.nlevels))
        if stack.shape[1] != desc.vstep or stack.shape[0] < rows:
            raise ValueError("stack shape %s does not hold the described pyramid (%d rows x %d)" % (
                stack.shape, rows, desc.vstep))
        kp = np.zeros(cap, np.uint32)
        lv = np.zeros(cap, np.uint8)
        d = np.zeros((cap, max(1, desc.words)), np.uint32)
        n = C.c_int(0)
        st = self.lib.pislam_b200_extract_pyramid(self._h, stack.ctypes.data, C.byref(desc), _p(kp, _u32p),
                                                  _p(lv, _u8p), _p(d, _u32p), cap, C.byref(n))
        if st == PISLAM_ERR_CAPACITY:
            raise PislamError(st, "capacity %d exceeded: %d keypoints found" % (cap, n.value))
        self._ok(st)
        n = n.value
        return kp[:n].copy(), lv[:n].copy(), d[:n, :desc.words].copy()

    def extract_batch(self, stacks, desc, cap=4096):
        """stacks: host uint8 [nframes][rows][vstep] -> list of (kp, level, desc, level_start, n_total)."""
        assert stacks.dtype == np.uint8 and stacks.ndim == 3 and stacks.flags.c_contiguous
        rows = max(desc.row_offsets[i] + desc.heights[i] for i in range(desc.nlevels))
        if stacks.shape[2] != desc.vstep or stacks.shape[1] < rows:
            raise ValueError("stacks shape %s does not hold the described pyramid (%d rows x %d)" % (
                stacks.shape, rows, desc.vstep))
        nf = stacks.shape[0]
        lay = record_layout(cap, desc.words)
        rec = np.zeros(nf * lay["bytes"], np.uint8)
        st = self.lib.pislam_b200_extract_batch(self._h, stacks.ctypes.data, stacks.shape[1] * stacks.shape[2], nf,
                                                C.byref(desc), cap, rec.ctypes.data)
        if st != PISLAM_ERR_CAPACITY:
            self._ok(st)
        return parse_records(rec, nf, cap, desc.words)

    # ---- pyramid builder in front of the path (include/pislam_b200.h) -----------------
    def build_pyramid(self, frames, desc, blur=False, fill=0):
        """frames: host uint8 [nframes][H][W] base frames -> stacks uint8 [nframes][rows][vstep]."""
        frames = np.ascontiguousarray(frames)
        assert frames.dtype == np.uint8 and frames.ndim == 3
        nf, H, W = frames.shape
        assert (W, H) == (desc.widths[0], desc.heights[0]), "base frames must have the size of level 0"
        rows = max(desc.row_offsets[i] + desc.heights[i] for i in range(desc.nlevels))
        stacks = np.full((nf, rows, desc.vstep), fill, np.uint8)
        self._ok(self.lib.pislam_b200_build_pyramid(self._h, frames.ctypes.data, W, W * H, nf, int(bool(blur)),
                                                    C.byref(desc), stacks.ctypes.data, rows * desc.vstep))
        return stacks

    def extract_frames(self, frames, desc, cap=4096, blur=False):
        """frames: host uint8 [nframes][H][W] base frames -> list of (kp, level, desc, level_start, n_total)."""
        frames = np.ascontiguousarray(frames)
        assert frames.dtype == np.uint8 and frames.ndim == 3
        nf, H, W = frames.shape
        assert (W, H) == (desc.widths[0], desc.heights[0]), "base frames must have the size of level 0"
        lay = record_layout(cap, desc.words)
        rec = np.zeros(nf * lay["bytes"], np.uint8)
        st = self.lib.pislam_b200_extract_frames(self._h, frames.ctypes.data, W, W * H, nf, int(bool(blur)),
                                                 C.byref(desc), cap, rec.ctypes.data)
        if st != PISLAM_ERR_CAPACITY:
            self._ok(st)
        return parse_records(rec, nf, cap, desc.words)

    def extract_frames_ptr(self, frames_ptr, frame_pitch, frame_stride, nframes, desc, cap, records_ptr, blur=False,
                           asynchronous=False):
        """Raw-pointer form (host or device memory); asynchronous=True only enqueues (see synchronize())."""
        f = self.lib.pislam_b200_extract_frames_async if asynchronous else self.lib.pislam_b200_extract_frames
        return f(self._h, frames_ptr, frame_pitch, frame_stride, nframes, int(bool(blur)), C.byref(desc), cap,
                 records_ptr)

    def extract_batch_ptr(self, stacks_ptr, frame_stride, nframes, desc, cap, records_ptr, asynchronous=False):
        """Raw-pointer form (host or device memory, e.g. torch tensors' data_ptr())."""
        f = self.lib.pislam_b200_extract_batch_async if asynchronous else self.lib.pislam_b200_extract_batch
        return f(self._h, stacks_ptr, frame_stride, nframes, C.byref(desc), cap, records_ptr)

    # ---- multi-GPU ---------------------------------------------------------------------------------------
    def comm_init(self, nranks, rank, unique_id):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._ok(self.lib.pislam_b200_comm_init(self._h, nranks, rank, buf))

    def comm_register(self, ptr, nbytes):
        """Collective: maps the device buffer into every peer for the size-aware exchange.  Returns True when the
        buffer will use it, False when every rank falls back to ncclAllGather for it."""
        st = self.lib.pislam_b200_comm_register(self._h, ptr, nbytes)
        if st == PISLAM_ERR_UNSUPPORTED:
            return False
        self._ok(st)
        return True

    def set_exchange_ctas(self, ctas):
        self._ok(self.lib.pislam_b200_set_exchange_ctas(self._h, int(ctas)))

    @property
    def p2p_windows(self):
        return self.lib.pislam_b200_comm_p2p_enabled(self._h)

    def comm_destroy(self):
        self._ok(self.lib.pislam_b200_comm_destroy(self._h))

    def join(self):
        self._ok(self.lib.pislam_b200_join(self._h))

    def allgather_ptr(self, all_ptr, bytes_per_rank):
        return self.lib.pislam_b200_allgather(self._h, all_ptr, bytes_per_rank)

    def extract_allgather_ptr(self, stacks_ptr, frame_stride, nframes_local, desc, cap, all_records_ptr,
                              slots_per_rank=None):
        """`slots_per_rank` (same on every rank, default nframes_local) sizes the rank's slot of the gather buffer;
        a rank may own fewer frames than slots, or none (sharding.shard_frames)."""
        slots = nframes_local if slots_per_rank is None else slots_per_rank
        return self.lib.pislam_b200_extract_allgather(self._h, stacks_ptr, frame_stride, nframes_local, slots,
                                                      C.byref(desc), cap, all_records_ptr)


def _frames_allgather(self, frames_ptr, frame_pitch, frame_stride, nframes_local, desc, cap, all_records_ptr,
                      slots_per_rank=None, blur=False):
    """pislam_b200_extract_frames_allgather: this rank's BASE frames (host or device) -> its slot, then the exchange."""
    slots = nframes_local if slots_per_rank is None else slots_per_rank
    return self.lib.pislam_b200_extract_frames_allgather(self._h, frames_ptr, frame_pitch, frame_stride, nframes_local,
                                                         int(bool(blur)), slots, C.byref(desc), cap, all_records_ptr)


Context.extract_frames_allgather_ptr = _frames_allgather


def comm_unique_id():
    lib = load_library()
    buf = C.create_string_buffer(128)
    st = lib.pislam_b200_comm_unique_id(buf)
    if st != PISLAM_OK:
        raise PislamError(st, lib.pislam_b200_error_string(st).decode())
    return bytes(buf.raw)


def brief_table():
    """The rotated + clamped rBRIEF pattern the device uses, int8[30][256][4] (Brief.h:28-50)."""
    lib = load_library()
    t = np.zeros((30, 256, 4), np.int8)
    st = lib.pislam_b200_brief_table(t.ctypes.data_as(C.POINTER(C.c_int8)))
    if st != PISLAM_OK:
        raise PislamError(st, "brief_table")
    return t


# Util.h:27-45
def encodeFast(score, x, y):
    return ((score & 0xFF) << 24) | ((x & 0xFFF) << 12) | (y & 0xFFF)


def rencodeFastScore(score, encoded):
    return ((score & 0xFF) << 24) | (encoded & 0xFFFFFF)


def decodeFastX(encoded):
    return (encoded >> 12) & 0xFFF


def decodeFastY(encoded):
    return encoded & 0xFFF


def decodeFastScore(encoded):
    return (encoded >> 24) & 0xFF
