"""the on-GPU pyramid builder in front of the hot path, and the fused
base-frame -> pyramid -> features call (pislam_b200_build_pyramid / pislam_b200_extract_frames).

PARITY UNPINNED BY THE REFERENCE: the reference ships no x1/1.2 resize and leaves pyramid construction to its caller
(README.md:28-31).  The definition is this repo's (oracle/pislam_oracle.h: po_resize_bilinear / po_build_pyramid,
pinned to pislam_b200/pyramid.py in tests/test_oracle.py); the CUDA path must equal it byte for byte, and the records
of extract_frames must equal the oracle's extraction from the oracle-built stack (VGA / 1080p / 4K)."""
import numpy as np
import pytest

from pislam_b200 import api
from pislam_b200.pyramid import FHD, UHD, VGA, PyramidLayout, pyramid_layout, synthetic_base

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def same(got, want, what):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, "%s: shape %s vs %s" % (what, got.shape, want.shape)
    d = np.flatnonzero(got.ravel() != want.ravel())
    assert len(d) == 0, "%s: %d diffs, first at %d: got %s want %s" % (what, len(d), d[0], got.ravel()[d[0]],
                                                                        want.ravel()[d[0]])


@pytest.mark.parametrize("name,lay", [("vga", VGA), ("fhd", FHD), ("uhd", UHD)])
def test_build_pyramid_baseline_layouts(ctx, oracle, name, lay):
    base = synthetic_base(lay.widths[0], lay.heights[0], 0x5EED0000 + 3)
    desc = api.make_desc(lay)
    for blur in (False, True):
        got = ctx.build_pyramid(base[None], desc, blur=blur, fill=0x77)[0]
        same(got, oracle.build_pyramid(base, lay, blur=blur), "%s blur=%s" % (name, blur))


def test_build_pyramid_ragged_and_batched(ctx, oracle):
    """Odd sizes, a pitch that is not a multiple of 4, up-scaling levels, 1-pixel levels, a stack with rows that
    belong to no level (must come back untouched), several frames per call."""
    rng = np.random.Generator(np.random.PCG64(9))
    cases = [
        (203, [(203, 117), (148, 85), (108, 62), (79, 45), (1, 1), (250 - 47, 5)], 3),
        (64, [(61, 47), (64, 50), (7, 3), (33, 47)], 0),
        (320, [(320, 240), (267, 200), (222, 167), (185, 139)], 0),
        (130, [(129, 9), (3, 3), (129, 9)], 7),
    ]
    for vstep, dims, gap in cases:
        offs, r = [], 0
        for w, h in dims:
            offs.append(r)
            r += h + gap
        lay = PyramidLayout(vstep, [w for w, h in dims], [h for w, h in dims], offs)
        nf = 3
        bases = rng.integers(0, 256, size=(nf, dims[0][1], dims[0][0]), dtype=np.uint8)
        desc = api.make_desc(lay)
        for blur in (False, True):
            got = ctx.build_pyramid(bases, desc, blur=blur, fill=0x5A)
            for f in range(nf):
                want = oracle.build_pyramid(bases[f], lay, blur=blur)
                covered = np.zeros(lay.rows, bool)
                for h, o in zip(lay.heights, lay.row_offsets):
                    covered[o:o + h] = True
                same(got[f][covered], want[covered], "vstep %d frame %d blur %s" % (vstep, f, blur))
                assert (got[f][~covered] == 0x5A).all(), "rows outside every level were written"


@pytest.mark.parametrize("name,lay,cap", [("vga", VGA, 4096), ("fhd", FHD, 8192), ("uhd", UHD, 16384)])
def test_extract_frames_matches_oracle(ctx, oracle, name, lay, cap):
    nf = 3 if name == "vga" else 1
    bases = np.stack([synthetic_base(lay.widths[0], lay.heights[0], 0x5EED0000 + 20 + f) for f in range(nf)])
    for blur, lb, bl in ((False, 4, 3), (True, 0, 5)):
        desc = api.make_desc(lay, log_bucket=lb, bucket_limit=bl)
        res = ctx.extract_frames(bases, desc, cap=cap, blur=blur)
        for f in range(nf):
            stack = oracle.build_pyramid(bases[f], lay, blur=blur)
            okp, olv, odesc = oracle.extract_pyramid(stack, lay.widths, lay.heights, lay.row_offsets, log_bucket=lb,
                                                     bucket_limit=bl)
            kp, lv, dsc, _, n_total = res[f]
            assert len(okp) > 200 and n_total == len(okp)
            same(kp, okp, "%s frame %d blur %s keypoints" % (name, f, blur))
            same(lv, olv, "levels")
            same(dsc, odesc, "descriptors")


def test_extract_frames_equals_extract_batch_and_device_pointers(ctx, oracle):
    """The base-frame call returns exactly the records of the stack call on the stacks the builder makes; host and
    device sources agree; chunking of device-resident frames (L2-sized sub-batches) does not show in the result."""
    import torch
    lay = pyramid_layout(320, 240, nlevels=5)
    nf = 70
    distinct = [synthetic_base(320, 240, 0x5EED0000 + 60 + i) for i in range(6)]
    bases = np.stack([distinct[i % 6] for i in range(nf)])
    desc = api.make_desc(lay, log_bucket=4, bucket_limit=3)
    cap = 2048
    from_frames = ctx.extract_frames(bases, desc, cap=cap)
    stacks = ctx.build_pyramid(bases, desc)
    from_stacks = ctx.extract_batch(stacks, desc, cap=cap)
    rec_bytes = api.record_layout(cap, 8)["bytes"]
    d_bases = torch.from_numpy(bases).cuda()
    d_rec = torch.zeros(nf * rec_bytes, dtype=torch.uint8, device="cuda")
    st = ctx.extract_frames_ptr(d_bases.data_ptr(), 320, 320 * 240, nf, desc, cap, d_rec.data_ptr())
    assert st == 0
    from_device = api.parse_records(d_rec.cpu().numpy(), nf, cap, 8)
    for f in range(nf):
        for other, what in ((from_stacks, "extract_batch"), (from_device, "device pointers")):
            same(from_frames[f][0], other[f][0], "frame %d keypoints vs %s" % (f, what))
            same(from_frames[f][2], other[f][2], "frame %d descriptors vs %s" % (f, what))
        same(from_frames[f][0], from_frames[f % 6][0], "frame %d vs its first copy" % f)
    for f in range(6):
        stack = oracle.build_pyramid(distinct[f], lay)
        okp, olv, odesc = oracle.extract_pyramid(stack, lay.widths, lay.heights, lay.row_offsets, log_bucket=4,
                                                 bucket_limit=3)
        same(from_frames[f][0], okp, "frame %d keypoints vs oracle" % f)
        same(from_frames[f][2], odesc, "frame %d descriptors vs oracle" % f)
    del d_bases, d_rec


def test_extract_frames_argument_errors(ctx):
    lay = pyramid_layout(64, 48, nlevels=2)
    desc = api.make_desc(lay)
    frames = np.zeros((1, 48, 64), np.uint8)
    rec = np.zeros(api.record_layout(64, 8)["bytes"], np.uint8)
    # pitch smaller than the width
    assert ctx.extract_frames_ptr(frames.ctypes.data, 32, 64 * 48, 1, desc, 64, rec.ctypes.data) == api.PISLAM_ERR_INVALID_ARG
    # two frames that would overlap
    assert ctx.extract_frames_ptr(frames.ctypes.data, 64, 100, 2, desc, 64, rec.ctypes.data) == api.PISLAM_ERR_INVALID_ARG
    assert ctx.extract_frames_ptr(frames.ctypes.data, 64, 64 * 48, 1, desc, 64, rec.ctypes.data) == 0
