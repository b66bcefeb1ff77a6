"""CPU checks of arithmetic the CUDA kernels rely on (numpy models of the device expressions; no GPU, no oracle).

* front_tile_kernel phase A uses a carry-tolerant SWAR compare (front_kernels.cuh: bytes_gt_loose) that must be a
  SUPERSET of the exact per-byte test |a-b| > t for every t < 128: phase B then decides exactly.
* the compass pre-test itself must never reject a FAST-9 corner (Fast.h:54-158 as restated by the oracle).
"""
import numpy as np

from pislam_b200.pyramid import noise_stack, pyramid_layout, synthetic_stack


def _loose(d_words, t):
    """bit 7 of every byte of d | (d + K), K = (127 - t) * 0x01010101, as uint32 arithmetic (device expression)."""
    K = np.uint64((127 - t) * 0x01010101)
    d = d_words.astype(np.uint64)
    v = (d | ((d + K) & np.uint64(0xffffffff)))
    return np.stack([(v >> np.uint64(8 * k + 7)) & np.uint64(1) for k in range(4)], -1).astype(bool)


def test_loose_compare_is_a_superset_with_tie_only_false_positives():
    # every (byte below, byte) pair in the two low bytes and in the two high bytes; the other bytes random
    lo, hi = np.meshgrid(np.arange(256, dtype=np.uint64), np.arange(256, dtype=np.uint64), indexing="ij")
    rng = np.random.Generator(np.random.PCG64(5))
    for t in list(range(0, 128, 7)) + [1, 2, 126, 127]:
        for shift in (0, 16):
            other = rng.integers(0, 1 << 16, size=lo.shape, dtype=np.uint64) << np.uint64(16 - shift)
            w = (lo << np.uint64(shift)) | (hi << np.uint64(shift + 8)) | other
            got = _loose(w, t)
            d = np.stack([(w >> np.uint64(8 * k)) & np.uint64(0xff) for k in range(4)], -1).astype(np.int64)
            exact = d > t
            assert not (exact & ~got).any(), "false negative at t=%d" % t
            fp = got & ~exact
            # a false positive needs a tie d == t and a carry out of the byte below (that byte >= 129 + t, or
            # 128 + t when it receives a carry itself)
            below = np.concatenate([np.zeros(d.shape[:-1] + (1,), np.int64), d[..., :-1]], -1)
            assert ((d == t) & (below >= 128 + t))[fp].all(), "unexplained false positive at t=%d" % t


def test_compass_pretest_never_rejects_a_corner(oracle):
    """(N or S) and (E or W) differ by more than t is necessary for 9 contiguous ring pixels (unsigned form)."""
    lay = pyramid_layout(320, 240, nlevels=2)
    for stack, t in ((synthetic_stack(lay, 5), 20), (noise_stack(lay, 6), 40), (noise_stack(lay, 7), 3)):
        for w, h, r in zip(lay.widths, lay.heights, lay.row_offsets):
            img = np.ascontiguousarray(stack[r:r + h])
            out = np.zeros_like(img)
            oracle.fast_detect(img, out, 16, w, h, t)
            I = img.astype(np.int32)
            sh = lambda dx, dy: np.roll(np.roll(I, -dy, 0), -dx, 1)  # noqa: E731
            u = [np.abs(sh(*o) - I) > t for o in ((0, -3), (3, 0), (0, 3), (-3, 0))]
            pre = (u[0] | u[2]) & (u[1] | u[3])
            corners = np.zeros_like(pre)
            corners[16:h - 16, 16:w - 16] = out[16:h - 16, 16:w - 16] != 0
            assert corners.sum() > 0
            assert not (corners & ~pre).any()


def test_diagonal_pretest_never_rejects_a_corner_and_removes_a_third_of_the_candidates(oracle):
    """Phase A2 of the tile kernel adds (NE or SW) and (SE or NW) on the ring pixels (+-2, +-2) to the compass test:
    a 9-arc contains two neighbouring diagonal ring pixels as well.  Necessary for every FAST-9 corner of the oracle,
    and on synthetic frames it removes more than a third of the compass test's candidates (2.6
    candidates per corner against 4.6 over the 8-level bench pyramids)."""
    lay = pyramid_layout(640, 480, nlevels=3)
    n0 = n1 = nc = 0
    for stack, t in ((synthetic_stack(lay, 0), 20), (synthetic_stack(lay, 1), 20), (noise_stack(lay, 6), 40)):
        for w, h, r in zip(lay.widths, lay.heights, lay.row_offsets):
            img = np.ascontiguousarray(stack[r:r + h])
            out = np.zeros_like(img)
            oracle.fast_detect(img, out, 16, w, h, t)
            I = img.astype(np.int32)
            sh = lambda dx, dy: np.roll(np.roll(I, -dy, 0), -dx, 1)  # noqa: E731
            u = {o: np.abs(sh(*o) - I) > t for o in ((0, -3), (3, 0), (0, 3), (-3, 0), (2, -2), (2, 2), (-2, 2), (-2, -2))}
            compass = (u[(0, -3)] | u[(0, 3)]) & (u[(3, 0)] | u[(-3, 0)])
            diag = (u[(2, -2)] | u[(-2, 2)]) & (u[(2, 2)] | u[(-2, -2)])
            inner = np.zeros_like(compass)
            inner[16:h - 16, 16:w - 16] = True
            corners = inner & (out != 0)
            assert not (corners & ~(compass & diag)).any()
            if t == 20:  # the synthetic frames only: uniform noise passes everything
                n0 += int((compass & inner).sum())
                n1 += int((compass & diag & inner).sum())
                nc += int(corners.sum())
    assert nc > 2000
    assert n1 < 0.65 * n0 and n1 < 4.0 * nc, (n0, n1, nc)
