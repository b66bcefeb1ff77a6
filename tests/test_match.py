"""Descriptor matcher: oracle self-checks on CPU, CUDA-vs-oracle parity on the GPU.
The reference has no matcher (README.md:125-139), so the oracle is pinned against scipy only: "parity unpinned"."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
import match_oracle  # noqa: E402


def _rand_desc(rng, n, words):
    return rng.integers(0, 1 << 32, size=(n, words), dtype=np.uint64).astype(np.uint32)


def test_oracle_distance_against_scipy():
    from scipy.spatial.distance import cdist
    rng = np.random.Generator(np.random.PCG64(1))
    for words in (1, 3, 8):
        q, t = _rand_desc(rng, 37, words), _rand_desc(rng, 53, words)
        bits = lambda a: np.unpackbits(a.view(np.uint8), axis=1)  # noqa: E731
        want = np.rint(cdist(bits(q), bits(t), "hamming") * 32 * words).astype(np.int32)
        assert np.array_equal(match_oracle.hamming_matrix(q, t), want)


def test_oracle_ties_window_and_empty():
    rng = np.random.Generator(np.random.PCG64(2))
    t = _rand_desc(rng, 10, 8)
    t[7] = t[2]  # duplicate train descriptor: the smaller index wins, the duplicate is the runner-up
    q = t[[2, 5]].copy()
    m = match_oracle.match_hamming(q, t)
    assert m[0].tolist() == [2, 0, 0] and m[1, 0] == 5 and m[1, 1] == 0 and m[1, 2] > 0
    # window: only same-level keypoints within the radius qualify
    enc = lambda x, y: np.uint32((x << 12) | y)  # noqa: E731
    t_kp = np.array([enc(100 + 3 * i, 50) for i in range(10)], np.uint32)
    t_lv = np.array([0] * 5 + [1] * 5, np.uint8)
    q_kp, q_lv = np.array([enc(100, 50), enc(400, 400)], np.uint32), np.array([0, 0], np.uint8)
    m = match_oracle.match_hamming(q, t, q_kp, q_lv, t_kp, t_lv, radius=6)
    assert m[0, 0] == 2 and m[0, 1] == 0        # candidates 0,1,2 (x = 100,103,106): index 2 is the duplicate source
    assert m[1].tolist() == [-1, match_oracle.MATCH_NONE, match_oracle.MATCH_NONE]
    assert match_oracle.match_hamming(q, t[:0]).tolist() == [[-1, match_oracle.MATCH_NONE, match_oracle.MATCH_NONE]] * 2
    assert match_oracle.match_hamming(q[:0], t).shape == (0, 3)
    one = match_oracle.match_hamming(q[:1], t[2:3])
    assert one.tolist() == [[0, 0, match_oracle.MATCH_NONE]]


@pytest.fixture(scope="module")
def ctx():
    from pislam_b200 import api
    c = api.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
def test_match_hamming_parity(ctx):
    rng = np.random.Generator(np.random.PCG64(3))
    for words, nq, nt in ((8, 1000, 1100), (8, 129, 127), (5, 64, 300), (1, 200, 40), (8, 1, 1), (8, 300, 0), (8, 0, 5)):
        q, t = _rand_desc(rng, nq, words), _rand_desc(rng, nt, words)
        if nt > 20 and nq > 20:  # planted duplicates and exact ties
            t[nt // 2] = t[3]
            q[:10] = t[rng.integers(0, nt, 10)]
            q[10] = t[3]
        got = ctx.match_hamming(q, t)
        assert np.array_equal(got, match_oracle.match_hamming(q, t)), (words, nq, nt)
    # one word with few distinct values: every distance is a many-way tie
    q = rng.integers(0, 4, size=(257, 1)).astype(np.uint32)
    t = rng.integers(0, 4, size=(513, 1)).astype(np.uint32)
    assert np.array_equal(ctx.match_hamming(q, t), match_oracle.match_hamming(q, t))


@pytest.mark.gpu
def test_match_windowed_and_records_parity(ctx):
    """Two consecutive synthetic frames (the second shifted by a few pixels): windowed matches from the records
    of a batch against the oracle on the parsed keypoints, plus the plain-array entry point."""
    from pislam_b200 import api
    from pislam_b200.pyramid import VGA, synthetic_stack
    a = synthetic_stack(VGA, 3)
    b = np.zeros_like(a)
    for w, h, r in zip(VGA.widths, VGA.heights, VGA.row_offsets):  # shift every level by (+3, +2)
        b[r + 2:r + h, 3:w] = a[r:r + h - 2, :w - 3]
    c = synthetic_stack(VGA, 4)
    stacks = np.stack([a, b, c])
    cap, words = 2048, 8
    d = api.make_desc(VGA, log_bucket=4, bucket_limit=3)
    lay = api.record_layout(cap, words)
    rec = np.zeros(3 * lay["bytes"], np.uint8)
    import ctypes
    st = ctx.lib.pislam_b200_extract_batch(ctx._h, stacks.ctypes.data, stacks.shape[1] * stacks.shape[2], 3,
                                           ctypes.byref(d), cap, rec.ctypes.data)
    assert st == 0
    frames = api.parse_records(rec, 3, cap, words)
    pairs = [(0, 1), (1, 0), (0, 2), (2, 2)]
    for radius in (0, 20, 5):
        got = ctx.match_records(rec, 3, cap, words, pairs, radius=radius)
        for p, (fa, fb) in enumerate(pairs):
            kq, lq, dq = frames[fa][:3]
            kt, lt, dt = frames[fb][:3]
            want = match_oracle.match_hamming(dq, dt, kq, lq, kt, lt, radius=radius)
            assert np.array_equal(got[p, :len(kq)], want), (radius, fa, fb)
            assert (got[p, len(kq):] == -2).all()  # rows beyond the query frame's keypoints stay untouched
            if p == 0:
                plain = ctx.match_hamming(dq, dt, kq, lq, kt, lt, radius=radius)
                assert np.array_equal(plain, want)
    # the shifted frame matches itself: with a 20-pixel window most keypoints find a close descriptor
    m = ctx.match_records(rec, 3, cap, words, [(0, 1)], radius=20)[0, :len(frames[0][0])]
    good = (m[:, 0] >= 0) & (m[:, 1] <= 40)
    assert good.mean() > 0.5
