"""CPU oracle of the descriptor matcher -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED BY THE REFERENCE: 0xfaded/pislam never released its matcher (README.md:125-139 only reports
"feature matching ... under 20 ms per frame using flann" with table-driven bit counts and a demo that traces "best
matches within a 20 pixel radius"), so there is no reference code, test or golden vector to pin this against.  The
definition is the one stated in include/pislam_b200.h; this file restates it in numpy, independently of the CUDA
kernel (a per-byte popcount table, like the reference's README describes, instead of POPC), and
tests/test_match.py cross-checks it against scipy's hamming cdist.

Only tests/ may import this module.
"""
import numpy as np

MATCH_NONE = 0x7fffffff
_POP8 = np.array([bin(i).count("1") for i in range(256)], np.int32)


def hamming_matrix(q_desc, t_desc):
    """int32 [nq][nt]: popcount of the XOR of the descriptor words."""
    if len(q_desc) == 0 or len(t_desc) == 0:
        return np.zeros((len(q_desc), len(t_desc)), np.int32)
    q = np.ascontiguousarray(q_desc, np.uint32).reshape(len(q_desc), -1)
    t = np.ascontiguousarray(t_desc, np.uint32).reshape(len(t_desc), -1)
    x = (q[:, None, :] ^ t[None, :, :]).view(np.uint8)  # [nq][nt][4*words]
    return _POP8[x].sum(-1, dtype=np.int32)


def decode_xy(kp):
    kp = np.asarray(kp, np.uint32)
    return ((kp >> 12) & 0xfff).astype(np.int64), (kp & 0xfff).astype(np.int64)  # Util.h:27-45


def match_hamming(q_desc, t_desc, q_kp=None, q_level=None, t_kp=None, t_level=None, radius=0):
    """int32 [nq][3]: (index of the closest candidate or -1, its distance, the runner-up's distance).
    Candidates: every train descriptor (radius <= 0) or the train keypoints on the same level with
    dx^2 + dy^2 <= radius^2.  The smallest index wins ties; a tie with the best is the runner-up."""
    d = hamming_matrix(q_desc, t_desc).astype(np.int64)
    nq, nt = d.shape
    out = np.empty((nq, 3), np.int32)
    out[:, 0] = -1
    out[:, 1:] = MATCH_NONE
    if nq == 0 or nt == 0:
        return out
    if radius > 0:
        qx, qy = decode_xy(q_kp)
        tx, ty = decode_xy(t_kp)
        ok = (np.asarray(q_level)[:, None] == np.asarray(t_level)[None, :]) & \
             ((qx[:, None] - tx[None, :]) ** 2 + (qy[:, None] - ty[None, :]) ** 2 <= radius * radius)
        d = np.where(ok, d, np.int64(1) << 40)
    for i in range(nq):
        row = d[i]
        j = int(np.argmin(row))  # first minimum = smallest index
        if row[j] >= (1 << 40):
            continue
        out[i, 0], out[i, 1] = j, row[j]
        rest = np.delete(row, j)
        if len(rest) and rest.min() < (1 << 40):
            out[i, 2] = rest.min()
    return out
