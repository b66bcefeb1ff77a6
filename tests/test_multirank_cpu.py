"""World-size-2 gloo test of the N>1 host logic (CPU only): frame sharding, the fixed-capacity record layout and
the rank-major all-gather.  The per-rank 'extraction' is the CPU oracle packed into records by a test-side
packer that follows include/pislam_b200.h, so what is checked is exactly what bench.py / a caller does around
pislam_b200_extract_allgather: shard -> extract -> one all-gather -> parse."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from pislam_b200 import api, sharding  # noqa: E402
from pislam_b200.pyramid import pyramid_layout, synthetic_stack  # noqa: E402

CAP, WORDS, NFRAMES = 512, 8, 5
LAY = pyramid_layout(160, 120, nlevels=3)


def pack_record(kp, lv, desc, cap, words, nlevels):
    lay = api.record_layout(cap, words)
    r = np.zeros(lay["bytes"], np.uint8)
    hdr = r[:128].view(np.uint32)
    n = len(kp)
    hdr[0] = hdr[1] = n
    for l in range(nlevels + 1):
        hdr[2 + l] = int((lv < l).sum())
    r[lay["kp"]:lay["kp"] + 4 * n] = kp.view(np.uint8)
    r[lay["desc"]:lay["desc"] + 4 * n * words] = desc.ravel().view(np.uint8)
    r[lay["level"]:lay["level"] + n] = lv
    return r


def oracle_frame(oracle, f):
    stack = synthetic_stack(LAY, 100 + f)
    return oracle.extract_pyramid(stack, LAY.widths, LAY.heights, LAY.row_offsets, fast_threshold=12,
                                  harris_threshold=1 << 12, words=WORDS)


def worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle = oracle_lib.Oracle()
    per = sharding.frames_per_rank(NFRAMES, world)
    start, count = sharding.shard_frames(NFRAMES, world, rank)
    rec_bytes = api.record_layout(CAP, WORDS)["bytes"]
    mine = np.zeros(per * rec_bytes, np.uint8)  # unused slots stay all-zero records (n = 0)
    for s in range(count):
        kp, lv, desc = oracle_frame(oracle, start + s)
        mine[s * rec_bytes:(s + 1) * rec_bytes] = pack_record(kp, lv, desc, CAP, WORDS, LAY.nlevels)
    gathered = torch.zeros(world * per * rec_bytes, dtype=torch.uint8)
    dist.all_gather_into_tensor(gathered, torch.from_numpy(mine))  # the ONE collective of the path
    recs = api.parse_records(gathered.numpy(), world * per, CAP, WORDS)
    ok = True
    for f, (r, s) in enumerate(sharding.gathered_index(NFRAMES, world)):
        kp, lv, desc = oracle_frame(oracle, f)
        gkp, glv, gdesc, level_start, n_total = recs[r * per + s]
        ok &= np.array_equal(gkp, kp) and np.array_equal(glv, lv) and np.array_equal(gdesc, desc)
        ok &= n_total == len(kp) and int(level_start[LAY.nlevels]) == len(kp)
    for slot in range(NFRAMES, world * per):  # padding slots
        ok &= len(recs[slot][0]) == 0
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_helpers():
    assert [sharding.shard_frames(5, 2, r) for r in range(2)] == [(0, 3), (3, 2)]
    assert [sharding.shard_frames(8, 8, r) for r in range(8)] == [(r, 1) for r in range(8)]
    assert [sharding.shard_frames(3, 4, r) for r in range(4)] == [(0, 1), (1, 1), (2, 1), (3, 0)]
    assert sharding.gathered_index(5, 2) == [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1)]
    for n in range(1, 20):
        for w in (1, 2, 4, 8):
            cover = []
            for r in range(w):
                s, c = sharding.shard_frames(n, w, r)
                cover += list(range(s, s + c))
            assert cover == list(range(n))


def test_two_rank_gather_gloo():
    api.load_library()
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
