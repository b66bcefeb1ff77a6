"""Two-GPU check of the path's only collective: every rank extracts its shard of frames
with pislam_b200_extract_allgather (records written in place into the gather buffer, one ncclAllGather on the
second stream) and EVERY rank must end up with all frames' records, bit-exact against the oracle; the same through
pislam_b200_extract_batch + pislam_b200_allgather, and with two alternating gather buffers (the overlapped form
bench.py times).  Needs >= 2 B200s (skipped otherwise): `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu`."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

CAP, WORDS, PER_RANK = 2048, 8, 3


def worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle_lib
    from pislam_b200 import api, sharding
    from pislam_b200.pyramid import VGA, synthetic_stack
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)  # host-side rendezvous only
    nframes = world * PER_RANK
    start, count = sharding.shard_frames(nframes, world, rank)
    stacks = np.stack([synthetic_stack(VGA, 500 + f) for f in range(start, start + count)])
    d_stacks = torch.from_numpy(stacks).cuda()
    desc = api.make_desc(VGA, log_bucket=4, bucket_limit=3, words=WORDS)
    rec_bytes = api.record_layout(CAP, WORDS)["bytes"]
    ctx = api.Context(rank)
    uid = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(world, rank, uid[0])
    bufs = [torch.zeros(world * PER_RANK * rec_bytes, dtype=torch.uint8, device="cuda") for _ in range(2)]
    ok = True
    oracle = oracle_lib.Oracle()
    want = [oracle.extract_pyramid(synthetic_stack(VGA, 500 + f), VGA.widths, VGA.heights, VGA.row_offsets,
                                   log_bucket=4, bucket_limit=3) for f in range(nframes)]

    def check(buf):
        good = True
        recs = api.parse_records(buf.cpu().numpy(), world * PER_RANK, CAP, WORDS)
        for f, (r, s) in enumerate(sharding.gathered_index(nframes, world)):
            kp, lv, dsc, _, n_total = recs[r * PER_RANK + s]
            good &= np.array_equal(kp, want[f][0]) and np.array_equal(lv, want[f][1]) and np.array_equal(dsc, want[f][2])
        return good

    # (1) fused extract + gather, alternating buffers, several rounds back to back (overlapped gathers)
    for it in range(6):
        st = ctx.extract_allgather_ptr(d_stacks.data_ptr(), VGA.stack_bytes, count, desc, CAP, bufs[it & 1].data_ptr())
        ok &= st == 0
    ctx.synchronize()
    ok &= check(bufs[0]) and check(bufs[1])
    # (2) the same buffer twice in a row (the second extraction must wait for the first gather)
    bufs[0].zero_()
    for it in range(2):
        ok &= ctx.extract_allgather_ptr(d_stacks.data_ptr(), VGA.stack_bytes, count, desc, CAP, bufs[0].data_ptr()) == 0
    ctx.synchronize()
    ok &= check(bufs[0])
    # (3) extract_batch into this rank's slot, then the stand-alone in-place all-gather
    bufs[1].zero_()
    slot = bufs[1][rank * PER_RANK * rec_bytes:(rank + 1) * PER_RANK * rec_bytes]
    ok &= ctx.extract_batch_ptr(stacks.ctypes.data, VGA.stack_bytes, count, desc, CAP, slot.data_ptr()) == 0
    ok &= ctx.allgather_ptr(bufs[1].data_ptr(), PER_RANK * rec_bytes) == 0
    ctx.synchronize()
    ok &= check(bufs[1])
    out[rank] = bool(ok)
    dist.barrier()
    ctx.comm_destroy()
    ctx.close()
    dist.destroy_process_group()


def test_two_gpu_allgather_matches_oracle():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
