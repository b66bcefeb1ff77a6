"""Multi-GPU check of the path's only exchange step: every rank extracts its shard of frames
with pislam_b200_extract_allgather (records written in place into the gather buffer, one collective on the second
stream) and EVERY rank must end up with all frames' records, bit-exact against the oracle.  Covered: even shards,
ragged shards (5 frames over 2 ranks, 11 over 8) and empty shards (1 frame over 2 ranks, 3 over 8: trailing ranks own
nothing and still take part), alternating gather buffers (the overlapped form bench.py times), the same buffer twice
in a row, a third buffer (eviction of the tracked gather), and pislam_b200_extract_batch + the stand-alone
pislam_b200_allgather.  Buffers 0 and 1 are registered with pislam_b200_comm_register and take the size-aware
exchange over NVLink peer mappings (only the live part of every record travels), buffer 2 is not and goes through
ncclAllGather: both must deliver the same records.  Runs at 2 ranks and at 8 ranks; each is skipped when the box has fewer GPUs:
`gpurun --gpus 8 -- python -m pytest tests/test_gpu_multirank.py -m gpu`."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

CAP, WORDS = 2048, 8


def worker(rank, world, port, frame_counts, out):
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
    desc = api.make_desc(VGA, log_bucket=4, bucket_limit=3, words=WORDS)
    rec_bytes = api.record_layout(CAP, WORDS)["bytes"]
    ctx = api.Context(rank)
    uid = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(world, rank, uid[0])
    oracle = oracle_lib.Oracle()
    nmax = max(frame_counts)
    all_stacks = [synthetic_stack(VGA, 500 + f) for f in range(nmax)]
    want = [oracle.extract_pyramid(s, VGA.widths, VGA.heights, VGA.row_offsets, log_bucket=4, bucket_limit=3)
            for s in all_stacks]
    failures = []
    # gather buffers for the largest case, alive for the whole test: 0 and 1 registered for the peer-mapped exchange
    max_per = max(sharding.frames_per_rank(n, world) for n in frame_counts)
    big = [torch.empty(world * max_per * rec_bytes, dtype=torch.uint8, device="cuda") for _ in range(3)]
    registered = [ctx.comm_register(b.data_ptr(), b.numel()) for b in big[:2]]
    if os.environ.get("PISLAM_B200_NO_P2P") is None and not all(registered):
        failures.append("comm_register fell back to NCCL: %s" % registered)
    if ctx.p2p_windows != sum(registered):
        failures.append("p2p_windows %d vs %s" % (ctx.p2p_windows, registered))

    for nframes in frame_counts:
        per = sharding.frames_per_rank(nframes, world)
        start, count = sharding.shard_frames(nframes, world, rank)
        stacks = np.stack(all_stacks[start:start + count]) if count else np.zeros((1, VGA.rows, VGA.vstep), np.uint8)
        d_stacks = torch.from_numpy(stacks).cuda()
        bufs = [b[:world * per * rec_bytes] for b in big]
        for b in bufs:
            b.fill_(0xA5)
        torch.cuda.synchronize()
        dist.barrier()  # nobody pushes into a buffer a slower rank is still filling with 0xA5

        def check(buf, what):
            recs = api.parse_records(buf.cpu().numpy(), world * per, CAP, WORDS)
            good = True
            seen = set()
            for f, (r, s) in enumerate(sharding.gathered_index(nframes, world)):
                kp, lv, dsc, _, n_total = recs[r * per + s]
                seen.add(r * per + s)
                good &= (n_total == len(want[f][0]) and np.array_equal(kp, want[f][0]) and
                         np.array_equal(lv, want[f][1]) and np.array_equal(dsc, want[f][2]))
            for i in set(range(world * per)) - seen:  # unused slots of ragged / empty shards read "no keypoints"
                good &= len(recs[i][0]) == 0 and recs[i][4] == 0
            if not good:
                failures.append("%d frames: %s" % (nframes, what))

        def extract(buf):
            st = ctx.extract_allgather_ptr(d_stacks.data_ptr() if count else 0, VGA.stack_bytes, count, desc, CAP,
                                           buf.data_ptr(), slots_per_rank=per)
            if st != 0:
                failures.append("%d frames: extract_allgather status %d" % (nframes, st))

        # (1) alternating buffers, several rounds back to back (overlapped gathers)
        for it in range(6):
            extract(bufs[it & 1])
        ctx.synchronize()
        check(bufs[0], "alternating buffer 0")
        check(bufs[1], "alternating buffer 1")
        # (2) the same buffer twice in a row (the second extraction must wait for the first gather), then a third
        # buffer (evicts the older tracked gather) and back
        bufs[0].fill_(0x5A)
        torch.cuda.synchronize()  # the fill runs on torch's stream, the exchange on the library's
        for b in (0, 0, 2, 1, 0):
            extract(bufs[b])
        ctx.synchronize()
        for b in range(3):
            check(bufs[b], "buffer %d after the 0,0,2,1,0 sequence" % b)
        # (2b) HOST stacks and HOST base frames as the source of the rank's shard (chunked pipeline into the slot, then
        # the same exchange), into a registered and into an unregistered buffer
        if count:
            from pislam_b200.pyramid import synthetic_base
            bases = np.stack([synthetic_base(640, 480, 0x5EED0000 + 500 + f) for f in range(start, start + count)])
        else:
            bases = np.zeros((1, 480, 640), np.uint8)
        for b in (1, 2):
            bufs[b].fill_(0x33)
        torch.cuda.synchronize()
        dist.barrier()
        st = ctx.extract_allgather_ptr(stacks.ctypes.data if count else 0, VGA.stack_bytes, count, desc, CAP,
                                       bufs[1].data_ptr(), slots_per_rank=per)
        st2 = ctx.extract_frames_allgather_ptr(bases.ctypes.data if count else 0, 640, 640 * 480, count, desc, CAP,
                                               bufs[2].data_ptr(), slots_per_rank=per)
        ctx.synchronize()
        if st != 0 or st2 != 0:
            failures.append("%d frames: host-source extract_allgather %d / frames_allgather %d" % (nframes, st, st2))
        check(bufs[1], "host stacks -> registered buffer")
        check(bufs[2], "host base frames -> unregistered buffer")
        # (3) extract_batch into this rank's slot, then the stand-alone in-place all-gather (even shards only:
        # pislam_b200_allgather moves whole slots and knows nothing about unused ones)
        if nframes == world * per:
            bufs[1].zero_()
            torch.cuda.synchronize()
            slot = bufs[1][rank * per * rec_bytes:(rank + 1) * per * rec_bytes]
            st = ctx.extract_batch_ptr(stacks.ctypes.data, VGA.stack_bytes, count, desc, CAP, slot.data_ptr())
            st2 = ctx.allgather_ptr(bufs[1].data_ptr(), per * rec_bytes)
            ctx.synchronize()
            if st != 0 or st2 != 0:
                failures.append("%d frames: extract_batch %d / allgather %d" % (nframes, st, st2))
            check(bufs[1], "extract_batch + allgather")
        ctx.synchronize()
        dist.barrier()
        del bufs, d_stacks
    out[rank] = failures
    dist.barrier()
    ctx.comm_destroy()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 8])
def test_multi_gpu_allgather_matches_oracle(world):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    # even / ragged / (mostly) empty shards
    frame_counts = [6, 5, 1] if world == 2 else [16, 11, 3]
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29600 + (os.getpid() % 2000) + world
    mp.spawn(worker, args=(world, port, frame_counts, out), nprocs=world, join=True)
    assert dict(out) == {r: [] for r in range(world)}
