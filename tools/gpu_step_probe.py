#!/usr/bin/env python
"""Device-resident step time at N ranks (torchrun) for several exchange settings: copy-kernel CTAs of the size-aware
exchange, ncclAllGather on unregistered buffers, and the gathered content checked against rank-local records.

    python -m torch.distributed.run --nproc-per-node N tools/gpu_step_probe.py [ctas ...]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
F, STEPS, CAP, WORDS = 256, 100, 2048, 8


def main():
    import torch
    import torch.distributed as dist
    from pislam_b200 import api, pyramid
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctas_list = [int(a) for a in sys.argv[1:]] or [32]
    lay = pyramid.VGA
    desc = api.make_desc(lay, log_bucket=4, bucket_limit=3, words=WORDS)
    rec_bytes = api.record_layout(CAP, WORDS)["bytes"]
    base = [pyramid.synthetic_stack(lay, i) for i in range(8)]
    d_stacks = torch.from_numpy(np.stack([base[i % 8] for i in range(F)])).cuda()
    ctx = api.Context(local)
    uid = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(world, rank, uid[0])
    reg = [torch.zeros(world * F * rec_bytes, dtype=torch.uint8, device="cuda") for _ in range(2)]
    unreg = [torch.zeros(world * F * rec_bytes, dtype=torch.uint8, device="cuda") for _ in range(2)]
    ok = all([ctx.comm_register(b.data_ptr(), b.numel()) for b in reg])
    stream = torch.cuda.ExternalStream(ctx.stream)

    def run(bufs, tag):
        turn = 0
        def step():
            nonlocal turn
            st = ctx.extract_allgather_ptr(d_stacks.data_ptr(), lay.stack_bytes, F, desc, CAP, bufs[turn].data_ptr())
            turn ^= 1
            assert st == 0, st
        for _ in range(10):
            step()
        dist.barrier(); ctx.synchronize(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        import time
        e0.record(stream)
        h0 = time.perf_counter()
        for _ in range(STEPS):
            step()
        host_ms = (time.perf_counter() - h0) * 1e3 / STEPS  # host time to ISSUE one step (no waiting for the GPU)
        ctx.join()
        e1.record(stream)
        dist.barrier(); ctx.synchronize(); torch.cuda.synchronize()
        # per-kernel and exchange spans (profiling events perturb the pipeline a little: separate, shorter run)
        ctx.set_profiling(True)
        for _ in range(20):
            step()
        kms, cnt = ctx.kernel_times()
        ctx.set_profiling(False)
        dist.barrier(); ctx.synchronize(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1), host_ms, kms[4] / max(cnt[4], 1), kms[0] / max(cnt[0], 1), kms[2] / max(cnt[2], 1)],
                         device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tag = "%s  [max over ranks: host issue %.3f ms/step, exchange span %.3f, front %.3f, orb %.3f]" % (
            tag, t[1].item(), t[2].item(), t[3].item(), t[4].item())
        # content: every slot of the gathered buffer equals this rank's own slot (all ranks hold the same frames)
        mine = bufs[0][rank * F * rec_bytes:(rank + 1) * F * rec_bytes]
        recs = api.parse_records(bufs[0].cpu().numpy(), world * F, CAP, WORDS)
        good = all(np.array_equal(recs[r * F + f][0], recs[rank * F + f][0]) and
                   np.array_equal(recs[r * F + f][2], recs[rank * F + f][2]) for r in range(world) for f in (0, F // 2, F - 1))
        g = torch.tensor([0 if good else 1], device="cuda")
        dist.all_reduce(g)
        if rank == 0:
            print("%s\n     -> %.4f ms/step  %8.0f Gpix/s  content %s" % (tag, t[0].item() / STEPS, lay.level_pixels * F * world * STEPS / t[0].item() / 1e6,
                                                                     "ok" if g.item() == 0 else "MISMATCH"), flush=True)

    if rank == 0:
        print("world %d, registered %s" % (world, ok), flush=True)
    for rnd in range(2):
        for c in ctas_list:
            ctx.set_exchange_ctas(c)
            run(reg, "p2p push, %d CTAs" % c)
        run(unreg, "ncclAllGather")
    dist.barrier()
    ctx.comm_destroy()
    del d_stacks, reg, unreg, stream
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
