#!/usr/bin/env python
"""Sub-batch sweep on one box: whole-step time of the device-resident fused path (256 VGA pyramids) for several
sub-batch sizes (pislam_b200_set_sub_batch: descriptor kernel of sub-batch k under the detector of sub-batch k+1),
records compared byte for byte with the un-split run.  Also sweeps PISLAM_B200_CHUNKS-like host chunking for the
base-frame path when --base is given.

    python tools/gpu_sub_ab.py [sub sizes ...]
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
F, STEPS, CAP, WORDS = 256, 30, 2048, 8


def main():
    import torch
    from pislam_b200 import api, pyramid
    subs = [int(a) for a in sys.argv[1:]] or [0, 16, 32, 64, 128]
    lay = pyramid.VGA
    desc = api.make_desc(lay, border=16, fast_threshold=20, harris_threshold=1 << 15, log_bucket=4, bucket_limit=3,
                         words=WORDS)
    rec_bytes = api.record_layout(CAP, WORDS)["bytes"]
    base = [pyramid.synthetic_stack(lay, i) for i in range(32)]
    frames = np.stack([base[i % 32] for i in range(F)])
    d_stacks = torch.from_numpy(frames).cuda()
    d_rec = torch.zeros(F * rec_bytes, dtype=torch.uint8, device="cuda")
    ctx = api.Context(0)
    s = torch.cuda.ExternalStream(ctx.stream)
    ref = None
    for rnd in range(2):
        for sub in subs:
            ctx.set_sub_batch(sub)
            d_rec.zero_()

            def step():
                st = ctx.extract_batch_ptr(d_stacks.data_ptr(), lay.stack_bytes, F, desc, CAP, d_rec.data_ptr(),
                                           asynchronous=True)
                assert st == 0, st
            for _ in range(5):
                step()
            ctx.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(STEPS):
                step()
            e1.record(s)
            ctx.synchronize()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / STEPS
            ctx.set_profiling(True)
            for _ in range(5):
                step()
            kms, cnt = ctx.kernel_times()
            ctx.set_profiling(False)
            rec = d_rec.cpu().numpy()
            h = hashlib.sha1(rec.tobytes()).hexdigest()[:12]
            if ref is None:
                ref = rec
            print("sub %4d  step %.4f ms  (front %.4f x%d, order %.4f, orb %.4f x%d per step)  sha1 %s  identical %s" % (
                sub, ms, kms[0] / 5, cnt[0] // 5, kms[1] / 5, kms[2] / 5, cnt[2] // 5, h, bool((rec == ref).all())),
                flush=True)
    del d_stacks, d_rec, s
    ctx.close()


if __name__ == "__main__":
    main()
