#!/usr/bin/env python
"""e2e (host buffers in, host records out) for several pipeline chunk counts, stack input and base-frame input.

    python tools/gpu_e2e_chunks.py [chunk counts ...]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
F, STEPS, CAP, WORDS = 256, 20, 2048, 8


def main():
    import torch
    from pislam_b200 import api, pyramid
    chunks = [int(a) for a in sys.argv[1:]] or [4, 8, 16, 32]
    lay = pyramid.VGA
    desc = api.make_desc(lay, log_bucket=4, bucket_limit=3, words=WORDS)
    rec_bytes = api.record_layout(CAP, WORDS)["bytes"]
    bases = [pyramid.synthetic_base(640, 480, 0x5EED0000 + i) for i in range(32)]
    stacks = [pyramid.build_stack(b, lay) for b in bases]
    h_stacks = torch.empty((F, lay.rows, lay.vstep), dtype=torch.uint8, pin_memory=True)
    h_bases = torch.empty((F, 480, 640), dtype=torch.uint8, pin_memory=True)
    for f in range(F):
        h_stacks.numpy()[f] = stacks[f % 32]
        h_bases.numpy()[f] = bases[f % 32]
    h_rec = torch.empty(F * rec_bytes, dtype=torch.uint8, pin_memory=True)
    for rnd in range(2):
        for ch in chunks:
            os.environ["PISLAM_B200_CHUNKS"] = str(ch)
            ctx = api.Context(0)
            res = []
            for name, fn in (("stacks", lambda: ctx.extract_batch_ptr(h_stacks.data_ptr(), lay.stack_bytes, F, desc, CAP, h_rec.data_ptr())),
                             ("bases", lambda: ctx.extract_frames_ptr(h_bases.data_ptr(), 640, 640 * 480, F, desc, CAP, h_rec.data_ptr()))):
                for _ in range(3):
                    assert fn() == 0
                t0 = time.perf_counter()
                for _ in range(STEPS):
                    assert fn() == 0
                ms = (time.perf_counter() - t0) / STEPS * 1e3
                res.append("%s %.3f ms (%.1f Gpix/s)" % (name, ms, lay.level_pixels * F / ms / 1e6))
            print("chunks %2d  " % ch + "   ".join(res), flush=True)
            ctx.close()


if __name__ == "__main__":
    main()
