#!/usr/bin/env python
"""pyramid_build_kernel alone: 256 VGA base frames resident in HBM -> stacks in HBM (pislam_b200_build_pyramid with
device pointers), CUDA events on the ctx stream.  python tools/gpu_pyramid_time.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import ctypes as C
    import torch
    from pislam_b200 import api, pyramid
    F = 256
    lay = pyramid.VGA
    desc = api.make_desc(lay)
    bases = np.stack([pyramid.synthetic_base(640, 480, 0x5EED0000 + i) for i in range(8)] * (F // 8))
    d_bases = torch.from_numpy(bases).cuda()
    d_stacks = torch.zeros((F, lay.rows, lay.vstep), dtype=torch.uint8, device="cuda")
    ctx = api.Context(0)
    s = torch.cuda.ExternalStream(ctx.stream)

    def run():
        st = ctx.lib.pislam_b200_build_pyramid(ctx._h, d_bases.data_ptr(), 640, 640 * 480, F, 0, C.byref(desc),
                                               d_stacks.data_ptr(), lay.stack_bytes)
        assert st == 0, st
    for _ in range(3):
        run()
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(20):
        run()
    e1.record(s)
    ctx.synchronize()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    want = pyramid.build_stack(bases[3], lay)
    got = d_stacks[3].cpu().numpy()
    print("pyramid_build_kernel: %.4f ms per %d VGA frames (%.0f Gpix/s of level pixels), frame 3 exact: %s" % (
        ms, F, lay.level_pixels * F / ms / 1e6, bool((got == want).all())), flush=True)
    del d_bases, d_stacks, s
    ctx.close()


if __name__ == "__main__":
    main()
