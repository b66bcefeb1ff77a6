#!/usr/bin/env python
"""Times pislam_b200_match_records on device-resident records: 256 synthetic VGA frames, the 255 consecutive
frame pairs, brute force and a 20-pixel window.  Prints ms per launch and comparisons/s."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pislam_b200 import api, pyramid  # noqa: E402

F, CAP, WORDS = 256, 2048, 8
lay = pyramid.VGA
desc = api.make_desc(lay, log_bucket=4, bucket_limit=3, words=WORDS)
base = [pyramid.synthetic_stack(lay, i) for i in range(32)]
stacks = torch.from_numpy(np.stack([base[i % 32] for i in range(F)])).cuda()
rec_bytes = api.record_layout(CAP, WORDS)["bytes"]
d_rec = torch.zeros(F * rec_bytes, dtype=torch.uint8, device="cuda")
ctx = api.Context(0)
assert ctx.extract_batch_ptr(stacks.data_ptr(), lay.stack_bytes, F, desc, CAP, d_rec.data_ptr()) == 0
hdr = d_rec.view(F, rec_bytes)[:, :4].contiguous().cpu().numpy().view(np.uint32)[:, 0].astype(np.int64)
pairs = np.array([(f, f + 1) for f in range(F - 1)], np.int32)
d_out = torch.zeros(len(pairs) * CAP * 3, dtype=torch.int32, device="cuda")
stream = torch.cuda.ExternalStream(ctx.stream)
comps = int((hdr[:-1] * hdr[1:]).sum())
for radius in (0, 20):
    def run():
        st = ctx.lib.pislam_b200_match_records(ctx._h, d_rec.data_ptr(), F, CAP, WORDS,
                                               pairs.ctypes.data_as(api._i32p), len(pairs), radius, d_out.data_ptr())
        assert st == 0, st
    for _ in range(3):
        run()
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(20):
        run()
    e1.record(stream)
    ctx.synchronize()
    ms = e0.elapsed_time(e1) / 20
    m = d_out.view(len(pairs), CAP, 3)[0, :hdr[0]].cpu().numpy()
    print("radius %2d: %.3f ms per launch (%d pairs, %.0f keypoints/frame): %.1f us per frame pair, %.2f T descriptor "
          "comparisons/s; pair 0: %d of %d matched, median distance %d" % (
              radius, ms, len(pairs), hdr.mean(), ms * 1e3 / len(pairs), comps / (ms * 1e-3) / 1e12,
              int((m[:, 0] >= 0).sum()), len(m), int(np.median(m[m[:, 0] >= 0, 1])) if (m[:, 0] >= 0).any() else -1))
del stacks, d_rec, d_out, stream
ctx.close()
