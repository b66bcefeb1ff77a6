"""Small fused + literal workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck)."""
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
import oracle_lib
from pislam_b200 import api
from pislam_b200.pyramid import pyramid_layout, synthetic_stack, noise_stack

lay = pyramid_layout(320, 240, nlevels=4)
o = oracle_lib.Oracle()
ok = True
with api.Context(0) as ctx:
    for stack, name in ((synthetic_stack(lay, 7), "synthetic"), (noise_stack(lay, 3), "noise")):
        for lb, bl in ((0, 5), (4, 3), (2, 1)):
            for tma in (1, 0):
                ctx.set_tma(tma)
                kp, lv, desc = ctx.extract_pyramid(stack, api.make_desc(lay, log_bucket=lb, bucket_limit=bl), cap=1 << 15)
                okp, olv, odesc = o.extract_pyramid(stack, lay.widths, lay.heights, lay.row_offsets, log_bucket=lb,
                                                    bucket_limit=bl, cap=1 << 15)
                good = np.array_equal(kp, okp) and np.array_equal(desc, odesc)
                ok &= good
                print(name, "buckets", (lb, bl), "tma", tma, "n", len(kp), "match", good)
    ctx.set_tma(1)
    # 48 frames: enough tiles for the host to pick the 128x64 instantiation (single frames above use 128x32)
    src = [synthetic_stack(lay, 20 + i) for i in range(5)] + [noise_stack(lay, 9)]
    stacks = np.stack([src[i % 6] for i in range(48)])
    for lb, bl in ((4, 3), (0, 5)):
        res = ctx.extract_batch(stacks, api.make_desc(lay, log_bucket=lb, bucket_limit=bl), cap=1 << 15)
        for i in range(6):
            okp, olv, odesc = o.extract_pyramid(src[i], lay.widths, lay.heights, lay.row_offsets, log_bucket=lb,
                                                bucket_limit=bl, cap=1 << 15)
            good = all(np.array_equal(res[j][0], okp) and np.array_equal(res[j][2], odesc) for j in range(i, 48, 6))
            ok &= good
            print("batch frame", i, "buckets", (lb, bl), "n", len(okp), "match", good)
    img = np.ascontiguousarray(stack[:240])
    out = np.zeros_like(img)
    ctx.fastDetect(320, 240, img, out, 20)
    ctx.fastScoreHarris(320, 240, img, 1 << 15, out)
    kp = ctx.fastExtract(320, 240, out)
    d = ctx.orbCompute(img, kp[:64])
    print("literal", len(kp), d.shape)
print("ALL MATCH" if ok else "MISMATCH")
