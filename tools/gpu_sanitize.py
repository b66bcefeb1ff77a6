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
    # pyramid builder in front of the path (host base frames, chunked pipeline, blur on and off) and the issue-only form
    from pislam_b200.pyramid import synthetic_base
    bases = np.stack([synthetic_base(320, 240, 0x5EED0000 + 70 + i) for i in range(3)] * 8)
    for blur in (False, True):
        d = api.make_desc(lay, log_bucket=4, bucket_limit=3)
        res = ctx.extract_frames(bases, d, cap=1 << 13, blur=blur)
        for i in range(3):
            st_ = o.build_pyramid(bases[i], lay, blur=blur)
            okp, olv, odesc = o.extract_pyramid(st_, lay.widths, lay.heights, lay.row_offsets, log_bucket=4, bucket_limit=3,
                                                cap=1 << 13)
            good = all(np.array_equal(res[j][0], okp) and np.array_equal(res[j][2], odesc) for j in range(i, 24, 3))
            good &= np.array_equal(ctx.build_pyramid(bases[i:i + 1], d, blur=blur)[0], st_)
            ok &= good
            print("base frame", i, "blur", blur, "n", len(okp), "match", good)
    # keypoints only with border 5: plain-load staging with linear addressing and the pitch-overrun spill
    nz = noise_stack(lay, 12)
    d5 = api.make_desc(lay, border=5, fast_threshold=45, harris_threshold=1 << 12, words=0)
    kp, lv, _ = ctx.extract_pyramid(nz, d5, cap=1 << 16)
    okp, olv, _ = o.extract_pyramid(nz, lay.widths, lay.heights, lay.row_offsets, border=5, fast_threshold=45,
                                    harris_threshold=1 << 12, words=0, cap=1 << 16)
    good = np.array_equal(kp, okp)
    ok &= good
    print("border 5 noise n", len(okp), "match", good)
    img = np.ascontiguousarray(stack[:240])
    out = np.zeros_like(img)
    ctx.fastDetect(320, 240, img, out, 20)
    ctx.fastScoreHarris(320, 240, img, 1 << 15, out)
    kp = ctx.fastExtract(320, 240, out)
    d = ctx.orbCompute(img, kp[:64])
    print("literal", len(kp), d.shape)
print("ALL MATCH" if ok else "MISMATCH")
