#!/usr/bin/env python
"""A/B of library builds on one box: kernel times (CUDA events around every launch, library profiling mode) and a
bit-exact check of the batch records against the first build's.

    python tools/gpu_front_ab.py ab/lib_a.so ab/lib_b.so ...       (parent: one subprocess per library, twice)
    python tools/gpu_front_ab.py --child <lib> <frames.npy> <out.npy>

The synthetic frames are built once and shared through /tmp; the child loads ONE library (PISLAM_B200_LIB).
Parity proper is tests/ (-m gpu); this only guards an A/B run against comparing a broken build.
"""
import hashlib
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
F, STEPS, CAP, WORDS = 256, 30, 2048, 8


def child(lib, frames_path, out_path):
    lib, _, envs = lib.partition("@")  # "ab/lib_x.so@VAR=VALUE,VAR2=VALUE2": environment of this variant
    for kv in filter(None, envs.split(",")):
        k, _, v = kv.partition("=")
        os.environ[k] = v
    tag = os.path.basename(lib) + ("@" + envs if envs else "")
    os.environ["PISLAM_B200_LIB"] = os.path.abspath(lib)
    import torch
    from pislam_b200 import api, pyramid
    lay = pyramid.VGA
    desc = api.make_desc(lay, border=16, fast_threshold=20, harris_threshold=1 << 15, log_bucket=4, bucket_limit=3,
                         words=WORDS)
    rec_bytes = api.record_layout(CAP, WORDS)["bytes"]
    frames = np.load(frames_path)
    d_stacks = torch.from_numpy(frames).cuda()
    d_rec = torch.zeros(F * rec_bytes, dtype=torch.uint8, device="cuda")
    ctx = api.Context(0)

    def step():
        st = ctx.extract_batch_ptr(d_stacks.data_ptr(), lay.stack_bytes, F, desc, CAP, d_rec.data_ptr(),
                                   asynchronous=True)
        assert st == 0, st
    for _ in range(5):
        step()
    ctx.synchronize()
    ctx.set_profiling(True)
    for _ in range(STEPS):
        step()
    ms, cnt = ctx.kernel_times()
    ctx.set_profiling(False)
    # whole-step time without the profiling events
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s = torch.cuda.ExternalStream(ctx.stream)
    ctx.synchronize()
    e0.record(s)
    for _ in range(STEPS):
        step()
    e1.record(s)
    ctx.synchronize()
    torch.cuda.synchronize()
    rec = d_rec.cpu().numpy()
    np.save(out_path, rec)
    hdr = rec.reshape(F, rec_bytes)[:, :8].copy().view(np.uint32)
    print("%-30s front %.4f ms  order %.4f  orb %.4f  step %.4f ms  kpts/frame %.1f  records sha1 %s" % (
        tag, ms[0] / max(cnt[0], 1), ms[1] / max(cnt[1], 1), ms[2] / max(cnt[2], 1),
        e0.elapsed_time(e1) / STEPS, hdr[:, 0].mean(), hashlib.sha1(rec.tobytes()).hexdigest()[:12]), flush=True)
    del d_stacks, d_rec, s
    ctx.close()


def main():
    if sys.argv[1] == "--child":
        return child(*sys.argv[2:5])
    libs = sys.argv[1:]
    from pislam_b200 import pyramid
    t0 = time.time()
    base = [pyramid.synthetic_stack(pyramid.VGA, i) for i in range(32)]
    frames = np.stack([base[i % 32] for i in range(F)])
    fp = "/tmp/ab_frames.npy"
    np.save(fp, frames)
    print("frames built in %.1f s" % (time.time() - t0), flush=True)
    ref = None
    rec_bytes = None
    for rnd in range(2):
        for lib in libs:
            out = "/tmp/ab_rec_%s.npy" % os.path.basename(lib).replace("@", "_").replace("=", "_").replace(",", "_")
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", lib, fp, out],
                                   stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
            except subprocess.TimeoutExpired:
                print("%s TIMED OUT" % lib, flush=True)
                continue
            sys.stdout.write(r.stdout)
            if r.returncode != 0:
                print("%s FAILED (%d): %s" % (lib, r.returncode, r.stderr[-600:]), flush=True)
                continue
            if rnd == 0:
                rec = np.load(out)
                if ref is None:
                    ref = rec
                else:
                    same = bool((rec == ref).all())
                    print("    records identical to %s: %s" % (os.path.basename(libs[0]), same), flush=True)
                    if not same:
                        describe_difference(ref, rec)


def describe_difference(ref, rec):
    """Which part of the records differs (counts, keypoints, descriptors, levels) and in how many frames."""
    from pislam_b200 import api
    lay = api.record_layout(CAP, WORDS)
    a, b = ref.reshape(F, lay["bytes"]), rec.reshape(F, lay["bytes"])
    n_a, n_b = a[:, :4].copy().view(np.uint32)[:, 0], b[:, :4].copy().view(np.uint32)[:, 0]
    parts = {"header": (0, lay["kp"]), "keypoints": (lay["kp"], lay["desc"]),
             "descriptors": (lay["desc"], lay["level"]), "levels": (lay["level"], lay["bytes"])}
    for name, (lo, hi) in parts.items():
        bad = int((a[:, lo:hi] != b[:, lo:hi]).any(axis=1).sum())
        print("      %-12s differs in %d of %d frames" % (name, bad, F), flush=True)
    print("      keypoints per frame: reference %.2f, this build %.2f; frames with another count: %d"
          % (n_a.mean(), n_b.mean(), int((n_a != n_b).sum())), flush=True)
    f = int(np.argmax((a != b).any(axis=1)))
    ka = a[f, lay["kp"]:lay["kp"] + 4 * int(n_a[f])].copy().view(np.uint32)
    kb = b[f, lay["kp"]:lay["kp"] + 4 * int(n_b[f])].copy().view(np.uint32)
    only_a, only_b = np.setdiff1d(ka, kb), np.setdiff1d(kb, ka)
    fmt = lambda v: ["(s%d x%d y%d)" % (e >> 24, (e >> 12) & 0xfff, e & 0xfff) for e in v[:6]]  # noqa: E731
    print("      first differing frame %d: only in reference %s, only in this build %s"
          % (f, fmt(only_a), fmt(only_b)), flush=True)


if __name__ == "__main__":
    main()
