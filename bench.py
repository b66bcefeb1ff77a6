#!/usr/bin/env python
"""Headline benchmark: ORB front-end throughput on the VGA 8-level 1.2x stacked pyramid.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--layout vga|fhd|uhd] [--impl ours|reference]

A "step" is one pass of the hot path (FAST-9 -> Harris -> NMS/buckets -> oriented rBRIEF, the caller loop of
reference demo/demo.cpp:77-101) over a batch of F synthetic stacked pyramids per GPU.  Prints ONE JSON line
(rank 0).  `value` is device-resident throughput in Mpix/s over level pixels; `e2e` is
the same metric through the C-ABI call with HOST (pinned) buffers, H2D of the stacks and D2H of the records
inside the timed region; `e2e_base_frames` is the same again starting from the BASE frames (the pyramid is built on
the GPU, SURVEY.md 8f row N1, so only 1/3 of the pixels cross PCIe).  `roofline` is for the dominant kernel
(front_tile_kernel) and `cpu_baseline` is the reference's own CPU implementation timed on this box's host cores.

`parity`: before the line is printed, the records of the timed steps (device-resident, e2e, base-frame) are compared
with the CPU oracle for every distinct bench frame, every copy of a frame with its first copy, and -- with N > 1 --
every rank's gathered buffer, slot by slot, with what the owning rank produced.  A mismatch prints the line with
"parity": {"ok": false, ...} and exits with status 2.

--impl reference runs only the CPU arm (oracle/_ref = the reference's own headers compiled through the NEON shim,
run natively over the same batch on all host threads; the C oracle port when that library is absent), same
`config`.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from pislam_b200 import pyramid as _pyr  # noqa: E402

METRIC = "orb_frontend_vga_pyramid_throughput"
UNIT = "Mpix/s"
BORDER, FAST_T, HARRIS_T, LOG_BUCKET, BUCKET_LIMIT, WORDS = 16, 20, 1 << 15, 4, 3, 8
# layout -> (pyramid.<name>, record capacity per frame, default frames per step per GPU)
LAYOUTS = {"vga": ("VGA", 2048, 256), "fhd": ("FHD", 4096, 64), "uhd": ("UHD", 8192, 16)}
LAY = _pyr.VGA  # rebound by --layout (the headline metric is quoted on VGA: BASELINE.json configs[1])
CAP = 2048


def n_distinct():
    return 32 if LAY is _pyr.VGA else 4  # distinct synthetic frames (large ones take seconds each); the batch cycles


def workload():
    name = {id(_pyr.VGA): "VGA", id(_pyr.FHD): "FHD", id(_pyr.UHD): "UHD"}[id(LAY)]
    return "%s 8-level 1.2x stacked pyramid (%dx%d), FAST 20 + Harris 2^15 + NMS buckets<4,3> + rBRIEF-256" % (
        name, LAY.rows, LAY.vstep)


def config(F):
    """Identical in both arms (the driver compares them): the workload and the batch one step processes per GPU."""
    return {"workload": workload(), "frames_per_step_per_gpu": F, "level_pixels_per_frame": LAY.level_pixels,
            "stack_bytes_per_frame": LAY.stack_bytes, "cap": CAP, "distinct_frames": min(F, n_distinct()),
            "l2_policy": "inputs larger than L2 (%d MB of stacks per GPU per step)" % (F * LAY.stack_bytes >> 20)
            if F * LAY.stack_bytes > (126 << 20) else
            "inputs smaller than L2 (%d MB): L2 flushed between steps is NOT done, quote only the default batch"
            % (F * LAY.stack_bytes >> 20)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_bases(n):
    nd = min(n, n_distinct())
    return [_pyr.synthetic_base(LAY.widths[0], LAY.heights[0], 0x5EED0000 + i) for i in range(nd)]


def frames_from_bases(bases, n):
    stacks = [_pyr.build_stack(b, LAY) for b in bases]
    return np.stack([stacks[i % len(stacks)] for i in range(n)])


# ---------------------------------------------------------------------------------------------- CPU arm
class CpuReference:
    """The reference's CPU implementation of the path on the host cores, one whole pyramid per task, run natively
    (no Python between the stages): oracle/ref_driver.cpp:ref_extract_batch over the reference's own headers when
    oracle/_ref is built, else the C oracle's po_extract_batch."""

    def __init__(self):
        import oracle_lib
        self.oracle = oracle_lib.Oracle()
        self.ref = None
        if oracle_lib.have_reference():
            try:
                self.ref = oracle_lib.Reference()
            except (OSError, AttributeError):
                self.ref = None
        self.kind = "reference" if self.ref is not None else "port"

    def run(self, stacks, nthreads):
        """Processes every frame once on `nthreads` host threads; returns (seconds, total keypoints)."""
        t0 = time.perf_counter()
        if self.ref is not None:
            counts = self.ref.extract_batch(stacks, LAY.widths, LAY.heights, LAY.row_offsets, nthreads,
                                            fast_threshold=FAST_T, harris_threshold=HARRIS_T, log_bucket=LOG_BUCKET,
                                            bucket_limit=BUCKET_LIMIT, cap=CAP * 4)
        else:
            counts = self.oracle.extract_batch(stacks, LAY.widths, LAY.heights, LAY.row_offsets, nthreads,
                                               border=BORDER, fast_threshold=FAST_T, harris_threshold=HARRIS_T,
                                               log_bucket=LOG_BUCKET, bucket_limit=BUCKET_LIMIT, words=WORDS,
                                               cap=CAP * 4)
        return time.perf_counter() - t0, int(counts.sum())


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline(cpu, stacks, cores):
    """All-cores and single-thread numbers on a bounded sample of the bench batch (the reference is single-threaded
    by design, demo/demo.cpp:75,103; frame-parallel threads are the fairest use of the whole host)."""
    cpu.run(stacks[:max(cores, 1)], cores)  # warm-up (page-in, thread start)
    n_all = len(stacks)
    dt, kp = cpu.run(stacks, cores)
    n1 = max(2, min(16, n_all))
    dt1, _ = cpu.run(stacks[:n1], 1)
    return {"value": LAY.level_pixels * n_all / dt / 1e6, "unit": UNIT, "cores": cores, "kind": cpu.kind,
            "keypoints_per_s": kp / dt,
            "sample": "%d frames (one step's batch) once on %d host threads, one pyramid per task" % (n_all, cores),
            "single_thread": {"value": LAY.level_pixels * n1 / dt1 / 1e6, "unit": UNIT, "cores": 1,
                              "ms_per_frame": 1e3 * dt1 / n1, "sample": "%d frames on 1 thread" % n1}}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cpu = CpuReference()
    cores = cpu_cores()
    F = args.frames
    stacks = frames_from_bases(make_bases(F), F)
    for _ in range(args.warmup):
        cpu.run(stacks[:max(cores, 1)], cores)
    times, kpts = [], 0
    for _ in range(args.steps):
        dt, kpts = cpu.run(stacks, cores)
        times.append(dt)
    total = sum(times)
    mpix = LAY.level_pixels * F * args.steps / total / 1e6
    n1 = max(2, min(16, F))
    dt1, _ = cpu.run(stacks[:n1], 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": mpix, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config(F),
        "keypoints_per_frame": kpts / F, "keypoints_per_s": kpts * args.steps / total,
        "cpu_baseline": {"value": mpix, "unit": UNIT, "cores": cores, "kind": cpu.kind,
                         "sample": "%d frames per step on %d host threads, one pyramid per task, %d steps" % (
                             F, cores, args.steps),
                         "single_thread": {"value": LAY.level_pixels * n1 / dt1 / 1e6, "unit": UNIT, "cores": 1,
                                           "ms_per_frame": 1e3 * dt1 / n1, "sample": "%d frames on 1 thread" % n1}},
        "e2e": {"value": mpix, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def emit(line):
    _OUT.write(json.dumps(line) + "\n")
    _OUT.flush()


_OUT = sys.stdout


# ---------------------------------------------------------------------------------------------- GPU arm
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
        self.lines = []
        if self.proc:
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [ln for (t, ln) in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or [ln for (_, ln) in self.lines]
        for ln in rows:
            f = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except (ValueError, IndexError):
                continue
            for k, nm in enumerate(names):
                if len(f) > 5 + k and f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


class ParityGate:
    """Compares records with the CPU oracle (tests/oracle_lib.py -- the checker, never the thing measured)."""

    def __init__(self, api, bases, F):
        import oracle_lib
        self.api, self.F = api, F
        oracle = oracle_lib.Oracle()
        self.want = []
        for b in bases:
            stack = oracle.build_pyramid(b, LAY)
            self.want.append(oracle.extract_pyramid(stack, LAY.widths, LAY.heights, LAY.row_offsets, border=BORDER,
                                                    fast_threshold=FAST_T, harris_threshold=HARRIS_T,
                                                    log_bucket=LOG_BUCKET, bucket_limit=BUCKET_LIMIT, words=WORDS))
        self.report = {"ok": True, "oracle_frames": len(self.want), "checks": {}}

    def check_records(self, name, buf):
        """buf: host uint8 copy of F records.  Every frame against the oracle result of its distinct source."""
        recs = self.api.parse_records(buf, self.F, CAP, WORDS)
        bad = 0
        for f, (kp, lv, desc, _, n_total) in enumerate(recs):
            okp, olv, odesc = self.want[f % len(self.want)]
            if not (n_total == len(okp) and np.array_equal(kp, okp) and np.array_equal(lv, olv) and
                    np.array_equal(desc, odesc)):
                bad += 1
        self.report["checks"][name] = {"frames": len(recs), "mismatches": bad}
        if bad:
            self.report["ok"] = False
        return recs

    @staticmethod
    def live_hashes(recs):
        """One digest per frame over the live part of its record (count, keypoints, levels, descriptors)."""
        out = []
        for kp, lv, desc, level_start, n_total in recs:
            h = hashlib.blake2b(digest_size=8)
            h.update(np.asarray([len(kp), n_total], np.uint32).tobytes())
            h.update(level_start.tobytes())
            h.update(kp.tobytes())
            h.update(lv.tobytes())
            h.update(desc.tobytes())
            out.append(h.hexdigest())
        return out

    def check_gathered(self, name, buf, world, rank, dist):
        """buf: host copy of this rank's gathered buffer (world x F records).  Slot r must hold exactly what rank r
        produced: per-frame digests of the live bytes, exchanged with all_gather_object."""
        recs = self.api.parse_records(buf, world * self.F, CAP, WORDS)
        mine = self.live_hashes(recs[rank * self.F:(rank + 1) * self.F])
        owners = [None] * world
        dist.all_gather_object(owners, mine)
        bad = 0
        for r in range(world):
            got = self.live_hashes(recs[r * self.F:(r + 1) * self.F])
            bad += sum(1 for a, b in zip(got, owners[r]) if a != b)
        # and the local slot itself against the oracle
        for f in range(self.F):
            okp, olv, odesc = self.want[f % len(self.want)]
            kp, lv, desc, _, n_total = recs[rank * self.F + f]
            if not (np.array_equal(kp, okp) and np.array_equal(desc, odesc)):
                bad += 1
        self.report["checks"][name] = {"slots": world * self.F, "mismatches": bad}
        if bad:
            self.report["ok"] = False


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from pislam_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    F = args.frames
    desc = api.make_desc(LAY, border=BORDER, fast_threshold=FAST_T, harris_threshold=HARRIS_T,
                         log_bucket=LOG_BUCKET, bucket_limit=BUCKET_LIMIT, words=WORDS)
    lay = api.record_layout(CAP, WORDS)
    rec_bytes = lay["bytes"]
    frame_bytes = LAY.stack_bytes
    W, H = LAY.widths[0], LAY.heights[0]

    bases = make_bases(F)  # every rank: its own F frames (weak scaling); same synthetic set per rank
    frames = frames_from_bases(bases, F)
    h_stacks = torch.empty((F, LAY.rows, LAY.vstep), dtype=torch.uint8, pin_memory=True)
    h_stacks.numpy()[:] = frames
    h_bases = torch.empty((F, H, W), dtype=torch.uint8, pin_memory=True)
    for f in range(F):
        h_bases.numpy()[f] = bases[f % len(bases)]
    h_records = torch.empty(F * rec_bytes, dtype=torch.uint8, pin_memory=True)
    d_stacks = h_stacks.to("cuda", non_blocking=False)
    # two gather buffers: the all-gather of step k runs under the kernels of step k+1
    nbufs = max(2, int(os.environ.get("PISLAM_BENCH_GATHER_BUFS", "2"))) if world > 1 else 1
    d_alls = [torch.zeros(world * F * rec_bytes, dtype=torch.uint8, device="cuda") for _ in range(nbufs)]
    d_all = d_alls[0]
    d_records = d_all[rank * F * rec_bytes:(rank + 1) * F * rec_bytes]
    turn = [0]

    ctx = api.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    gather = "none"
    if world > 1:
        uid = [api.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(world, rank, uid[0])
        p2p = all([ctx.comm_register(b.data_ptr(), b.numel()) for b in d_alls])
        gather = ("size-aware exchange: every rank stores the live part of its records straight into the peers' "
                  "gather buffers over NVLink peer mappings (p2p_push_kernel), on libpislam_b200's second stream"
                  if p2p else
                  "one ncclAllGather of the fixed-capacity records per step, issued by libpislam_b200 on its second "
                  "stream")
    gate = None if args.no_parity else ParityGate(api, bases, F)

    def step_device():
        """inputs resident in HBM; results (this rank's records, all-gathered when N>1) left in HBM"""
        if world > 1:
            buf = d_alls[turn[0]]
            turn[0] = (turn[0] + 1) % len(d_alls)
            st = ctx.extract_allgather_ptr(d_stacks.data_ptr(), frame_bytes, F, desc, CAP, buf.data_ptr())
        else:
            st = ctx.extract_batch_ptr(d_stacks.data_ptr(), frame_bytes, F, desc, CAP, d_records.data_ptr(),
                                       asynchronous=True)
        if st != 0:
            raise RuntimeError("extract failed: %d" % st)

    d2h_stream = torch.cuda.Stream() if world > 1 else None

    def e2e_multi(call):
        """N > 1: HOST frames in; the records are extracted into this rank's slot of a (registered) gather buffer through
        the library's chunked pipeline, exchanged over NVLink on the library's second stream, and the rank's own
        records are read back to the host on a third stream -- nothing in the step blocks, consecutive steps overlap"""
        buf = d_alls[turn[0]]
        turn[0] = (turn[0] + 1) % len(d_alls)
        st = call(buf.data_ptr())
        if st not in (0, api.PISLAM_ERR_CAPACITY):
            raise RuntimeError("extract failed: %d" % st)
        ev = torch.cuda.Event()
        ev.record(stream)  # this rank's records are complete here (the exchange runs on the library's second stream)
        d2h_stream.wait_event(ev)
        with torch.cuda.stream(d2h_stream):
            h_records.copy_(buf[rank * F * rec_bytes:(rank + 1) * F * rec_bytes], non_blocking=True)

    def e2e_single(st):
        if st not in (0, api.PISLAM_ERR_CAPACITY):
            raise RuntimeError("extract failed: %d" % st)

    def step_e2e(asynchronous=True):
        """the call a user makes: HOST stacks in, HOST records out (H2D + kernels + D2H inside the call).  The
        issue-only form (pislam_b200_extract_batch_async on pinned host buffers) lets consecutive calls overlap: the
        H2D of step k+1 runs under the kernels and the D2H of step k; the timed region ends after the last D2H."""
        if world == 1:
            e2e_single(ctx.extract_batch_ptr(h_stacks.data_ptr(), frame_bytes, F, desc, CAP, h_records.data_ptr(),
                                             asynchronous=asynchronous))
        else:
            e2e_multi(lambda all_ptr: ctx.extract_allgather_ptr(h_stacks.data_ptr(), frame_bytes, F, desc, CAP, all_ptr))
            if not asynchronous:
                ctx.synchronize()
                d2h_stream.synchronize()

    def step_e2e_blocking():
        step_e2e(False)

    def step_e2e_base(asynchronous=True):
        """HOST base frames in, HOST records out: the pyramid is built on the GPU"""
        if world == 1:
            e2e_single(ctx.extract_frames_ptr(h_bases.data_ptr(), W, W * H, F, desc, CAP, h_records.data_ptr(),
                                              asynchronous=asynchronous))
        else:
            e2e_multi(lambda all_ptr: ctx.extract_frames_allgather_ptr(h_bases.data_ptr(), W, W * H, F, desc, CAP,
                                                                       all_ptr))
            if not asynchronous:
                ctx.synchronize()
                d2h_stream.synchronize()

    def step_e2e_base_blocking():
        step_e2e_base(False)

    def barrier():
        if world > 1:
            dist.barrier()
        ctx.synchronize()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize; device time by CUDA events on the ctx stream."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.launch_count
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        ctx.join()  # the last exchange (second stream) and record copies belong to the timed region
        if d2h_stream is not None:
            stream.wait_stream(d2h_stream)
        e1.record(stream)
        barrier()
        t1 = time.perf_counter()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ctx.launch_count - l0, t0, t1

    W_ = max(args.warmup, 3)
    for _ in range(W_):
        step_device()
    sampler = ClockSampler(local) if rank == 0 else None
    ms_dev, launches, t0, t1 = timed(step_device, args.steps)
    # nvidia-smi samples every 100 ms and the timed region may be a few ms: keep the identical load running
    # (untimed, all ranks: the step contains a collective) until the sampler has seen >= 1 s of it
    t_load = t1
    extra_steps = 0
    if ms_dev < 1000.0:  # ms_dev is the max over ranks: the same number on every rank
        n_more = int(min(20000, (1000.0 - ms_dev) / max(ms_dev / args.steps, 1e-3))) + 1
        for _ in range(n_more):
            step_device()
        extra_steps = n_more
        barrier()
        t_load = time.perf_counter()
    clocks = sampler.stop(t0, t_load) if sampler else None
    if clocks is not None:
        clocks["window"] = ("timed region + %d identical untimed steps right after it (region %.1f ms < sampling "
                            "period)" % (extra_steps, (t1 - t0) * 1e3)) if extra_steps else "timed region"

    # keypoints actually produced (from the records of the last step) and the parity gate on them
    hdr = d_records.view(F, rec_bytes)[:, :8].contiguous().cpu().numpy().view(np.uint32)
    kpts_local = int(hdr[:, 0].sum())
    overflow = int((hdr[:, 1] > hdr[:, 0]).sum())
    if gate:
        if world > 1:
            for i, buf in enumerate(d_alls):
                gate.check_gathered("gathered_buffer_%d" % i, buf.cpu().numpy(), world, rank, dist)
        else:
            gate.check_records("device_resident", d_records.cpu().numpy())

    for _ in range(W_):
        step_e2e()
    # events on the ctx stream span the H2D copies, the kernels, the D2H copies and any host-side gaps
    h_records.zero_()
    ms_e2e, _, _, _ = timed(step_e2e, args.steps)
    if gate:
        gate.check_records("e2e_host_records", h_records.numpy())
    h_records.zero_()
    ms_e2e_blk, _, _, _ = timed(step_e2e_blocking, max(3, args.steps // 2))
    ms_e2e_blk *= args.steps / max(3, args.steps // 2)
    if gate:
        gate.check_records("e2e_blocking_host_records", h_records.numpy())

    for _ in range(W_):
        step_e2e_base()
    h_records.zero_()
    ms_e2e_base, launches_base, _, _ = timed(step_e2e_base, args.steps)
    if gate:
        gate.check_records("e2e_base_frames_host_records", h_records.numpy())
    ms_e2e_base_blk, _, _, _ = timed(step_e2e_base_blocking, max(3, args.steps // 2))
    ms_e2e_base_blk *= args.steps / max(3, args.steps // 2)

    # roofline of the dominant kernel: per-kernel CUDA events (library profiling mode), same inputs
    roof = latency = demo_ms = base_dev = None
    if rank == 0:
        roof = kernel_roofline(ctx, api, d_stacks, frame_bytes, F, desc, d_records, kpts_local, args.steps)
        base_dev = base_frames_device(ctx, api, torch, h_bases, W, H, F, desc, d_records, args.steps, stream)
        latency = single_frame_latency(ctx, h_stacks[0].numpy(), desc)  # pinned host frame
        if LAY is _pyr.VGA:
            demo_ms = literal_demo_loop_ms(ctx, h_stacks[0].numpy())

    if world > 1:
        t = torch.tensor([kpts_local, overflow, 0 if (gate is None or gate.report["ok"]) else 1], device="cuda",
                         dtype=torch.int64)
        dist.all_reduce(t)
        kpts_total, overflow, parity_bad_ranks = int(t[0].item()), int(t[1].item()), int(t[2].item())
    else:
        kpts_total = kpts_local
        parity_bad_ranks = 0 if (gate is None or gate.report["ok"]) else 1

    ok = True
    if rank == 0:
        pix = LAY.level_pixels * F * world
        value = pix * args.steps / (ms_dev * 1e-3) / 1e6
        e2e_value = pix * args.steps / (ms_e2e * 1e-3) / 1e6
        e2e_base_value = pix * args.steps / (ms_e2e_base * 1e-3) / 1e6
        cpu_line = None
        if not args.no_cpu:
            cpu = CpuReference()
            cpu_line = cpu_baseline(cpu, frames, cpu_cores())
        parity = None
        if gate:
            parity = gate.report
            parity["ranks_with_mismatches"] = parity_bad_ranks
            parity["ok"] = parity["ok"] and parity_bad_ranks == 0
            ok = parity["ok"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": W_, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config(F),
            "keypoints_per_frame": kpts_total / (F * world),
            "keypoints_per_s": kpts_total * args.steps / (ms_dev * 1e-3),
            "details": {"capacity_overflows": overflow, "gather": gather,
                        "parallelism": "frames sharded one batch per GPU, one all-gather of records per step "
                                       "(overlapping the next step's kernels)" if world > 1 else "single GPU",
                        "numa_node_rank0": numa},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": F * LAY.level_pixels,
                    "h2d_note": "level pixels only: the padding right of each level is not uploaded",
                    "d2h_bytes_per_step": F * rec_bytes, "ms_per_step": ms_e2e / args.steps,
                    "keypoints_per_s": kpts_total * args.steps / (ms_e2e * 1e-3),
                    "call": "pislam_b200_extract_batch_async on pinned host buffers, steps issued back to back "
                            "(the copies of consecutive calls overlap), timed until the last D2H has landed",
                    "blocking_call": {"value": pix * args.steps / (ms_e2e_blk * 1e-3) / 1e6, "unit": UNIT,
                                      "ms_per_step": ms_e2e_blk / args.steps,
                                      "call": "pislam_b200_extract_batch, one blocking call per step"}},
            "e2e_base_frames": {"value": e2e_base_value, "unit": UNIT, "h2d_bytes_per_step": F * W * H,
                                "d2h_bytes_per_step": F * rec_bytes, "ms_per_step": ms_e2e_base / args.steps,
                                "note": "host BASE frames in, host records out: pislam_b200_extract_frames builds "
                                        "the 8-level pyramid on the GPU; Mpix/s over the same level pixels",
                                "gpu_launches": launches_base, "device_resident": base_dev,
                                "call": "pislam_b200_extract_frames_async, steps issued back to back",
                                "blocking_call": {"value": pix * args.steps / (ms_e2e_base_blk * 1e-3) / 1e6,
                                                  "unit": UNIT, "ms_per_step": ms_e2e_base_blk / args.steps}},
            "gpu_launches": launches,
            "single_frame_latency_us": latency,
            "literal_stage_demo_loop_ms_per_frame": demo_ms,
            "parity": parity,
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu_line,
        }
        emit(line)
    # teardown order matters: torch may still hold blocks it used on the ctx stream (an ExternalStream);
    # release them before the stream is destroyed with the context
    if world > 1:
        dist.barrier()
    barrier()
    del d_records, d_all, d_alls, d_stacks, h_stacks, h_records, h_bases, stream
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    if world > 1:
        try:
            ctx.comm_destroy()
        except Exception:  # noqa: BLE001
            pass
        dist.destroy_process_group()
    ctx.close()
    if parity_bad_ranks:
        sys.stderr.write("bench.py: PARITY MISMATCH against the oracle -- the numbers above are void\n")
        sys.exit(2)
    return ok


def single_frame_latency(ctx, stack, desc, iters=200):
    """One pyramid at a time through pislam_b200_extract_pyramid (pinned host stack in, host keypoints +
    descriptors out, blocking): the latency a per-frame caller such as the reference demo loop sees.  Wall clock,
    mean."""
    stack = np.ascontiguousarray(stack)
    for _ in range(10):
        ctx.extract_pyramid(stack, desc, cap=CAP)
    t0 = time.perf_counter()
    for _ in range(iters):
        ctx.extract_pyramid(stack, desc, cap=CAP)
    return (time.perf_counter() - t0) / iters * 1e6


def literal_demo_loop_ms(ctx, stack, iters=5):
    """The reference demo's own loop (demo/demo.cpp:77-101) through the literal stage calls -- what a user who only
    swaps the include path gets: per level fastDetect -> fastScoreHarris -> fastExtract (img and out round-trip
    through the device on every call), then one orbCompute.  examples/demo (C++, same loop against
    include/pislam/*.h) when it has been built, else the ctypes mirror.  ms per frame, wall clock."""
    exe = os.path.join(ROOT, "examples", "demo")
    if os.path.exists(exe):
        try:
            import tempfile
            with tempfile.NamedTemporaryFile(suffix=".raw") as f:
                f.write(np.ascontiguousarray(stack).tobytes())
                f.flush()
                best = None
                for _ in range(3):
                    out = subprocess.run([exe, f.name], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
                                         timeout=120).stdout
                    for ln in out.splitlines():
                        if ln.startswith("stages:"):
                            ms = float(ln.split("descriptors", 1)[1].split(",", 1)[1].split()[0])
                            best = ms if best is None else min(best, ms)
                if best is not None:
                    return {"ms": best, "how": "examples/demo (C++ against include/pislam/*.h): third pass of the "
                                               "loop in one process (context and buffers warm), best of 3 processes"}
        except Exception:  # noqa: BLE001
            pass
    out = np.zeros_like(stack)

    def once():
        out[:] = 0
        pts = []
        for w, h, r in zip(LAY.widths, LAY.heights, LAY.row_offsets):
            ctx.fastDetect(w, h, stack, out, FAST_T, border=BORDER, row0=r)
            ctx.fastScoreHarris(w, h, stack, HARRIS_T, out, border=BORDER, row0=r)
            kp = ctx.fastExtract(w, h, out, border=BORDER, row0=r)
            pts.append(kp + np.uint32(r))
        ctx.orbCompute(stack, np.concatenate(pts), words=WORDS)
    once()
    t0 = time.perf_counter()
    for _ in range(iters):
        once()
    return {"ms": (time.perf_counter() - t0) / iters * 1e3, "how": "ctypes mirror of the literal stage calls"}


def base_frames_device(ctx, api, torch, h_bases, W, H, F, desc, d_records, steps, stream):
    """Device-resident base frames -> records (pyramid_build_kernel + the fused path in L2-sized chunks)."""
    d_bases = h_bases.to("cuda")
    for _ in range(3):
        assert ctx.extract_frames_ptr(d_bases.data_ptr(), W, W * H, F, desc, CAP, d_records.data_ptr()) == 0
    ctx.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        assert ctx.extract_frames_ptr(d_bases.data_ptr(), W, W * H, F, desc, CAP, d_records.data_ptr()) == 0
    e1.record(stream)
    ms, cnt = ctx.kernel_times()
    ctx.set_profiling(False)
    total = e0.elapsed_time(e1)
    del d_bases
    return {"ms_per_step": total / steps, "value": LAY.level_pixels * F * steps / (total * 1e-3) / 1e6, "unit": UNIT,
            "pyramid_build_kernel_ms_per_step": ms[3] / steps, "note": "blocking call per step, profiling events on"}


def kernel_roofline(ctx, api, d_stacks, frame_bytes, F, desc, d_records, kpts, steps):
    """Average duration of front_tile_kernel from CUDA events recorded around each of its launches on the
    ctx stream (pislam_b200_set_profiling), against its algorithmic bytes:
    every level pixel read once + 4 B per keypoint written."""
    ctx.set_profiling(True)
    for _ in range(steps):
        st = ctx.extract_batch_ptr(d_stacks.data_ptr(), frame_bytes, F, desc, CAP, d_records.data_ptr(),
                                   asynchronous=True)
        assert st == 0
    ms, cnt = ctx.kernel_times()
    ctx.set_profiling(False)
    peak, peak_src = measured_peak_gbs()
    if cnt[0] == 0:
        return None
    t_front = ms[0] / cnt[0] * 1e-3
    alg_bytes = LAY.level_pixels * F + 4 * kpts
    achieved = alg_bytes / t_front / 1e9
    traffic, traffic_src = None, None
    try:  # DRAM bytes of one launch from the committed ncu --set full capture (same workload and batch size)
        with open(os.path.join(ROOT, "profiles", "front_kernel_traffic.json")) as f:
            tr = json.load(f)
        if tr["frames_per_launch"] == F and LAY is _pyr.VGA:
            traffic, traffic_src = tr["dram_bytes_read"] + tr["dram_bytes_write"], tr["source"]
    except Exception:  # noqa: BLE001
        pass
    return {"kernel": "front_tile_kernel<TMA>", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": t_front * 1e3,
            "launches_timed": int(cnt[0]),
            "other_kernels_avg_ms": {"order_kernel": ms[1] / max(cnt[1], 1), "orb_kernel": ms[2] / max(cnt[2], 1)}}


def bind_to_gpu_numa(local_rank):
    """Pin this process (and so its first-touch pinned buffers) to the CPUs of the GPU's NUMA node: sysfs first, the
    CPU-affinity column of `nvidia-smi topo -m` when the container hides the node."""
    def apply(cpus):
        cpus = set(cpus) & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return True
        return False

    def parse_list(s):
        cpus = set()
        for part in s.strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        return cpus
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node >= 0 and apply(parse_list(open("/sys/devices/system/node/node%d/cpulist" % node).read())):
            return node
    except Exception:  # noqa: BLE001
        pass
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
                             timeout=20).stdout
        import re
        clean = re.sub(r"\x1b\[[0-9;]*m", "", out)
        head = None
        for ln in clean.splitlines():
            cols = [c.strip() for c in ln.split("\t") if c.strip()]
            if not cols:
                continue
            if head is None and any("CPU Affinity" in c for c in cols):
                head = cols
                continue
            if head and cols[0] == "GPU%d" % local_rank:
                # the header lacks the row-label column: values align with head shifted by one
                idx = next(i for i, c in enumerate(head) if "CPU Affinity" in c) + 1
                nidx = next((i for i, c in enumerate(head) if "NUMA Affinity" in c), None)
                if idx < len(cols) and apply(parse_list(cols[idx])):
                    try:
                        return int(cols[nidx + 1]) if nidx is not None else "cpus " + cols[idx]
                    except (ValueError, IndexError):
                        return "cpus " + cols[idx]
    except Exception:  # noqa: BLE001
        pass
    return None


def main():
    # the contract is ONE JSON line on stdout: libraries that chat on fd 1 (NCCL prints its version there) are
    # sent to stderr, and the line goes to the saved descriptor
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=0, help="stacked pyramids per step per GPU (default: 256 VGA, "
                                                          "64 1080p, 16 4K)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--layout", default="vga", choices=sorted(LAYOUTS),
                    help="pyramid geometry (default vga = the headline config; fhd/uhd = BASELINE configs 3 and 5)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (sweeps)")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-bench parity gate (never for a quoted number)")
    args = ap.parse_args()
    global LAY, CAP
    name, CAP, default_frames = LAYOUTS[args.layout]
    LAY = getattr(_pyr, name)
    if args.frames <= 0:
        args.frames = default_frames
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
