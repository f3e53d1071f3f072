#!/usr/bin/env python
"""bench.py -- stereo frames/sec of the circularMatching() hot path on B200 (+ the CPU reference arm).

  python bench.py --gpus N --steps K --warmup W            this library (one process per GPU)
  python bench.py --impl reference --gpus N --steps K ...  the reference's OpenCV CPU path on the host cores

One "step" = one pass of the whole path (FAST -> select 2000 -> pyramids -> LK ring -> filters ->
triangulation -> PnP/RANSAC) over `--units` independent KITTI-shaped synthetic stereo pairs per GPU.
Prints ONE JSON line (see README "bench contract"):
  value      frames/s, inputs resident in HBM when the timed region starts (vo_batch_run only)
  e2e        frames/s through the C-ABI with pinned HOST buffers: H2D of the 4 images per unit +
             run + D2H of the result records inside the timed region (vo_frame_batch)
  roofline   LK ring kernel: algorithmic bytes (17044 B per feature-ring) / its own
             CUDA-event time on the launching stream, vs the measured HBM copy bandwidth
  cpu_baseline  cv2 (the OpenCV the reference links) through the reference glue, timed on this host
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W_IMG, H_IMG, N_FEAT = 1241, 376, 2000
LK_BYTES_PER_FEATURE = 4 * (4 * ((21 + 3) ** 2 + (21 + 1) ** 2) + 21)      # 17044, SURVEY.md 8(d)
METRIC = "stereo frames/sec at 1241x376, 2000 feats; LK kernel HBM GB/s vs roofline"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--units", type=int, default=8, help="independent stereo pairs per step per GPU")
    ap.add_argument("--features", type=int, default=N_FEAT)
    ap.add_argument("--cpu-sample", type=int, default=400, help="frames timed for cpu_baseline (~10 s of host work)")
    ap.add_argument("--sequence", type=int, default=48,
                    help="frames of the streaming-mode (vo_seq_push) side measurement at N=1; 0 = skip")
    ap.add_argument("--width", type=int, default=W_IMG)
    ap.add_argument("--height", type=int, default=H_IMG)
    ap.add_argument("--calib", default="kitti", choices=["kitti", "zed"], help="intrinsics of the synthetic rig")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.perf_counter() + 0.5          # a very short timed region: wait for the first sample (clocks are still up)
        while not self.lines and time.perf_counter() < t_end:
            time.sleep(0.01)
        self.proc.terminate()          # exact PID we started
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[2 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def _cpu_one_frame(u, n_feat):
    """One work unit through the reference's CPU path: cv2 (the OpenCV the reference's calls resolve to) behind the
    verbatim glue of oracle/ref_path.py -- FAST, stride selection, 4-call LK ring, filters, triangulation, PnP."""
    from oracle import ref_path
    from visual_odom_b200 import synth
    corners = ref_path.fast_cv2(u["l0"])
    pts = synth.select_features(corners, n_feat)
    fs = ref_path.FeatureSet(); fs.points = pts; fs.ages = np.zeros(len(pts), np.int32)
    cm = ref_path.circular_matching(u["l0"], u["r0"], u["l1"], u["r1"], pts, fs, "cv2")
    ok = ref_path.check_valid_match(cm["l0"], cm["l0_ret"], 0)
    pL0, pR0, pL1 = (ref_path.remove_invalid_points(cm[k], ok) for k in ("l0", "r0", "l1"))
    X = ref_path.triangulate(u["P_l"], u["P_r"], pL0, pR0, "cv2")
    return ref_path.tracking_frame2frame(u["P_l"], pL0, pL1, X, np.array([0.0, 0.0, -0.8]), "cv2")


def cpu_reference_frames(units, n_feat, frames, threads=None):
    """Sequential frames, OpenCV's own thread pool inside each call (how the reference program runs)."""
    import cv2
    if threads is not None:
        cv2.setNumThreads(threads)
    for i in range(min(3, len(units))):
        _cpu_one_frame(units[i], n_feat)                      # warm-up
    t0 = time.perf_counter()
    for i in range(frames):
        _cpu_one_frame(units[i % len(units)], n_feat)
    dt = time.perf_counter() - t0
    return frames / dt, dt, cv2.getNumThreads()


_POOL_STATE = {}


def _pool_init(w, h, calib, n_feat, threads):
    """Runs once in every worker process: its own work unit + OpenCV thread count + one warm-up frame."""
    import cv2
    from visual_odom_b200 import synth
    cv2.setNumThreads(threads)
    cal = synth.KITTI00 if calib == "kitti" else synth.ZED
    _POOL_STATE["unit"] = synth.stereo_unit(w, h, os.getpid() % 64, cal=cal)
    _POOL_STATE["n_feat"] = n_feat
    _cpu_one_frame(_POOL_STATE["unit"], n_feat)


def _pool_run(reps):
    for _ in range(reps):
        _cpu_one_frame(_POOL_STATE["unit"], _POOL_STATE["n_feat"])
    return reps


def cpu_reference_parallel(n_proc, w, h, calib, n_feat, frames, threads_per_proc):
    """Independent work units on `n_proc` host processes at once (each with `threads_per_proc` OpenCV threads): what a
    CPU deployment of the batched workload would do with all the cores.  Returns (frames/s, wall seconds, frames)."""
    import multiprocessing as mp
    ctxm = mp.get_context("spawn")                 # no fork: OpenCV's thread pool does not survive one
    with ctxm.Pool(n_proc, initializer=_pool_init, initargs=(w, h, calib, n_feat, threads_per_proc)) as pool:
        pool.map(_pool_run, [1] * (2 * n_proc), chunksize=1)          # every worker initialised and warm
        chunk = 2
        tasks = max(n_proc, int(round(frames / chunk)))
        tasks = ((tasks + n_proc - 1) // n_proc) * n_proc               # whole waves
        t0 = time.perf_counter()
        done = sum(pool.map(_pool_run, [chunk] * tasks, chunksize=1))
        dt = time.perf_counter() - t0
    return done / dt, dt, done


def cpu_reference_best(units, args, frames):
    """The better of (a) sequential frames with OpenCV's internal threads and (b) one process per core group."""
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    fps_a, dt_a, cv_threads = cpu_reference_frames(units, args.features, frames)
    best = {"value": fps_a, "cores": cv_threads, "seconds": dt_a, "frames": frames,
            "how": f"sequential frames, {cv_threads} OpenCV threads (cv2 default)"}
    tried = [f"sequential x{cv_threads} threads: {fps_a:.1f} fps"]
    try:
        n_proc = max(1, min(cores, 64))
        per = max(1, cores // n_proc)
        fps_b, dt_b, done = cpu_reference_parallel(n_proc, W_IMG, H_IMG, args.calib, args.features, max(frames, 4 * n_proc), per)
        tried.append(f"{n_proc} processes x{per} threads: {fps_b:.1f} fps")
        if fps_b > best["value"]:
            best = {"value": fps_b, "cores": n_proc * per, "seconds": dt_b, "frames": done,
                    "how": f"{n_proc} processes x {per} OpenCV thread(s), independent work units, {done} frames"}
    except Exception as e:                       # never lose the line to the pool
        tried.append(f"process pool failed: {str(e)[:80]}")
    best["tried"] = tried
    best["affinity_cores"] = cores
    return best


def sequence_mode(ctx, torch, cal, n_frames):
    """SURVEY.md 8f row N1, reported beside the headline: the reference's actual usage pattern -- one stereo pair
    at a time through vo_seq_push (pinned host images in, pose out, main-loop state resident on the GPU) -- against
    the same loop on the CPU (cv2 through oracle/ref_path.matching_features).  Latency-bound: ~300 bucketed
    features per frame, one frame in flight."""
    from oracle import ref_path
    from visual_odom_b200 import synth
    step_r = np.array([0.001, -0.004, 0.0005]); step_t = np.array([0.01, -0.003, -0.2])
    base = synth.stereo_unit(W_IMG, H_IMG, 31, cal=cal)
    frames = [(base["l0"], base["r0"])]
    for k in range(1, n_frames + 1):
        u = synth.stereo_unit(W_IMG, H_IMG, 31, rvec=step_r * k, tvec=step_t * k, cal=cal)
        frames.append((u["l1"], u["r1"]))
    pin = []
    for l, r in frames:
        a = torch.empty((2, H_IMG, W_IMG), dtype=torch.uint8, pin_memory=True)
        a.numpy()[0] = l; a.numpy()[1] = r
        pin.append(a.numpy())
    lat = []
    for rep in range(2):                      # rep 0 warms up (graph capture), rep 1 is timed
        ctx.seq_begin(pin[0][0], pin[0][1], base["P_l"], base["P_r"])
        lat = []
        t0 = time.perf_counter()
        for k in range(1, n_frames + 1):
            t1 = time.perf_counter()
            got = ctx.seq_push(pin[k][0], pin[k][1], want_points=False)
            lat.append(time.perf_counter() - t1)
        dt_sync = time.perf_counter() - t0
    # pipelined: frame k+1 submitted before frame k is waited for (one frame of result lag)
    for rep in range(2):
        ctx.seq_begin(pin[0][0], pin[0][1], base["P_l"], base["P_r"])
        t0 = time.perf_counter()
        ctx.seq_submit(pin[1][0], pin[1][1])
        for k in range(1, n_frames + 1):
            if k + 1 <= n_frames:
                ctx.seq_submit(pin[k + 1][0], pin[k + 1][1])
            got_p = ctx.seq_wait(want_points=False)
        dt = time.perf_counter() - t0
    assert got_p["n_inliers"] == got["n_inliers"] and np.array_equal(got_p["tvec"], got["tvec"])
    gpu_fps = n_frames / dt
    # row N3: the same frames from a KITTI-layout PNG directory: vo_reader (worker threads decode ahead into a pinned
    # ring) -> vo_seq_submit / vo_seq_wait, against cv2.imread + cvtColor + the cv2 loop below
    png = None
    try:
        import cv2, shutil, tempfile
        from visual_odom_b200 import capi
        d = tempfile.mkdtemp(prefix="vo_png_")
        os.makedirs(os.path.join(d, "image_0")); os.makedirs(os.path.join(d, "image_1"))
        for i, (l, r) in enumerate(frames):
            cv2.imwrite(os.path.join(d, "image_0", "%06d.png" % i), l)
            cv2.imwrite(os.path.join(d, "image_1", "%06d.png" % i), r)
        threads = max(2, min(16, len(os.sched_getaffinity(0)) // 2))
        for rep in range(2):
            rd = capi.SequenceReader(d, 0, n_frames + 1, threads=threads, depth=threads + 3)
            t0 = time.perf_counter()
            lp, rp, rw, rh, rpitch, ch, fid = rd.next_ptr()
            ctx.seq_begin_ptr(rw, rh, lp, rp, rpitch, base["P_l"], base["P_r"], ch)
            lp, rp, rw, rh, rpitch, ch, fid = rd.next_ptr()
            ctx.seq_submit_ptr(lp, rp, rpitch, ch)
            for k in range(1, n_frames + 1):
                if k + 1 <= n_frames:
                    lp, rp, rw, rh, rpitch, ch, fid = rd.next_ptr()
                    ctx.seq_submit_ptr(lp, rp, rpitch, ch)
                got_f = ctx.seq_wait(want_points=False)
            dt_png = time.perf_counter() - t0
            rd.close()
        t0 = time.perf_counter()
        for i in range(min(n_frames + 1, 13)):
            for cam in (0, 1):
                cv2.cvtColor(cv2.imread(os.path.join(d, "image_%d" % cam, "%06d.png" % i), cv2.IMREAD_COLOR), cv2.COLOR_BGR2GRAY)
        cpu_load_ms = 1e3 * (time.perf_counter() - t0) / min(n_frames + 1, 13)
        shutil.rmtree(d, ignore_errors=True)
        png = {"value": n_frames / dt_png, "unit": "frames/s", "decode_threads": threads, "ring_depth": threads + 3,
               "same_result_as_memory_path": bool(got_f["n_inliers"] == got["n_inliers"] and np.array_equal(got_f["tvec"], got["tvec"])),
               "cpu_imread_cvtcolor_ms_per_frame_pair": cpu_load_ms,
               "note": "PNG files -> vo_reader (decode ahead, pinned ring) -> vo_seq_submit / vo_seq_wait; includes vo_seq_begin"}
    except Exception as e:
        png = {"error": str(e)[:200]}
    pose = ctx.seq_pose()
    # CPU: same loop, bounded sample
    fs = ref_path.FeatureSet(); translation = np.zeros(3)
    ncpu = min(n_frames, 12)
    t0 = time.perf_counter()
    for k in range(1, ncpu + 1):
        l0, r0 = frames[k - 1]; l1, r1 = frames[k]
        pL0, pR0, pL1, pR1, info = ref_path.matching_features(l0, r0, l1, r1, fs, backend="cv2")
        X = ref_path.triangulate(base["P_l"], base["P_r"], pL0, pR0, "cv2")
        R, translation, inl, rvec = ref_path.tracking_frame2frame(base["P_l"], pL0, pL1, X, translation, "cv2")
    cpu_fps = ncpu / (time.perf_counter() - t0)
    return {"value": gpu_fps, "unit": "frames/s", "frames": n_frames, "synchronous_fps": n_frames / dt_sync,
            "median_latency_ms": 1e3 * float(np.median(lat)),
            "max_latency_ms": 1e3 * float(np.max(lat)), "cpu_reference": cpu_fps, "cpu_frames": ncpu,
            "features_last_frame": int(got["n_features"]), "inliers_last_frame": int(got["n_inliers"]),
            "pose_translation": [float(x) for x in pose[:3, 3]], "from_png": png,
            "note": "value: vo_seq_submit / vo_seq_wait with two frames in flight (wall clock incl. H2D of every new pair and "
                    "the pose read-back); synchronous_fps / latency: one vo_seq_push at a time"}


def run_reference(args, rank, world):
    """--impl reference: the reference's own OpenCV CPU implementation on the host cores."""
    if rank != 0:
        return
    from visual_odom_b200 import synth
    global W_IMG, H_IMG
    W_IMG, H_IMG = args.width, args.height
    from visual_odom_b200 import synth as _s
    cal = _s.KITTI00 if args.calib == "kitti" else _s.ZED
    units = [synth.stereo_unit(W_IMG, H_IMG, s, cal=cal) for s in range(args.units)]
    # each step = a bounded sample of the workload (args.units frames per process group) on the CPU
    for _ in range(min(args.warmup, 1)):
        cpu_reference_frames(units, args.features, len(units))
    frames_total = max(len(units), min(args.steps * len(units), 600))
    best = cpu_reference_best(units, args, frames_total)
    cores = best["cores"]
    t_total = args.steps * len(units) / best["value"]
    value = args.steps * len(units) / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/i32 fixed point + f32 (LK), f64 (pose)", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{best['frames']} frames of the workload in {best['seconds']:.1f} s; {best['how']}; cv2 "
                                   f"{__import__('cv2').__version__} (the OpenCV build the reference's calls resolve to) through the "
                                   f"oracle/ref_path.py glue restatement; affinity cores={best['affinity_cores']}; tried: {best['tried']}"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, world):
    return {"workload": f"{args.calib}-calibrated synthetic stereo {W_IMG}x{H_IMG}, {args.features} FAST features (thr 20, even-stride "
                        f"selection), LK 21x21 maxLevel=3 (4 images) 30 it / 0.01, PnP RANSAC 500/0.5/0.999; "
                        f"{args.units} independent stereo pairs per step per GPU",
            "units_per_gpu": args.units, "global_units": args.units * world, "features": args.features,
            "l2": "flushed before every step's submission by a 256 MiB write (inside the timed region)", "parallelism": f"units sharded over {world} GPU(s), no data-path collective"}


# ------------------------------------------------------------------------------------------------
def main():
    global W_IMG, H_IMG
    args = parse()
    W_IMG, H_IMG = args.width, args.height
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from visual_odom_b200 import synth
    from visual_odom_b200.capi import Context

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: this library has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from visual_odom_b200 import dist as vd
    B = args.units
    # work-queue scatter: rank 0 owns the unit table (seeds), broadcast over NCCL; unit u -> rank u mod world
    table = vd.broadcast_unit_table(np.arange(world * B) if rank == 0 else np.zeros(world * B, np.int64), device="cuda")
    my_units = vd.unit_assignment(world * B, world)[rank]
    seeds = [int(table[u]) for u in my_units]
    cal = synth.KITTI00 if args.calib == "kitti" else synth.ZED
    units = [synth.stereo_unit(W_IMG, H_IMG, s, cal=cal) for s in seeds]

    # pinned host copies of the images (what a capture / decode thread would hand over)
    pinned = []
    for u in units:
        d = {}
        for k in ("l0", "r0", "l1", "r1"):
            t = torch.empty((H_IMG, W_IMG), dtype=torch.uint8, pin_memory=True)
            t.numpy()[:] = u[k]
            d[k] = t.numpy()
        pinned.append(d)
    keep_alive = pinned

    ctx = Context(local_rank, max_features=max(2048, args.features), max_units=2 * B)
    # a real (non-default) stream shared by torch's events and the library's kernels
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    for opt in ("graphs", "priorities", "batch_graphs"):   # A/B switches for experiments, e.g. VO_OPT_PRIORITIES=0 VO_OPT_BATCH_GRAPHS=1
        if os.environ.get("VO_OPT_" + opt.upper()) is not None:
            ctx.set_option(opt, float(os.environ["VO_OPT_" + opt.upper()]))
    ctx.batch_configure(W_IMG, H_IMG, 2 * B, units[0]["P_l"], units[0]["P_r"])     # two slot ranges of B (pipelined e2e)
    arr, keep, pitch = ctx.make_units([dict(p, n_select=args.features, t_prev=(0.0, 0.0, -0.8)) for p in pinned])

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- resident-input throughput (`value`) ----------------
    # Both slot ranges (2 x B units) are uploaded once; a step re-runs one resident range of B units
    # (vo_batch_submit with units = NULL) and reads its B result records back.  Two steps are in flight, as in `e2e`;
    # the L2 is flushed by a 256 MiB write before every submission (on the caller's stream, so the submission waits
    # for it; the flush is INSIDE the timed region).
    arr2, keep2, _ = ctx.make_units([dict(p, n_select=args.features, t_prev=(0.0, 0.0, -0.8)) for p in pinned + pinned])
    ctx.batch_upload(arr2, pitch)

    def resident_steps(n, ev_pair=None):
        out = None
        if ev_pair:
            ev_pair[0].record(stream)
        flush.fill_(1)
        ctx.batch_submit(None, 0, pitch, n_units=B)
        for s in range(n):
            if s + 1 < n:
                flush.fill_(s & 0xFF)
                ctx.batch_submit(None, ((s + 1) & 1) * B, pitch, n_units=B)
            out = ctx.batch_wait((s & 1) * B, B)
        if ev_pair:
            ev_pair[1].record(stream)
        return out

    resident_steps(max(2, args.warmup))
    torch.cuda.synchronize()
    ctx.lk_kernel_time(reset=True)
    launches0 = ctx.kernel_launches()
    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    res = resident_steps(args.steps, ev)
    barrier()
    t_dev_ms = ev[0].elapsed_time(ev[1])
    launches = ctx.kernel_launches() - launches0
    feats_per_launch = sum(r["n_features"] for r in res)

    ctx.batch_upload(arr, pitch)          # back to one resident range of B units for the single-stream pass
    # ---------------- LK kernel alone (roofline): one stream, so its CUDA-event time is not shared ----------------
    ctx.set_option("batch_streams", 1)
    ctx.set_option("graphs", 0)          # plain launches: the LK kernel is bracketed by its own CUDA events
    for _ in range(args.warmup):
        ctx.batch_run()
    torch.cuda.synchronize()
    ctx.lk_kernel_time(reset=True)
    ev1 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for s in range(args.steps):
        flush.fill_(s & 0xFF)
        ev1[s][0].record(stream)
        ctx.batch_run()
        ev1[s][1].record(stream)
    torch.cuda.synchronize()
    t_single_ms = sum(a.elapsed_time(b) for a, b in ev1)
    lk_ms, lk_n = ctx.lk_kernel_time(reset=True)
    ctx.set_option("batch_streams", 2)
    ctx.set_option("graphs", 1)

    # ---------------- end-to-end through the C-ABI with host buffers (`e2e`) ----------------
    # Pipelined submissions (vo_batch_submit / vo_batch_wait): every step uploads its B stereo pair-of-pairs from pinned
    # host memory into one of two resident slot ranges, runs the whole path and reads its B result records back; step
    # s+1 is submitted before step s is waited for, so the copy and the latency-bound PnP tail of one step run under
    # the LK ring of the other.  Every step's H2D, kernels and D2H are inside the timed region.
    def e2e_steps(n, ev_pair=None):
        out = None
        if ev_pair:
            ev_pair[0].record(stream)
        ctx.batch_submit(arr, 0, pitch)
        for s in range(n):
            if s + 1 < n:
                ctx.batch_submit(arr, ((s + 1) & 1) * B, pitch)
            out = ctx.batch_wait((s & 1) * B, B)
            if world > 1:                                    # result gather: fixed-size records over NCCL
                vd.gather_records([vd.result_to_record(r) for r in out], my_units, world * B, device="cuda")
        if ev_pair:
            ev_pair[1].record(stream)
        return out

    e2e_steps(max(2, args.warmup))
    torch.cuda.synchronize()
    barrier()
    e2e_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    t_wall0 = time.perf_counter()
    res_e2e = e2e_steps(args.steps, e2e_ev)
    barrier()
    t_e2e_wall = time.perf_counter() - t_wall0
    t_e2e_ms = e2e_ev[0].elapsed_time(e2e_ev[1])
    clocks = sampler.stop() if sampler else None

    # max over ranks (device-timed)
    tt = torch.tensor([t_dev_ms, t_e2e_ms, lk_ms, float(feats_per_launch)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = tt.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tt.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    else:
        tmax = tt; tsum = tt
    t_dev_ms, t_e2e_ms = float(tmax[0]), float(tmax[1])

    if rank == 0:
        frames = world * B * args.steps
        value = frames / (t_dev_ms * 1e-3)
        e2e_value = frames / (t_e2e_ms * 1e-3)
        peak, peak_src = peaks()
        lk_avg_ms = lk_ms / max(lk_n, 1)
        alg_bytes = LK_BYTES_PER_FEATURE * feats_per_launch
        achieved = alg_bytes / (lk_avg_ms * 1e-3) / 1e9 if lk_avg_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "lk_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        if world == 1:
            cpu_best = cpu_reference_best(units, args, args.cpu_sample)
        else:                                    # the host baseline is an N = 1 measurement (see --impl reference)
            cpu_best = {"value": None, "seconds": 0.0, "cores": 0, "frames": 0, "how": "not measured at N > 1", "tried": [],
                        "affinity_cores": len(os.sched_getaffinity(0))}
        cpu_fps, cpu_dt, cores = cpu_best["value"], cpu_best["seconds"], cpu_best["cores"]
        seq = None
        if world == 1 and args.sequence > 0:
            try:
                seq = sequence_mode(ctx, torch, cal, args.sequence)
            except Exception as e:           # a side measurement must never cost the headline line
                seq = {"error": str(e)[:200]}
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/i32 fixed point + f32 (LK), f64 (pose)", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": B * 4 * W_IMG * H_IMG + B * 32,
                    "d2h_bytes_per_step": B * 152, "ms_per_step": t_e2e_ms / args.steps,
                    "wall_ms_per_step": 1e3 * t_e2e_wall / args.steps,
                    "mode": "vo_batch_submit / vo_batch_wait, two submissions of units_per_gpu in flight; timed from the "
                            "first submit to the last wait, every step's H2D + kernels + D2H inside",
                    "equals_resident_results": all(a["n_inliers"] == b["n_inliers"] and np.array_equal(a["tvec"], b["tvec"])
                                                   for a, b in zip(res_e2e, res))},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_lk_ring", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "features_per_launch": feats_per_launch,
                         "avg_launch_ms": lk_avg_ms, "lk_share_of_step": lk_ms / t_single_ms if t_single_ms else None,
                         "single_stream_ms_per_step": t_single_ms / args.steps,
                         "note": "algorithmic bytes per SURVEY.md 8(d); the kernel is ALU/latency bound, see DESIGN.md"},
            "cpu_baseline": {"value": cpu_fps, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{cpu_best['frames']} frames of the same workload in {cpu_dt:.1f} s; {cpu_best['how']}; cv2 (the "
                                       f"OpenCV the reference's calls resolve to) through oracle/ref_path.py glue; affinity cores="
                                       f"{cpu_best['affinity_cores']}; tried: {cpu_best['tried']}"},
            "clocks": clocks,
            "sequence_mode": seq,
            "parity": {"n_valid": [r["n_valid"] for r in res], "n_inliers": [r["n_inliers"] for r in res]},
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
