#!/usr/bin/env python
"""bench.py -- stereo pairs/s of the B200 hot path at BASELINE.json config 2 (640x480, 1000 ORB kp,
full detect + L/R match + tracking + F-matrix RANSAC + Gauss-Newton), one process per GPU.

  python bench.py --gpus 1 --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                     the reference's CPU path (oracle: real OpenCV
                                                           calls + C++ restatement of the reference's glue)

A "step" advances B independent stereo sequences (one slot each) by one frame.  Rank 0 prints ONE JSON line.
  value  : pairs/s, inputs resident in HBM, timed with CUDA events on the library's stream
  e2e    : pairs/s through svo_process_batch with pinned HOST images (H2D + results D2H inside the timing)
  roofline : the dominant kernel's algorithmic HBM bytes / measured time vs the measured HBM peak
  cpu_baseline : the CPU oracle on the host cores, bounded sample
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "stereo_pairs_per_s"
UNIT = "pairs/s"
CONFIG = "c2"
N_SEQ = 8            # distinct synthetic sequences (seeds 0..7), replicated over the slots with phase offsets
N_FRAMES = 16        # frames per sequence, played forward and backward so consecutive steps are adjacent frames
WORKLOAD = ("c2: 640x480 synthetic stereo sequences, 1000 ORB kp (orb_nfeats=500), full FAST+ORB detect / L-R match / "
            "tracking + F-matrix RANSAC / robust Gauss-Newton path; 8 distinct 16-frame sequences played forward and backward")
HM_POPC, HM_ALU = 4, 17   # full 256-bit pair-distance (raw kernel): POPC-pipe and ALU-pipe (16 LOP3 + 1 min) instructions
HM_POPC_PFX, HM_ALU_PFX = 3, 7   # 128-bit prefix test of the pipeline kernels (6 LOP3 + 1 compare)


def pingpong(t, n):
    period = 2 * (n - 1)
    t %= period
    return t if t < n else period - t


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json (of measured)"
    return 6650.0, "fallback 6.65 TB/s of /opt/skills/guides/B200_PROFILING.md (of fallback)"


def _make_one(args):
    from stereo_vo_b200 import synth
    config, seed, n_frames, two = args
    return (synth.make_sequence_two_motions if two else synth.make_sequence)(config, seed=seed, n_frames=n_frames)


def make_frames(n_seq=N_SEQ, n_frames=N_FRAMES, config=CONFIG, two_motions=False):
    """the synthetic sequences (host side, outside every timed region), rendered by a process pool"""
    import multiprocessing as mp
    jobs = [(config, s, n_frames, two_motions) for s in range(n_seq)]
    try:
        with mp.get_context("fork").Pool(min(n_seq, os.cpu_count() or 1)) as pool:
            res = pool.map(_make_one, jobs)
    except Exception:
        res = [_make_one(j) for j in jobs]
    return [r[0] for r in res], res[0][1]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device = device
        self.stop_flag = False
        self.rows = []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, nme in enumerate(names):
                if len(r) > 3 + i and r[3 + i].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline: the oracle with the real OpenCV calls, one process per core
# ------------------------------------------------------------------------------------------------
def _cpu_estimator(backend, kind):
    from oracle import oracle as orc
    P = orc.Params(orb_nfeats=500)
    rs = "cv2_tolerant" if backend == "cv2" else "off"
    if kind == "reference":
        from oracle import ref
        return ref.RefEstimator(P, backend=backend, ransac=rs)
    return orc.OracleEstimator(P, backend=backend, ransac=rs)


def _cpu_worker(args):
    frames, cam, n_pairs, backend, kind = args
    try:
        import cv2
        cv2.setNumThreads(1)
    except Exception:
        backend = "port"
    from oracle import oracle as orc
    camo = orc.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    est = _cpu_estimator(backend, kind)
    kw = dict(light=True) if kind == "reference" else {}  # the class API's own output only (no parity read-backs)
    est.process_new_image_pair(*frames[0], camo, **kw)    # first frame (untimed: no tracking / GN yet)
    t0 = time.perf_counter()
    valid = 0
    for t in range(1, n_pairs + 1):
        o = est.process_new_image_pair(*frames[pingpong(t, len(frames))], camo, **kw)
        valid += int(o["valid"])
    return time.perf_counter() - t0, valid


def _cpu_single(args):
    frames, cam, n_pairs, threads = args
    import cv2
    cv2.setNumThreads(threads)            # 1 = serial OpenCV; -1 (cv2's reset value) = OpenCV's default thread pool
    from oracle import oracle as orc
    camo = orc.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    kind = cpu_kind()
    est = _cpu_estimator("cv2", kind)
    kw = dict(light=True) if kind == "reference" else {}
    est.process_new_image_pair(*frames[0], camo, **kw)
    t0 = time.perf_counter()
    for t in range(1, n_pairs + 1):
        est.process_new_image_pair(*frames[pingpong(t, len(frames))], camo, **kw)
    return n_pairs / (time.perf_counter() - t0), cv2.getNumThreads()


def cpu_single_process_variants(seqs, cam, n_pairs=40):
    """one process with one OpenCV thread, and one process with OpenCV's default threads (each in a
    fresh process, so that the thread setting cannot leak into the throughput workers)."""
    import multiprocessing as mp
    try:
        import cv2  # noqa: F401
    except Exception:
        return None
    out = {}
    ctxmp = mp.get_context("fork")
    for name, threads in (("1_process_1_thread", 1), ("1_process_default_threads", -1)):
        with ctxmp.Pool(1) as pool:
            rate, nthr = pool.map(_cpu_single, [(seqs[0], cam, n_pairs, threads)])[0]
        out[name] = {"pairs_per_s": round(rate, 2), "opencv_threads": nthr, "sample": f"{n_pairs} consecutive {CONFIG} pairs"}
    return out


def cpu_kind():
    """'reference': oracle/_ref/libsvo_ref.so = the reference's own stage sources compiled against stand-in headers, its
    OpenCV calls served by the real OpenCV (cv2); 'port': the restated oracle (when that library is not there)."""
    try:
        from oracle import ref
        return "reference" if os.path.exists(ref.LIB) or os.path.isdir(ref.REF_SRC) else "port"
    except Exception:
        return "port"


def cpu_throughput(seqs, cam, pairs_per_proc, procs):
    """pairs/s of the reference's CPU path with `procs` processes x 1 OpenCV thread, one sequence each."""
    import multiprocessing as mp
    try:
        import cv2  # noqa: F401
        backend = "cv2"
    except Exception:
        backend = "port"
    from oracle import oracle as orc
    orc.build()
    orc.lib()                                   # the parent loads what its workers run, so the record shows it
    kind = cpu_kind()
    if kind == "reference":
        from oracle import ref
        ref.lib()
    ctxmp = mp.get_context("fork")
    jobs = [(seqs[i % len(seqs)], cam, pairs_per_proc, backend, kind) for i in range(procs)]
    t0 = time.perf_counter()
    with ctxmp.Pool(procs) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    busy = max(r[0] for r in res)
    total = pairs_per_proc * procs
    what = ("the reference's own stage sources (oracle/_ref, compiled unmodified against stand-in headers)" if kind == "reference"
            else "the C++ restatement of the reference's glue (oracle/svo_oracle.cpp)")
    return total / busy, dict(kind=kind, cores=procs, backend=backend,
                              sample=f"{procs} processes x {pairs_per_proc} consecutive {CONFIG} pairs "
                                     f"({total} pairs, {busy:.1f} s busy, {wall:.1f} s wall); {what}; "
                                     f"third-party calls via {'cv2 ' + __import__('cv2').__version__ if backend == 'cv2' else 'the C++ port'}",
                              valid=int(sum(r[1] for r in res)))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    seqs, cam = make_frames()
    procs = os.cpu_count() or 1
    # warm-up doubles as a rate probe: the per-step sample is sized so that the whole run stays near one minute
    rate, _ = cpu_throughput(seqs, cam, 4, procs)
    budget_s = 60.0
    per = int(budget_s * rate / max(1, args.steps) / procs)
    per = max(2, min(args.ref_pairs, per))
    vals = []
    info = None
    t_all = time.perf_counter()
    for _ in range(args.steps):
        v, info = cpu_throughput(seqs, cam, per, procs)
        vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * (time.perf_counter() - t_all) / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8+f64", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "cpu_arm": "the reference's CPU implementation of the path, one process per host core, each step a bounded "
                              "sample of the workload"},
        "cpu_baseline": dict(info, value=value, unit=UNIT),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# algorithmic HBM bytes per kernel launch (no re-reads counted)
# ------------------------------------------------------------------------------------------------
def algorithmic_bytes(B, W, H, N, M, T):
    """COMPULSORY HBM bytes per launch, SURVEY.md section 8(d): detect per image = W*H + 44*N (image in once; 12 B keypoint
    + 32 B descriptor out), charged once -- the image read to fast_blur_kernel, the keypoints to select_kernel, the
    descriptors to brief_kernel; the blur / score planes those kernels exchange are implementation traffic and are NOT
    counted (they show up in `traffic`).  match(NL, NR) = 32(NL+NR) + 8 NL, track = 2 [32(Mp+Mc) + 8 Mp], GN = 32 T + 8 T + 96;
    the small glue kernels move index lists."""
    img = W * H
    return {
        "fast_blur_kernel": 2 * B * img,
        "select_kernel": 2 * B * 12 * N,
        "brief_kernel": 2 * B * 32 * N,
        "hamming_kernel(lr)": B * (32 * (N + N) + 8 * N),
        "lr_filter_kernel": B * (8 * N + 16 * M),
        "hamming_kernel(track)": 2 * B * (32 * (M + M) + 8 * M),
        "track_filter_kernel": B * (16 * M + 16 * M),
        "ransac_prepare_kernel": B * (16 * M + 2 * 16 * M),
        "fundmat_kernel": 2 * B * (16 * M + M),
        "consistency_kernel": B * (16 * M + 2 * M + 8 * T),
        "gn_kernel": B * (32 * T + 8 * T + 96),
        "begin_step": B * 64,
    }


def reduce_over_ranks(ms, e2e_ms, pairs, n_valid, launches, device):
    """max over ranks of the elapsed times + the end-of-batch gather of the per-rank counters: the only
    collectives on this path (NCCL on the GPU box; gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return ms, e2e_ms, pairs, n_valid, launches
    tm = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=device)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    cnt = torch.tensor([pairs, n_valid, launches], dtype=torch.int64, device=device)
    gathered = [torch.zeros_like(cnt) for _ in range(dist.get_world_size())]
    dist.all_gather(gathered, cnt)
    return (float(tm[0]), float(tm[1]), int(sum(int(g[0]) for g in gathered)), int(sum(int(g[1]) for g in gathered)),
            int(sum(int(g[2]) for g in gathered)))



# ------------------------------------------------------------------------------------------------
# BASELINE config 5: descriptor-match microbench (4096 L x 4096 R, 256-bit Hamming brute force) and the
# NR sweep across the roofline ridge: small NR = HBM-bound streaming of the queries,
# large NR = integer-pipe bound.
# ------------------------------------------------------------------------------------------------
C5_N = 4096


def hamming_algo_bytes(nq, nt, reps):
    return reps * (32 * (nq + nt) + 4 * nq)        # every descriptor once in, one packed (dist<<16|idx) word out


def time_hamming(ctx, torch, stream, d_q, nq, d_t, nt, reps, d_out, iters, warm=3, flush=None):
    """mean milliseconds per svo_match_hamming_device launch (CUDA events on the library stream)."""
    for _ in range(warm):
        ctx.match_hamming_device(d_q.data_ptr(), nq, d_t.data_ptr(), nt, reps, d_out.data_ptr())
    ctx.synchronize()
    tot = 0.0
    for _ in range(iters):
        if flush is not None:
            with torch.cuda.stream(stream):
                flush.add_(1)                      # > L2: the next launch reads its inputs from HBM
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.match_hamming_device(d_q.data_ptr(), nq, d_t.data_ptr(), nt, reps, d_out.data_ptr())
        e1.record(stream)
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def c5_microbench(ctx, torch, stream, device, pipe, peak, reps=64, iters=10, sweep=True, seed=5):
    g = torch.Generator(device="cpu").manual_seed(seed)
    q = torch.randint(0, 256, (reps, C5_N, 32), dtype=torch.uint8, generator=g).to(device)
    t = torch.randint(0, 256, (reps, C5_N, 32), dtype=torch.uint8, generator=g).to(device)
    out = torch.zeros((reps, C5_N), dtype=torch.int32, device=device)
    flush = torch.zeros(64 << 20, dtype=torch.int32, device=device)       # 256 MB > 126 MB L2
    ms = time_hamming(ctx, torch, stream, q, C5_N, t, C5_N, reps, out, iters, flush=flush)
    pd = reps * C5_N * C5_N / (ms * 1e-3)
    int_roof = 1.0 / max(HM_POPC / pipe["popc_per_s"], HM_ALU / pipe["lop3_per_s"])
    res = {"problem": f"{reps} independent {C5_N}x{C5_N} problems per launch", "ms": ms, "pair_distances_per_s": pd,
           "frac_of_int_pipe_roof": pd / int_roof, "algo_bytes": hamming_algo_bytes(C5_N, C5_N, reps),
           "hbm_gbs": hamming_algo_bytes(C5_N, C5_N, reps) / (ms * 1e-3) / 1e9}
    res["hbm_frac"] = res["hbm_gbs"] / peak
    if sweep:
        # one problem, NL = 8 Mi queries (256 MB of descriptors, larger than L2) against NR train rows: crosses the
        # ridge (NR ~ 8); below it the kernel streams the queries once and "fraction of HBM peak" is the figure
        nq = 8 << 20
        qb = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, device=device)
        ob = torch.zeros((nq,), dtype=torch.int32, device=device)
        rows = []
        for nt in (1, 2, 4, 8, 16, 64, 256, 1024):
            m = time_hamming(ctx, torch, stream, qb, nq, t, nt, 1, ob, 3, warm=1)
            ab = hamming_algo_bytes(nq, nt, 1)
            rows.append({"nr": nt, "nl": nq, "us": round(m * 1e3, 2), "hbm_gbs": round(ab / (m * 1e-3) / 1e9, 1),
                         "hbm_frac": round(ab / (m * 1e-3) / 1e9 / peak, 4),
                         "int_frac": round(nq * nt / (m * 1e-3) / int_roof, 4)})
        res["sweep"] = rows
        del qb, ob
    return res, (q, t, out)



def stage1_microbench(ctx, torch, stream, device, W, H, peak, n=128, iters=5):
    """stage 1 on the device (gray conversion + rectification, csrc/prepare.cu): HBM-bound streaming kernels, timed
    alone on n stereo pairs (inputs larger than L2).  Algorithmic bytes per output pixel: channels (source) + 1."""
    pitch = (W + 15) & ~15
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float64)
    r2 = ((xs - W / 2) / W) ** 2 + ((ys - H / 2) / W) ** 2
    maps = []
    for side in (0, 1):          # mild radial distortion, CV_16SC2 + CV_16UC1 layout of cv::initUndistortRectifyMap
        fx = np.rint((xs + (xs - W / 2) * (-0.06 + 0.02 * side) * r2) * 32).astype(np.int64)
        fy = np.rint((ys + (ys - H / 2) * (-0.06 + 0.02 * side) * r2) * 32).astype(np.int64)
        maps.append((np.stack([fx >> 5, fy >> 5], -1).astype(np.int16), ((fy & 31) * 32 + (fx & 31)).astype(np.uint16)))
    out = {}
    dst = [torch.zeros((n, H, pitch), dtype=torch.uint8, device=device) for _ in range(2)]
    for ch, rect in ((3, True), (1, True), (3, False)):
        if rect:
            ctx.set_rectify_maps(maps[0][0], maps[0][1], maps[1][0], maps[1][1])
        else:
            ctx.set_rectify_maps()
        src = [torch.randint(0, 256, (n, H, W * ch), dtype=torch.uint8, device=device) for _ in range(2)]
        tot = 0.0
        for it in range(iters + 2):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.prepare_rectify_device(src[0].data_ptr(), src[1].data_ptr(), n, W * ch, H * W * ch, ch, dst[0].data_ptr(), dst[1].data_ptr())
            e1.record(stream)
            e1.synchronize()
            if it >= 2:
                tot += e0.elapsed_time(e1)
        ms = tot / iters
        ab = 2 * n * W * H * (ch + 1)
        out["%s%s" % ("bgr" if ch == 3 else "gray", "+rectify" if rect else "->gray")] = {
            "ms": round(ms, 4), "pairs": n, "algo_bytes": ab, "hbm_gbs": round(ab / (ms * 1e-3) / 1e9, 1),
            "hbm_frac": round(ab / (ms * 1e-3) / 1e9 / peak, 4)}
        del src
    ctx.set_rectify_maps()
    return out

def single_pair_latency(capi, seqs, gc, W, H, nfeats, device_index, n=60):
    """what one caller with ONE sequence sees: svo_process_pair per frame (host images in, result out, synchronous)"""
    ctx = capi.Context(width=W, height=H, orb_nfeats=nfeats, max_seqs=1, device=device_index)
    frames = seqs[0]
    ts = []
    valid = 0
    for t in range(n + 5):
        L, R = frames[pingpong(t, len(frames))]
        t0 = time.perf_counter()
        r = ctx.process_pair(L, R, gc)
        dt = time.perf_counter() - t0
        if t >= 5:
            ts.append(dt)
            valid += int(r.valid)
    ctx.close()
    ts = np.array(ts) * 1e3
    return {"median_ms": round(float(np.median(ts)), 3), "p90_ms": round(float(np.percentile(ts, 90)), 3),
            "pairs_per_s": round(1e3 / float(np.median(ts)), 1), "valid": f"{valid}/{n}",
            "note": "batch of one: the kernels are latency-bound here, throughput needs many sequences per step"}


def octaves_throughput(capi, torch, dev, gc, W, H, nfeats, device_index, peak, n_octaves=3, batch=256, steps=20, warmup=3):
    """params_rectify.nOctaves = 3, the reference's ctor default (stage1_rectify.cpp:28): the pyramid is built on the
    device and stages 2-4 run per octave before one flattened stage 5.  Device-resident images, CUDA events on the
    library's stream; a second line beside the headline (which runs the one-octave INI of SURVEY Appendix C)."""
    batch = min(batch, dev[0][0].shape[0])
    ctx = capi.Context(width=W, height=H, orb_nfeats=nfeats, max_seqs=batch, device=device_index, n_octaves=n_octaves)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", device_index))
    cams = [gc] * batch

    def step(t):
        L, R = dev[t % len(dev)]
        ctx.process_batch_device(L.data_ptr(), R.data_ptr(), batch, W, W * H, cams)
    for t in range(warmup):
        step(t)
    ctx.synchronize()
    l0 = ctx.kernel_launches()
    with torch.cuda.stream(stream):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(steps):
            step(warmup + t)
        e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / steps
    res = ctx.fetch_results(batch)
    stats = np.mean([ctx.octave_stats(s) for s in range(0, batch, max(1, batch // 16))], axis=0)
    # the pyramid alone: the one streaming, HBM-bound kernel of this path.  Algorithmic bytes per level = source image
    # read once + a quarter of it written; inputs cycle through the distinct batches (larger than L2 together)
    pyr_ms = []
    with torch.cuda.stream(stream):
        for t in range(3 + 10):
            L, R = dev[t % len(dev)]
            p0 = torch.cuda.Event(enable_timing=True)
            p1 = torch.cuda.Event(enable_timing=True)
            p0.record()
            ctx.build_pyramid_device(L.data_ptr(), R.data_ptr(), batch, W, W * H)
            p1.record()
            p1.synchronize()
            if t >= 3:
                pyr_ms.append(p0.elapsed_time(p1))
    pyr_bytes = sum(2 * batch * (w * h + (w // 2) * (h // 2)) for (w, h) in [tuple(o[:2]) for o in ctx.octave_info[:-1]])
    pyr_t = float(np.median(pyr_ms))
    out = {"n_octaves": n_octaves, "pairs_per_step": batch, "ms_per_step": round(ms, 4),
           "pairs_per_s": round(batch / ms * 1e3, 1), "valid_results": f"{sum(int(r.valid) for r in res)}/{batch}",
           "launches_per_step": (ctx.kernel_launches() - l0) // steps,
           "per_octave_mean": {"keypoints_left": [round(float(v), 1) for v in stats[:, 0]],
                               "stereo_matches": [round(float(v), 1) for v in stats[:, 2]],
                               "tracked_pairs": [round(float(v), 1) for v in stats[:, 3]]},
           "octave_sizes": [list(o[:2]) for o in ctx.octave_info], "quotas": [o[2] for o in ctx.octave_info],
           "pyramid": {"kernel": "pyramid_kernel (all levels of one step)", "ms": round(pyr_t, 4), "algo_bytes": pyr_bytes,
                       "hbm_gbs": round(pyr_bytes / (pyr_t * 1e-3) / 1e9, 1),
                       "hbm_frac": round(pyr_bytes / (pyr_t * 1e-3) / 1e9 / peak, 4) if peak else None,
                       "note": "two launches (levels 1 and 2) timed together with CUDA events on the library's stream"}}
    ctx.close()
    return out


def run_c5(args):
    """--workload c5: the microbench on N GPUs.  Query rows are sharded across the ranks, the train set is
    replicated (broadcast from rank 0), the packed results are all-gathered (the one exchange step, NCCL)."""
    import torch
    import torch.distributed as dist
    from stereo_vo_b200 import capi
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    ctx = capi.Context(width=320, height=240, orb_nfeats=100, max_seqs=1, device=local)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=device)
    pipe = ctx.pipe_peaks()
    peak, peak_src = load_peaks()
    reps = args.c5_reps
    micro, (q, t, out) = c5_microbench(ctx, torch, stream, device, pipe, peak, reps=reps, iters=5, sweep=(rank == 0), seed=5 + rank)
    if world > 1:
        dist.broadcast(t, src=0)                       # replicated train set
    gathered = torch.zeros((world,) + tuple(out.shape), dtype=out.dtype, device=device)
    flush = torch.zeros(64 << 20, dtype=torch.int32, device=device)

    def step():
        with torch.cuda.stream(stream):
            flush.add_(1)
        ctx.match_hamming_device(q.data_ptr(), C5_N, t.data_ptr(), C5_N, reps, out.data_ptr())
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(gathered, out)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    sampler.stop_flag = True
    ms = e0.elapsed_time(e1)
    if world > 1:
        tm = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm[0])
        assert torch.equal(gathered[rank], out)
    if rank == 0:
        pd = world * reps * C5_N * C5_N * args.steps / (ms * 1e-3)
        ab = hamming_algo_bytes(C5_N, C5_N, reps)
        kms = micro["ms"]
        line = {"metric": "hamming_pair_distances_per_s", "value": pd, "unit": "pair-distances/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": f"c5: {reps} x ({C5_N} L x {C5_N} R) 256-bit Hamming brute force per GPU per step; "
                                       "query rows sharded over the ranks, train set replicated, results all-gathered",
                           "l2": "256 MB flush buffer written between launches"},
                "gpu_launches": args.steps * world,
                "roofline": {"bound": "hbm", "kernel": "hamming_raw_kernel", "achieved": ab / (kms * 1e-3) / 1e9, "peak": peak,
                             "unit": "GB/s", "frac": ab / (kms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                             "note": "N x N brute force is integer-pipe bound (ridge at NR ~ 8); see sweep for the HBM-bound regime"},
                "match_microbench": micro, "clocks": sampler.summary()}
        emit(line)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()

def bind_to_gpu_numa(torch, local):
    """Pin this rank's threads (and, by first touch, the pinned host buffers it allocates afterwards) to the CPUs next to
    its GPU: NVML's CPU affinity of the device = the NUMA node of its PCIe root."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
        node = None
        try:
            node = int(open(f"/sys/bus/pci/devices/{bus[4:].lower()}/numa_node").read())
        except Exception:
            pass
        return {"pci": bus, "cpus": f"{min(cpus)}-{max(cpus)} ({len(cpus)})" if cpus else None, "numa_node": node}
    except Exception as e:          # noqa: BLE001
        return {"error": str(e)[:120]}


def gather_floats(vals, device):
    """per-rank values -> list of per-rank lists (all_gather over NCCL; one rank: itself)"""
    import torch
    import torch.distributed as dist
    t = torch.tensor(vals, dtype=torch.float64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [t.tolist()]
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.tolist() for o in out]


class Playback:
    """Slot i plays sequence (i + rank) % n_seq forward and backward, starting i frames in.  Device side: one (B, H, W)
    tensor pair per phase (period of them, together far larger than L2).  Host side: every slot's image is its own
    buffer inside a pinned pool of frames -- the C-ABI gets one pointer per image, as a caller with one buffer per camera
    image has them."""

    def __init__(self, torch, seqs, B, rank, device, pinned=True):
        self.n_seq, self.n_frames = len(seqs), len(seqs[0])
        self.period = 2 * (self.n_frames - 1)
        H, W = seqs[0][0][0].shape
        self.B, self.H, self.W = B, H, W
        pool = torch.empty((self.n_seq, self.n_frames, 2, H, W), dtype=torch.uint8)
        if pinned:
            pool = pool.pin_memory()
        for s_, frames in enumerate(seqs):
            for f, (L, R) in enumerate(frames):
                pool[s_, f, 0].copy_(torch.from_numpy(L))
                pool[s_, f, 1].copy_(torch.from_numpy(R))
        self.pool = pool
        self.seq_of = [(i + rank) % self.n_seq for i in range(B)]
        self.device = device
        self.dev = None

    def frame_of(self, t, i):
        return pingpong(t + i, self.n_frames)

    def build_device_batches(self, torch):
        pd = self.pool.to(self.device, non_blocking=True)
        sidx = torch.tensor(self.seq_of, device=self.device)
        self.dev = []
        for t in range(self.period):
            fidx = torch.tensor([self.frame_of(t, i) for i in range(self.B)], device=self.device)
            self.dev.append((pd[sidx, fidx, 0].contiguous(), pd[sidx, fidx, 1].contiguous()))
        torch.cuda.synchronize()
        del pd

    def pointer_arrays(self):
        """per phase: two ctypes arrays of B host pointers (left images, right images)"""
        base = self.pool.data_ptr()
        plane = self.H * self.W
        out = []
        for t in range(self.period):
            lp = (C.c_void_p * self.B)()
            rp = (C.c_void_p * self.B)()
            for i in range(self.B):
                off = ((self.seq_of[i] * self.n_frames + self.frame_of(t, i)) * 2) * plane
                lp[i] = base + off
                rp[i] = base + off + plane
            out.append((lp, rp))
        return out


def e2e_loop(lib, capi, ctx, ptrs, B, W, cam_arr, steps, warm, detail, cap):
    """K steps through svo_submit_batch / svo_wait_batch[_detail] with host images: every step's images go host -> device
    and every step's results come back device -> host inside the timed region; three batches in flight."""
    DEPTH = 3
    outs = [(capi.Result * B)() for _ in range(DEPTH)]
    resid = np.zeros((B, cap)) if detail else None
    outl = np.zeros((B, cap), np.int64) if detail else None
    stats = np.zeros((B, 16), np.int32) if detail else None
    period = len(ptrs)

    def submit(t):
        lp, rp = ptrs[t % period]
        rc = lib.svo_submit_batch(ctx._h, B, None, lp, rp, C.c_size_t(W), cam_arr, None)
        if rc < 0:
            raise RuntimeError(lib.svo_last_error(ctx._h).decode())

    def wait(k):
        if detail:
            rc = lib.svo_wait_batch_detail(ctx._h, outs[k % DEPTH], resid.ctypes.data_as(C.c_void_p), outl.ctypes.data_as(C.c_void_p),
                                           cap, stats.ctypes.data_as(C.c_void_p))
        else:
            rc = lib.svo_wait_batch(ctx._h, outs[k % DEPTH])
        if rc < 0:
            raise RuntimeError(lib.svo_last_error(ctx._h).decode())

    if lib.svo_set_result_detail(ctx._h, 1 if detail else 0) != 0:
        raise RuntimeError(lib.svo_last_error(ctx._h).decode())
    ctx.reset_all()
    t = 0
    for _ in range(max(3, warm)):                      # warm-up through the same public calls
        submit(t)
        wait(t)
        t += 1
    ctx.synchronize()
    t0 = time.perf_counter()
    for k in range(steps):
        submit(t + k)
        if k >= DEPTH - 1:
            wait(k - (DEPTH - 1))
    for k in range(max(0, steps - (DEPTH - 1)), steps):
        wait(k)
    ms = 1e3 * (time.perf_counter() - t0)
    out = outs[(steps - 1) % DEPTH]
    valid = sum(int(r.valid) for r in out)
    d2h = B * C.sizeof(capi.Result) + (B * C.sizeof(capi.Result) // 3)            # results + slot states
    if detail:
        d2h += B * cap * 12 + B * 64
        nres = int(sum(r.n_residuals for r in out))
        assert nres > 0 and float(np.abs(resid[0, :out[0].n_residuals]).sum()) >= 0.0
    return ms, valid, d2h


def side_workload(capi, torch, local, config, nfeats, batch, n_seq, n_frames, steps, warmup, two_motions=False, **ctxkw):
    """a secondary device-resident line (BASELINE config 3, the outlier-heavy variant): pairs/s with CUDA events on the
    library's stream, plus the data-dependent work it did (RANSAC iterations, Gauss-Newton iterations, inlier ratios)"""
    device = torch.device("cuda", local)
    seqs, cam = make_frames(n_seq, n_frames, config, two_motions)
    gc = capi.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    pb = Playback(torch, seqs, batch, 0, device, pinned=False)
    pb.build_device_batches(torch)
    H, W = pb.H, pb.W
    ctx = capi.Context(width=W, height=H, orb_nfeats=nfeats, max_seqs=batch, device=local,
                       max_corners=32768 if config == "c3" else 16384, **ctxkw)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=device)
    cams = [gc] * batch

    def step(t):
        L, R = pb.dev[t % pb.period]
        ctx.process_batch_device(L.data_ptr(), R.data_ptr(), batch, W, W * H, cams)
    for t in range(warmup):
        step(t)
    ctx.synchronize()
    l0 = ctx.kernel_launches()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for t in range(steps):
        step(warmup + t)
    e1.record(stream)
    e1.synchronize()
    ms = e0.elapsed_time(e1) / steps
    res = ctx.fetch_results(batch)
    rs = np.array([ctx.ransac_stats(s_) for s_ in range(0, batch, max(1, batch // 32))], np.float64)    # (k, 2, 3)
    stats = np.mean([ctx.octave_stats(s_)[0] for s_ in range(0, batch, max(1, batch // 32))], axis=0)
    out = {"workload": f"{config}: {W}x{H}, {2 * nfeats} ORB kp, {n_seq} distinct {n_frames}-frame sequences"
                       + (", 38 % of every image under a second rigid motion (false tracks)" if two_motions else ""),
           "pairs_per_step": batch, "ms_per_step": round(ms, 4), "pairs_per_s": round(batch / ms * 1e3, 1),
           "valid_results": f"{sum(int(r.valid) for r in res)}/{batch}",
           "launches_per_step": (ctx.kernel_launches() - l0) // steps,
           "image_bytes_per_step": 2 * batch * W * H,
           "mean": {"keypoints_left": round(float(stats[0]), 1), "stereo_matches": round(float(stats[2]), 1),
                    "tracked_pairs": round(float(stats[3]), 1),
                    "ransac_points": round(float(rs[:, :, 0].mean()), 1),
                    "ransac_inlier_ratio": round(float((rs[:, :, 1] / np.maximum(rs[:, :, 0], 1)).mean()), 4),
                    "ransac_iterations": round(float(rs[:, :, 2].mean()), 2),
                    "gn_iterations": round(float(np.mean([r.num_it + r.num_it_final for r in res])), 2),
                    "gn_inliers": round(float(np.mean([r.n_outliers for r in res])), 1)}}
    ctx.close()
    del pb
    return out


def c5_sharded(capi, torch, dist, local, world, rank, pipe, reps, steps=10, warmup=3):
    """BASELINE config 5 on N GPUs inside the default run: the query rows of `world * reps` 4096 x 4096 problems are sharded
    over the ranks (reps problems each), the train sets are replicated (broadcast from rank 0) and the packed (distance,
    index) results are all-gathered every step -- the one exchange step of this path (NCCL over NVLink)."""
    device = torch.device("cuda", local)
    ctx = capi.Context(width=320, height=240, orb_nfeats=100, max_seqs=1, device=local)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=device)
    g = torch.Generator(device="cpu").manual_seed(5 + rank)
    q = torch.randint(0, 256, (reps, C5_N, 32), dtype=torch.uint8, generator=g).to(device)
    t = torch.randint(0, 256, (reps, C5_N, 32), dtype=torch.uint8, generator=g).to(device)
    out = torch.zeros((reps, C5_N), dtype=torch.int32, device=device)
    if world > 1:
        dist.broadcast(t, src=0)
    gathered = torch.zeros((world,) + tuple(out.shape), dtype=out.dtype, device=device)
    flush = torch.zeros(64 << 20, dtype=torch.int32, device=device)

    def step():
        with torch.cuda.stream(stream):
            flush.add_(1)
        ctx.match_hamming_device(q.data_ptr(), C5_N, t.data_ptr(), C5_N, reps, out.data_ptr())
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(gathered, out)
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    e1.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tm = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm[0])
        assert torch.equal(gathered[rank], out)
    ctx.close()
    pd = world * reps * C5_N * C5_N * steps / (ms * 1e-3)
    int_roof = 1.0 / max(HM_POPC / pipe["popc_per_s"], HM_ALU / pipe["lop3_per_s"])
    return {"n_gpus": world, "problems_per_gpu_per_step": reps, "ms_per_step": round(ms / steps, 4),
            "pair_distances_per_s": pd, "frac_of_int_pipe_roof_per_gpu": pd / world / int_roof,
            "all_gather_bytes_per_step": world * reps * C5_N * 4 if world > 1 else 0,
            "note": "whole-job figure over all ranks, max-over-ranks device time; includes the 256 MB L2 flush write of every step"}


def class_api_bench(seqs, cam, B, steps, warmup, nfeats, local, rank):
    """e2e through the reference's CLASS API: B rso::CStereoOdometryEstimator instances of one CStereoOdometryPool
    (stereo_vo_b200/host/pool_driver.cpp), one processNewImagePair per estimator per step, complete results back."""
    import struct
    import tempfile
    drv = os.path.join(ROOT, "stereo_vo_b200", "host", "svo_pool_driver")
    if not os.path.exists(drv):
        return None
    H, W = seqs[0][0][0].shape
    with tempfile.NamedTemporaryFile(suffix=".bin", prefix=f"svo_seqs_r{rank}_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as f:
        f.write(struct.pack("<4i5d", W, H, len(seqs), len(seqs[0]), cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"]))
        for frames in seqs:
            for (L, R) in frames:
                f.write(L.tobytes())
                f.write(R.tobytes())
        f.flush()
        p = subprocess.run([drv, "bench", f.name, str(B), str(steps), str(warmup), str(nfeats), str(local)], capture_output=True, text=True)
    if p.returncode != 0:
        return {"error": (p.stderr or p.stdout)[-300:]}
    return json.loads(p.stdout.strip().splitlines()[-1])


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from stereo_vo_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    # pageable host images are gathered into pinned staging by host threads of the library: share the cores between the ranks
    os.environ.setdefault("SVO_GATHER_THREADS", str(max(4, min(16, (os.cpu_count() or 16) // world))))
    numa = bind_to_gpu_numa(torch, local) if not args.no_numa else {"disabled": True}
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    B = args.batch
    W, H, nfeats = dict(c1=(320, 240, 100), c2=(640, 480, 500), c3=(1242, 375, 1000))[CONFIG]
    seqs, cam = make_frames()
    gc = capi.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    pb = Playback(torch, seqs, B, rank, device, pinned=True)
    pb.build_device_batches(torch)
    period = pb.period

    ctx = capi.Context(width=W, height=H, orb_nfeats=nfeats, max_seqs=B, device=local, match_method=args.match_method,
                       if_match_method=args.if_match_method)
    cams = [gc] * B
    stream = torch.cuda.ExternalStream(ctx.stream(), device=device)

    def step_device(t):
        L, R = pb.dev[t % period]
        ctx.process_batch_device(L.data_ptr(), R.data_ptr(), B, W, W * H, cams)

    def barrier():
        torch.cuda.synchronize()
        ctx.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ---------------------------------------------------------------------
    if args.chunks:
        ctx.set_concurrency(args.chunks)
    for t in range(args.warmup):
        step_device(t)
    ctx.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.kernel_launches()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(args.steps):
        step_device(args.warmup + k)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.kernel_launches() - launches0
    sampler.stop_flag = True
    # ---- per-kernel times: the same steps once more with one CUDA event after every kernel.  In this pass the
    #      library keeps the whole batch on one stream, so each kernel is timed alone (in the headline pass above,
    #      groups of slots run on concurrent streams and their kernels overlap) ------------------------------
    n_kt = min(args.steps, 8)
    ctx.enable_stage_timing(True)
    t_ser0 = time.perf_counter()
    for k in range(n_kt):
        step_device(args.warmup + args.steps + k)
    ctx.synchronize()
    serial_ms = 1e3 * (time.perf_counter() - t_ser0) / n_kt
    kernel_ms = ctx.stage_times()
    pipe = ctx.pipe_peaks()
    ctx.enable_stage_timing(False)
    res = ctx.fetch_results(B)
    n_valid = sum(int(r.valid) for r in res)
    Nkp = float(np.mean([r.detected_left for r in res]))
    M = float(np.mean([r.stereo_matches for r in res]))
    T = float(np.mean([r.tracked_feats_from_last_frame for r in res]))
    rs = np.array([ctx.ransac_stats(s_) for s_ in range(0, B, max(1, B // 64))], np.float64)
    work = {"ransac_iterations_mean": round(float(rs[:, :, 2].mean()), 2),
            "ransac_inlier_ratio_mean": round(float((rs[:, :, 1] / np.maximum(rs[:, :, 0], 1)).mean()), 4),
            "gn_iterations_mean": round(float(np.mean([r.num_it + r.num_it_final for r in res])), 2),
            "gn_inliers_mean": round(float(np.mean([r.n_outliers for r in res])), 1)}

    # ---- end to end through the C-ABI with host buffers ---------------------------------------------------
    # headline e2e: every image its own pinned host buffer (one pointer per image), complete TStereoOdometryResult back
    # (PODs + out_residual + outliers + per-octave counters) -- what the reference's boundary hands its caller
    lib = capi.lib()
    cam_arr = (capi.Camera * B)(*cams)
    ptrs = pb.pointer_arrays()
    cap = 2 * nfeats
    h2d_bytes = 2 * B * W * H
    e2e_ms, e2e_valid, d2h_bytes = e2e_loop(lib, capi, ctx, ptrs, B, W, cam_arr, args.steps, args.warmup, True, cap)
    variants = {}
    if not args.no_micro:
        n_var = max(8, min(args.steps, 60))
        vm, vv, vd = e2e_loop(lib, capi, ctx, ptrs, B, W, cam_arr, n_var, 3, False, cap)
        variants["pod_results_only"] = {"pairs_per_s": round(B * n_var / (vm * 1e-3), 1), "ms_per_step": round(vm / n_var, 4),
                                        "d2h_bytes_per_step": vd, "note": "svo_wait_batch: the 296-byte result structs only"}
        pbp = Playback(torch, seqs, B, rank, device, pinned=False)                     # pageable host images (malloc'd, as an MRPT CImage)
        vm, vv, vd = e2e_loop(lib, capi, ctx, pbp.pointer_arrays(), B, W, cam_arr, n_var, 3, True, cap)
        variants["pageable_host_images"] = {"pairs_per_s": round(B * n_var / (vm * 1e-3), 1), "ms_per_step": round(vm / n_var, 4),
                                            "note": "the same with the images in ordinary (not page-locked) host memory"}
        del pbp
        # one contiguous pinned block per step (the best-case layout round 1 reported): two large copies per step; two
        # blocks (phases 0 and 1 of the playback) alternate, so consecutive steps are adjacent frames
        cptrs = []
        blks = []
        for ph in (0, 1):
            blk = [torch.empty((B, H, W), dtype=torch.uint8).pin_memory() for _ in range(2)]
            blk[0].copy_(pb.dev[ph][0]); blk[1].copy_(pb.dev[ph][1])
            blks.append(blk)
            cptrs.append(((C.c_void_p * B)(*[blk[0][i].data_ptr() for i in range(B)]), (C.c_void_p * B)(*[blk[1][i].data_ptr() for i in range(B)])))
        torch.cuda.synchronize()
        vm, vv, vd = e2e_loop(lib, capi, ctx, cptrs, B, W, cam_arr, n_var, 3, True, cap)
        variants["one_contiguous_pinned_block"] = {"pairs_per_s": round(B * n_var / (vm * 1e-3), 1), "ms_per_step": round(vm / n_var, 4),
                                                   "note": "all images of a step in one pinned block: two large copies per step instead of the batched copy call"}
        del blks
    # the host->device link with every rank copying at the same time: a barrier, then one step's worth of pinned image
    # bytes host -> device, five times; e2e cannot be faster than this
    stage = [torch.empty((B, H, W), dtype=torch.uint8, device=device) for _ in range(2)]
    hblk = [torch.empty((B, H, W), dtype=torch.uint8).pin_memory() for _ in range(2)]
    barrier()
    c0 = torch.cuda.Event(enable_timing=True)
    c1 = torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        stage[0].copy_(hblk[0], non_blocking=True)
        stage[1].copy_(hblk[1], non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    c0.record()
    for _ in range(8):
        stage[0].copy_(hblk[0], non_blocking=True)
        stage[1].copy_(hblk[1], non_blocking=True)
    c1.record()
    c1.synchronize()
    h2d_ms = c0.elapsed_time(c1) / 8
    del stage, hblk
    per_rank = gather_floats([e2e_ms / args.steps, h2d_ms, ms / args.steps], device)

    # ---- max over ranks + the end-of-batch gather (the only collective on this path) ------------------------
    pairs = B * args.steps
    ms, e2e_ms, pairs_all, n_valid_all, launches_all = reduce_over_ranks(ms, e2e_ms, pairs, n_valid, launches, "cuda")

    sharded = None
    if not args.no_micro or world > 1:
        sharded = c5_sharded(capi, torch, dist, local, world, rank, pipe, reps=64)          # weak scaling: 64 problems per GPU

    if rank == 0:
        value = pairs_all / (ms * 1e-3)
        e2e_value = pairs_all / (e2e_ms * 1e-3)
        peak, peak_src = load_peaks()
        ab = algorithmic_bytes(B, W, H, Nkp, M, T)
        total_kernel_ms = sum(kernel_ms.values()) or 1.0
        kernels = {}
        for name, kms in kernel_ms.items():
            gbs = ab.get(name, 0) / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
            kernels[name] = {"ms": round(kms, 4), "share": round(kms / total_kernel_ms, 4),
                             "algo_bytes": int(ab.get(name, 0)), "gbs": round(gbs, 2)}
        dom = max(kernel_ms, key=kernel_ms.get)
        dom_gbs = ab.get(dom, 0) / (kernel_ms[dom] * 1e-3) / 1e9
        prof = {}
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                prof = json.load(open(tp))
            except Exception:
                prof = {}
        traffic = prof.get("dram_bytes_per_launch", {}).get(dom)
        inst = prof.get("warp_instructions_per_launch", {}).get(dom)
        clocks = sampler.summary()
        # the binding resource of the dominant kernel is instruction issue (ncu: DRAM throughput ~6 %, SM throughput ~83 %):
        # warp-instructions per launch (ncu, static) / live kernel time against 4 issue slots per SM per clock
        issue_peak = 148 * 4 * (clocks.get("sm_mhz") or 1965.0) * 1e6
        binding = {"what": "SM instruction issue (4 warp-instructions / clk / SM)", "unit": "warp-instructions/s",
                   "peak": issue_peak, "achieved": None, "frac": None,
                   "source": prof.get("source", "profiles/traffic.json"),
                   "note": "instruction count per launch from the committed ncu --set full capture (static), time and clock measured in this run"}
        if inst:
            binding["achieved"] = inst * (B / float(prof.get("pairs_per_launch", B))) / (kernel_ms[dom] * 1e-3)
            binding["frac"] = binding["achieved"] / issue_peak
        # Hamming stage against its binding (integer-pipe) roof.  The pipeline kernels test a 128-bit prefix of every
        # pair (3 POPC + 7 ALU-pipe ops) and finish the distance only where the prefix beats the best so far (a few
        # percent of the warp-instructions); the raw kernel of the microbench computes every full distance (4 POPC +
        # 17 ALU).  Denominators: POPC and LOP3 lane rates measured on this GPU (svo_measure_pipe_peaks).
        hm_ms = kernel_ms.get("hamming_kernel(lr)", 0) + kernel_ms.get("hamming_kernel(track)", 0)
        pairdist = B * (Nkp * Nkp + 2 * M * M)
        int_roof = 1.0 / max(HM_POPC_PFX / pipe["popc_per_s"], HM_ALU_PFX / pipe["lop3_per_s"])
        micro = stage1 = latency = octaves = c3 = outl = dmorb = klass = None
        if not args.no_micro:
            octaves = octaves_throughput(capi, torch, pb.dev, gc, W, H, nfeats, local, peak)
            micro, _ = c5_microbench(ctx, torch, stream, device, pipe, peak, reps=64, iters=5)
            stage1 = stage1_microbench(ctx, torch, stream, device, W, H, peak)
            latency = single_pair_latency(capi, seqs, gc, W, H, nfeats, local)
        ctx.close()
        del pb
        torch.cuda.empty_cache()
        if not args.no_micro:
            klass = class_api_bench(seqs, cam, B, max(8, min(args.steps, 60)), 3, nfeats, local, rank)
            c3 = side_workload(capi, torch, local, "c3", 1000, 128, 4, 8, 20, 3)
            outl = side_workload(capi, torch, local, "c2", 500, 256, 8, 8, 20, 3, two_motions=True)
            dmorb = side_workload(capi, torch, local, "c2", 500, 128, 8, 8, 10, 3, detect_method=0, max_y_diff=2.0)
            dmorb["workload"] += "; detect_method = dmORB (cv::ORB::detectAndCompute with 750 features on 8 levels, oriented, then the reference's NMS): SURVEY 8(f) rank 4"
        if micro is not None and sharded is not None:
            micro["sharded"] = sharded
        elif sharded is not None:
            micro = {"sharded": sharded}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8+f64", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "pairs_per_step_per_gpu": B, "distinct_sequences": N_SEQ, "frames_per_sequence": N_FRAMES,
                       "matchers": {"match_method": ["smDescBF", "smDescRbR"][args.match_method],
                                    "if_match_method": ["ifmDescBF", "ifmDescWin"][args.if_match_method]},
                       "streams": "slots of a step run in %s groups on concurrent streams (value, e2e); the per-kernel times "
                                  "and the roofline come from a serialised pass of the same steps (%.3f ms per step)"
                                  % (args.chunks or "the library's default (2)", serial_ms),
                       "l2": f"inputs larger than L2: {2 * B * W * H / 1e6:.0f} MB of images per step per GPU, "
                             f"{period} distinct batches cycled ({period * 2 * B * W * H / 1e9:.1f} GB resident)",
                       "keypoints_per_image": Nkp, "stereo_matches": M, "tracked_pairs": T,
                       "data_dependent_work": work,
                       "valid_results": f"{n_valid_all}/{B * world}"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": e2e_ms / args.steps, "valid_results": f"{e2e_valid}/{B}",
                    "layout": "every image a separate pinned host buffer (one pointer per image, batched copy call); the complete "
                              "TStereoOdometryResult of every estimator comes back: result struct, out_residual, outliers, "
                              "per-octave counters (svo_wait_batch_detail)",
                    "variants": variants or None,
                    "class_api": klass,
                    "per_rank": {"e2e_ms_per_step": [round(r[0], 4) for r in per_rank],
                                 "concurrent_h2d_copy_ms_per_step": [round(r[1], 4) for r in per_rank],
                                 "device_ms_per_step": [round(r[2], 4) for r in per_rank],
                                 "e2e_over_copy_ceiling": [round(r[1] / r[0], 4) if r[0] else None for r in per_rank],
                                 "note": "copy ceiling: after a barrier every rank copies one step's pinned image bytes host -> device "
                                         "at the same time (eight times); e2e_over_copy_ceiling = that time / the e2e step time"},
                    "numa": numa},
            "gpu_launches": launches_all,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": dom_gbs, "peak": peak, "unit": "GB/s",
                         "frac": dom_gbs / peak, "traffic": traffic, "peak_source": peak_src,
                         "achieved_is": "compulsory bytes of SURVEY section 8(d) (the image read once: 2*B*W*H per launch) / CUDA-event "
                                        "time of the kernel in the serialised pass of this run",
                         "traffic_is": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full "
                                       "capture (static, not measured in this run): " + str(prof.get("source")),
                         "binding_resource": binding,
                         "note": "this path is not HBM-bound anywhere (SURVEY 8(d)): the dominant kernel is instruction-issue bound, "
                                 "so the HBM fraction is structurally small; binding_resource carries the fraction of the roof that binds"},
            "kernels": kernels,
            "match_stage": {"pair_distances_per_s": pairdist / (hm_ms * 1e-3) if hm_ms else None,
                            "int_pipe_roof_pair_distances_per_s": int_roof,
                            "frac_of_int_pipe_roof": (pairdist / (hm_ms * 1e-3) / int_roof) if hm_ms else None,
                            "measured_popc_lane_ops_per_s": pipe["popc_per_s"], "measured_lop3_lane_ops_per_s": pipe["lop3_per_s"],
                            "ops_per_pair_distance": {"popc": HM_POPC_PFX, "alu": HM_ALU_PFX,
                                                      "note": "this kernel's own algorithm: a 128-bit prefix per pair; pairs whose prefix "
                                                              "beats the best so far (threshold-initialised) get the second half.  The roof "
                                                              "is the integer-pipe roof of THAT operation count, not of the naive 8 POPC"},
                            "hbm_gbs": (ab["hamming_kernel(lr)"] + ab["hamming_kernel(track)"]) / (hm_ms * 1e-3) / 1e9 if hm_ms else None,
                            "hbm_frac": (ab["hamming_kernel(lr)"] + ab["hamming_kernel(track)"]) / (hm_ms * 1e-3) / 1e9 / peak if hm_ms else None},
            "match_microbench": micro,
            "c3": c3,
            "outlier_heavy": outl,
            "dmorb": dmorb,
            "stage1": stage1,
            "single_pair": latency,
            "octaves": octaves,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            procs = os.cpu_count() or 1
            v, info = cpu_throughput(seqs, cam, args.cpu_pairs, procs)
            line["cpu_baseline"] = dict(info, value=v, unit=UNIT, variants=cpu_single_process_variants(seqs, cam))
        emit(line)
    else:
        ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line on the real stdout (everything else, NCCL's version banner included, goes to stderr)"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                    # libraries that print to stdout (NCCL banner) must not pollute the JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="stereo pairs (sequence slots) per step per GPU")
    ap.add_argument("--cpu-pairs", type=int, default=300, help="cpu_baseline: pairs per process")
    ap.add_argument("--ref-pairs", type=int, default=100, help="--impl reference: pairs per process per step")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the rank to the CPUs next to its GPU")
    ap.add_argument("--no-micro", action="store_true", help="skip the microbenches (match c5 + sweep, stage 1, single-pair latency)")
    ap.add_argument("--match-method", type=int, default=0, help="0 = smDescBF (headline), 1 = smDescRbR (row-by-row band matcher)")
    ap.add_argument("--if-match-method", type=int, default=0, help="0 = ifmDescBF (headline), 1 = ifmDescWin (windowed tracker)")
    ap.add_argument("--chunks", type=int, default=0, help="groups of slots on concurrent streams (0 = library default)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"], help="c2 = the headline pipeline; c5 = Hamming microbench")
    ap.add_argument("--c5-reps", type=int, default=64, help="c5: independent 4096x4096 problems per GPU per step")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c5":
        run_c5(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
