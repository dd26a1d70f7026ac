#!/usr/bin/env python
"""bench.py -- frames/s of the SLAM frontend hot path (extract + consecutive-frame match) on B200.

Contract (one JSON line on stdout from rank 0):
  python bench.py --gpus N --steps K --warmup W            our CUDA path (this repo)
  python bench.py --impl reference ...                     the reference's CPU algorithm on the host cores
For N > 1 launch through torch.distributed.run (one rank per GPU, NCCL).

Workload = BASELINE.json configs[1] shape: KITTI-shape synthetic grayscale 1241x376, 2000 ORB keypoints, 8 levels.
A "step" is one pass of the hot path over ONE sequence of N x `--frames` frames, sharded by contiguous frame range:
rank r owns frames [r F, (r + 1) F) and re-extracts frame (r + 1) F (a 1-frame halo, recomputed, never exchanged) so that
every pair (f, f + 1) it owns is local; detectAndCompute on every frame, match(f, f + 1) on every pair, then ONE NCCL
all-gather of the per-frame counts {keypoints, matches, raw corners, status} -- the path's only exchange --
inside the timed region.  Per-GPU work is fixed as N grows ("scaling": "weak").

  value    : whole-job frames/s with the frames already resident in HBM (device-timed with CUDA events, max over ranks)
  e2e      : same metric through the public host API: pinned host frames -> H2D -> extract -> match -> device
             compaction of the results into dense pinned buffers (only the defined rows cross the link) -> count
             all-gather, every step inside the timed region
  roofline : dominant kernel of the step, per-launch duration from CUDA events recorded on the launching stream over a
             second pass of the same K steps (slamcu_profile_*), against MEASURED_PEAKS.json (HBM-bound kernels) or the
             popc ceiling measured in this run (the matcher)
  cpu_baseline : cv2's own ORB + BFMatcher (ORB mode) or the reference's frontend (oracle/_ref when built, else its
             C++ port) on a bounded sample, frame-parallel over all host cores in separate processes
Extra legs (N = 1, default invocation only, outside the headline's timed region): the reference's own algorithm
(`mode_reference`), BASELINE configs[2] with RANSAC (`config3`) and the configs[4] Hamming sweep (`hamming_sweep`).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "frames/s"
# The default (`kitti`) is the configuration BASELINE.json's metric is quoted on; the other two are BASELINE.json's
# configs[2] / configs[3] at full size, selectable for extra measured lines (ORB mode only).
WORKLOADS = {
    "kitti": dict(rows=376, cols=1241, pitch_px=14, nfeatures=2000, frames=1000, max_kp=2560, chunk=500, k4=None,
                  metric="frames/s ORB extract+match @1241x376, 2k kp",
                  text="KITTI-shape synthetic grayscale sequence 1241x376 (BASELINE.json configs[1]: 1000 frames, 2000 ORB keypoints/frame, 8-level pyramid)"),
    "tum": dict(rows=480, cols=640, pitch_px=17, nfeatures=1000, frames=1000, max_kp=1280, chunk=500, k4=(525.0, 525.0, 319.5, 239.5),
                metric="frames/s ORB extract+match+findEssentialMat @640x480, 1k kp",
                text="TUM-RGB-D-shape 640x480 synthetic sequence (BASELINE.json configs[2]: 1000 keypoints/frame, consecutive-frame matching + RANSAC essential matrix per pair)"),
    "4k": dict(rows=2160, cols=3840, pitch_px=28, nfeatures=10000, frames=48, max_kp=12288, chunk=24, k4=None,
               metric="frames/s ORB extract+match @3840x2160, 10k kp",
               text="4K 3840x2160 synthetic sequence (BASELINE.json configs[3]: 10000 keypoints/frame, sharded by frame range across the GPUs)"),
}
MODE_TEXT = {
    "orb": "OpenCV-ORB-compatible: 8-level pyramid, FAST-9 + Harris, keypoint budget per frame as the workload names, rBRIEF-256, BF Hamming kNN k=2 + ratio 0.75",
    "reference": "reference algorithm (FAST-12 + SAD NMS + BRIEF + BF Hamming with keypoint penalty)",
}
SCENE = 16  # frames per synthetic scene (the crop offsets have period 16); a pair across a scene cut has no true matches


def make_frames(w, lo, n):
    """Frames [lo, lo + n) of the global synthetic sequence of workload `w` (scene g = frames [16 g, 16 g + 16), seed g)."""
    from slam_cin0051_b200.synth import make_sequence
    out = np.empty((n, w["rows"], w["cols"]), np.uint8)
    f = lo
    while f < lo + n:
        g, k = divmod(f, SCENE)
        m = min(SCENE - k, lo + n - f)
        out[f - lo:f - lo + m] = make_sequence(w["rows"], w["cols"], SCENE, w["pitch_px"], seed=g)[k:k + m]
        f += m
    return out


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def config_of(args, w):
    """Identical on both arms (the product and --impl reference): it names the workload, not a run."""
    return {"workload": w["text"], "mode": MODE_TEXT[args.mode], "frames_per_step_per_gpu": args.frames, "keypoint_budget": w["nfeatures"],
            "rows": w["rows"], "cols": w["cols"], "ransac": bool(w["k4"]),
            "parallelism": "one sequence of n_gpus x frames_per_step_per_gpu frames sharded by contiguous frame range, 1 recomputed halo frame per "
                           "range boundary, one NCCL all-gather of per-frame counts per step"}


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the OpenCV path BASELINE.json names (ORB mode) / the reference's own frontend (reference mode)
# ---------------------------------------------------------------------------------------------------------------------
_W = {}


def _cpu_orb_range(job):
    """One worker process: frames [lo, hi) + halo of the shared batch -> ORB, then match(f, f+1) for its pairs.
    Plain cv2 API calls, the way a user of OpenCV writes this loop (Lowe's ratio idiom on the DMatch objects)."""
    import cv2
    lo, hi, last = job
    frames, nfeatures, k4, ratio = _W["frames"], _W["nfeatures"], _W["k4"], 0.75
    cv2.setNumThreads(1)
    orb = cv2.ORB_create(nfeatures=nfeatures, scaleFactor=1.2, nlevels=8, edgeThreshold=31, firstLevel=0, WTA_K=2,
                         scoreType=cv2.ORB_HARRIS_SCORE, patchSize=31, fastThreshold=20)
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    K = None if k4 is None else np.array([[k4[0], 0, k4[2]], [0, k4[1], k4[3]], [0, 0, 1.0]])
    feats = []
    for f in range(lo, min(hi + 1, last)):  # + the halo frame, recomputed like the GPU ranks do
        k, d = orb.detectAndCompute(frames[f], None)
        feats.append((k, d if d is not None else np.zeros((0, 32), np.uint8)))
    out = []
    for f in range(lo, hi):
        k1, d1 = feats[f - lo]
        nm = 0
        if f + 1 < last:
            k2, d2 = feats[f + 1 - lo]
            if len(d1) and len(d2) >= 2:
                good = [m for m, n in bf.knnMatch(d1, d2, k=2) if not (m.distance >= ratio * n.distance)]
                nm = len(good)
                if K is not None and nm >= 8:  # PoseEstimator::estimate's guard (pose_estimator.cpp:22-26)
                    p1 = cv2.KeyPoint_convert(k1, [m.queryIdx for m in good])
                    p2 = cv2.KeyPoint_convert(k2, [m.trainIdx for m in good])
                    cv2.findEssentialMat(p1, p2, K, cv2.RANSAC, 0.999, 1.0, 1000)
        out.append((len(k1), nm))
    return lo, out


def cpu_orb_run(frames, workers, w):
    """OpenCV's own ORB + BFMatcher(k=2) + ratio test (+ findEssentialMat when the workload has it) on the host cores:
    `workers` PROCESSES (no shared interpreter lock), each a contiguous frame range.  Returns (seconds, counts[n, 2])."""
    _W.update(frames=frames, nfeatures=w["nfeatures"], k4=w["k4"])
    return _pool_run(_cpu_orb_range, frames, workers)


def _cpu_ref_range(job):
    """One worker process: the reference's OWN frontend (oracle/_ref/libslam_ref.so: its unmodified feature_detector.cpp and
    feature_matcher.cpp) on frames [lo, hi) + halo: detectAndCompute, then match(f, f+1) with keypoints."""
    from oracle import ref_build
    lo, hi, last = job
    frames = _W["frames"]
    feats = [ref_build.detect_and_compute(frames[f]) for f in range(lo, min(hi + 1, last))]
    out = []
    for f in range(lo, hi):
        k1, d1 = feats[f - lo]
        nm = 0
        if f + 1 < last:
            k2, d2 = feats[f + 1 - lo]
            if len(d1) and len(d2):
                nm = len(ref_build.match(d1, d2, k1, k2)[0])
        out.append((len(k1), nm))
    return lo, out


def _pool_run(fn, frames, workers):
    import multiprocessing as mp
    n = len(frames)
    workers = max(1, min(workers, n))
    bounds = [round(i * n / workers) for i in range(workers + 1)]
    jobs = [(bounds[i], bounds[i + 1], n) for i in range(workers) if bounds[i + 1] > bounds[i]]
    counts = np.zeros((n, 2), np.int64)
    if workers == 1:
        t0 = time.perf_counter()
        res = [fn(jobs[0])]
        sec = time.perf_counter() - t0
    else:
        with mp.get_context("fork").Pool(workers) as pool:  # workers inherit the frames; the pool is up before the clock starts
            pool.map(int, range(workers))
            t0 = time.perf_counter()
            res = pool.map(fn, jobs, chunksize=1)
            sec = time.perf_counter() - t0
    for lo, out in res:
        counts[lo:lo + len(out)] = out
    return sec, counts


def cpu_reference_run(frames, workers, w):
    """The reference's own frontend on the host cores: its unmodified sources (oracle/_ref, kind "reference") when that
    library is there, else their C++ port (oracle/ref_frontend.cpp, kind "port")."""
    from oracle import ref_build
    if ref_build.available():
        ref_build.lib()
        _W.update(frames=frames)
        return _pool_run(_cpu_ref_range, frames, workers)
    from oracle import ref_oracle
    ref_oracle.build()
    sec, counts = ref_oracle.frontend_run(frames, with_kp=True, threads=workers)
    return sec, counts[:, :2].astype(np.int64)


def cpu_kind(args, w):
    if args.mode == "orb":
        import cv2
        return "reference", (f"cv2 {cv2.__version__} ORB_create({w['nfeatures']}).detectAndCompute + BFMatcher(NORM_HAMMING).knnMatch(k=2) + ratio 0.75"
                             f"{' + findEssentialMat(RANSAC) per pair' if w['k4'] else ''} (the OpenCV path BASELINE.json names), one process per host core, "
                             "cv2.setNumThreads(1) each")
    from oracle import ref_build
    if ref_build.available():
        return "reference", ("oracle/_ref/libslam_ref.so: the reference's unmodified src/frontend/feature_detector.cpp + feature_matcher.cpp (g++ -O2, header stand-ins "
                             "for Eigen / OpenCV-core / spdlog), detectAndCompute + match(f, f+1) with keypoints, one process per host core")
    return "port", "oracle/ref_frontend.cpp (C++ port of the reference's src/frontend, pinned bit for bit to the reference's own sources by tests/test_ref_build.py), frame-parallel std::thread pool, g++ -O2"


def cpu_baseline_leg(args, w, frames, gpu_counts=None):
    """Bounded sample on all host cores + a single-core run of part of it (scaling evidence)."""
    threads = host_threads()
    run = cpu_orb_run if args.mode == "orb" else cpu_reference_run
    n_s = len(frames)
    run(frames[: max(2, min(n_s, threads))], threads, w)  # warm (page cache, cv2 init)
    sec, c_cpu = run(frames, threads, w)
    n1 = max(4, min(8, n_s))
    sec1, _ = run(frames[:n1], 1, w)
    half = max(1, threads // 2)
    sec_h, _ = run(frames[: max(half, n_s // 2)], half, w)
    kind, what = cpu_kind(args, w)
    out = {"value": n_s / sec, "unit": UNIT, "cores": threads, "kind": kind,
           "sample": f"first {n_s} frames of the same sequence, frame-parallel over {threads} host cores; {what}",
           "single_thread_value": n1 / sec1, "scaling_vs_linear": (n_s / sec) / (threads * n1 / sec1),
           "half_cores": {"workers": half, "value": max(half, n_s // 2) / sec_h, "scaling_vs_linear": (max(half, n_s // 2) / sec_h) / (half * n1 / sec1)},
           "scaling_note": "separate processes (no shared interpreter lock, no Python glue inside the timed calls beyond the cv2 / reference API itself); what is "
                           "left of the gap to linear is the host: compare half_cores (the box's vCPUs are not all independent cores)"}
    if gpu_counts is not None:
        m = min(n_s, len(gpu_counts))
        out["counts_match_gpu"] = bool((c_cpu[:m, 0] == gpu_counts[:m, 0]).all() and (c_cpu[:m - 1, 1] == gpu_counts[:m - 1, 1]).all())
    return out


def run_reference(args, w, rank, world):
    if rank != 0:
        return 0
    threads = host_threads()
    # bounded sample per step: ~0.1 s of single-core work per frame -> a few seconds per step
    n = max(threads, min(args.frames, 8 * threads))
    frames = make_frames(w, 0, n)
    run = cpu_orb_run if args.mode == "orb" else cpu_reference_run
    for _ in range(args.warmup):
        run(frames[: max(2, n // 4)], threads, w)
    t = 0.0
    counts = np.zeros((n, 2))
    for _ in range(args.steps):
        sec, counts = run(frames, threads, w)
        t += sec
    value = n * args.steps / t
    kind, how = cpu_kind(args, w)
    line = {
        "impl": "reference", "metric": w["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config_of(args, w),
        "stats": {"frames_per_step_sampled": n, "keypoints_per_frame_mean": float(counts[:, 0].mean())},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{n} frames/step x {args.steps} steps of the same sequence; {how}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(gpu_index)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        rows = [l.strip().split(", ") for l in open(self.tmp.name) if l.strip()]
        os.unlink(self.tmp.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def bind_to_gpu_numa(local_rank):
    """Pins this rank (and therefore the first-touch placement of its pinned buffers) to the CPUs next to its GPU."""
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        before = sorted(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        after = sorted(os.sched_getaffinity(0))
        if not after:
            os.sched_setaffinity(0, before)
            after = before
        info = {"bound": after != before, "cpus": f"{after[0]}-{after[-1]} ({len(after)})", "all_cpus": len(before)}
        try:
            bus = pynvml.nvmlDeviceGetPciInfo(h).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            node = open(f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node").read().strip()
            info["gpu_numa_node"] = int(node)
        except Exception:
            pass
    except Exception as e:  # no NVML / not permitted: run unbound
        info["error"] = repr(e)[:80]
    return info


def orb_level_pixels(rows, cols, nlevels=8, sf=1.2):
    out = []
    for l in range(nlevels):
        s = np.float32(np.power(np.float64(np.float32(sf)), float(l)))
        out.append(int(np.rint(np.float32(rows) * (np.float32(1) / s))) * int(np.rint(np.float32(cols) * (np.float32(1) / s))))
    return out


def kernel_model(name, w, frames, n_raw, n_kp, n_match, cmp_per_step, orb):
    """Algorithmic work per STEP: bytes for HBM-bound kernels, comparisons for the matcher.
    Kernels launched once per pyramid level are summed over the levels."""
    px = w["rows"] * w["cols"]
    mask = px / 8
    if orb:
        lp = orb_level_pixels(w["rows"], w["cols"])
        P = float(sum(lp))
        n_cand, n_sel = n_raw, 2.0 * n_kp
        table = {
            "pyr_down": ("hbm", frames * float(sum(lp[l - 1] + lp[l] for l in range(1, len(lp))))),
            "fast9_mask": ("hbm", frames * (P + P / 8)),
            "orb_select": ("hbm", frames * (P / 8 + 12 * n_cand + 16 * n_cand + 4 * n_sel)),
            "harris": ("hbm", frames * (81 * n_sel + 8 * n_sel)),
            "orb_retain": ("hbm", frames * (8 * n_sel + 8 * n_kp)),
            "orb_assemble": ("hbm", frames * (8 * n_kp + 28 * n_kp)),
            "blur7": ("hbm", frames * (2 * P)),
            "orb_describe": ("hbm", frames * ((709 + 512) * n_kp + 44 * n_kp)),
            "desc_or": ("hbm", frames * (32 * n_kp)),
            "match": ("popc", cmp_per_step),
            "match_expand": ("hbm", frames * ((32 + 256 + 4) * n_kp)),  # 32 B descriptor -> 256 operand bytes + one key constant
            "match_finalize": ("hbm", frames * (16 * n_kp + 12 * n_match)),
            "pack_counts": ("hbm", frames * 32.0),
        }
        return table.get(name, ("hbm", 0.0))
    table = {
        "fast_mask": ("hbm", frames * (px + mask)),
        "corner_list": ("hbm", frames * (mask + 17 * n_raw + 8 * n_raw)),
        "sort": ("hbm", frames * (8 * n_raw)),
        "nms": ("hbm", frames * (8 * n_raw + 20 * n_kp)),
        "blur5": ("hbm", frames * (2 * px)),
        "describe": ("hbm", frames * (px + 52 * n_kp)),
        "desc_or": ("hbm", frames * (32 * n_kp)),
        "match": ("popc", cmp_per_step),
        "match_finalize": ("hbm", frames * (16 * n_kp + 12 * 20)),
        "pack_counts": ("hbm", frames * 32.0),
    }
    return table.get(name, ("hbm", 0.0))


def run_ours(args, w, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import slam_cin0051_b200 as S

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: this framework has no CPU fallback")
    all_cpus = sorted(os.sched_getaffinity(0))
    numa = bind_to_gpu_numa(local_rank) if world > 1 else {"bound": False}
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ROWS, COLS = w["rows"], w["cols"]
    from slam_cin0051_b200.sharding import frame_range
    B = args.frames                         # frames every rank owns
    lo, hi, halo = frame_range(world * B, rank, world)  # contiguous range + 1 halo frame (the next range's first, re-extracted here)
    assert hi - lo == B
    NF = B + halo
    NP = NF - 1                             # local pairs: (f, f + 1) for every owned f but the sequence's last frame
    ctx = S.Context(local_rank)
    stream = torch.cuda.Stream()  # a real (non-null) stream shared by torch's events / NCCL and the library's launches
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    data = os.path.join(ROOT, "test", "data")
    orb = args.mode == "orb"
    sfx = "_orb" if orb else ""
    with_kp = not orb  # the reference's matcher applies its image-distance penalty; BFMatcher has none
    det_cfg = os.path.join(data, f"feature_detector{sfx}.yml")
    if orb and w["nfeatures"] != 2000:
        from slam_cin0051_b200.config import read_yaml
        det_cfg = {**read_yaml(det_cfg), "MaxFeatures": w["nfeatures"]}
    det = S.FeatureDetector(det_cfg, ctx)
    mat = S.FeatureMatcher(os.path.join(data, f"feature_matcher{sfx}.yml"), ctx)
    max_kp = args.max_keypoints
    k4 = w["k4"]
    seq = S.FrameSequence(ROWS, COLS, NF, desc_bytes=det.descriptor_bytes, max_keypoints=max_kp, context=ctx)

    frames_np = make_frames(w, lo, NF)
    host_frames = torch.empty((NF, ROWS, COLS), dtype=torch.uint8, pin_memory=True)
    host_frames.numpy()[:] = frames_np
    # dense pinned result buffers of the end-to-end leg (capacity = every frame at the keypoint cap; only the defined rows move)
    kp_cap, m_cap = NF * max_kp, NF * max_kp

    def out_set():
        return (torch.empty((kp_cap, 5), dtype=torch.float32, pin_memory=True), torch.empty((kp_cap, det.descriptor_bytes), dtype=torch.uint8, pin_memory=True),
                torch.empty((m_cap, 3), dtype=torch.int32, pin_memory=True), torch.empty((NF, 4), dtype=torch.int32, pin_memory=True))

    # the step's only exchange: per-frame counts of the owned frames, all-gathered over NCCL on the compute stream
    d_counts = torch.zeros((B, 4), dtype=torch.int32, device="cuda")
    g_counts = torch.zeros((world * B, 4), dtype=torch.int32, device="cuda")

    def gather_counts(s):
        s.counts_device(d_counts.data_ptr(), 0, B)
        if world > 1:
            dist.all_gather_into_tensor(g_counts, d_counts)
        else:
            g_counts.copy_(d_counts, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()

    def step_resident():
        seq.extract_match(det, mat, 0, NF, with_keypoints=with_kp, chunk=args.lane_chunk)  # detectAndCompute + match(f, f+1)
        if k4:
            seq.essential(k4, 0, NP)
        gather_counts(seq)

    # end-to-end leg: N_SEQ sequences rotate and the host stays AHEAD steps in front of the results it collects, so step
    # k+2's H2D runs under step k+1's kernels while step k's results are still being downloaded (a streaming deployment);
    # every step still uploads its frames from pinned host memory and downloads all its results
    N_SEQ, AHEAD = 3, 2
    seqs = [seq] + [S.FrameSequence(ROWS, COLS, NF, desc_bytes=det.descriptor_bytes, max_keypoints=max_kp, context=ctx) for _ in range(N_SEQ - 1)]
    outs = [out_set() for _ in range(N_SEQ)]

    submit_s = [0.0]

    def submit_e2e(i):
        t_sub = time.perf_counter()
        k, d, m, c = outs[i % N_SEQ]
        seqs[i % N_SEQ].process_dense_ptrs(det, mat, host_frames.data_ptr(), NF, chunk=args.chunk, with_keypoints=with_kp,
                                       kps_ptr=k.data_ptr(), desc_ptr=d.data_ptr(), matches_ptr=m.data_ptr(), counts_ptr=c.data_ptr(),
                                       kp_capacity=kp_cap, match_capacity=m_cap)
        if k4:  # E, inlier masks and counts stay on the device (read per pair with essential_result)
            seqs[i % N_SEQ].essential(k4, 0, NP)
        gather_counts(seqs[i % N_SEQ])
        submit_s[0] += time.perf_counter() - t_sub

    def collect_e2e(i):
        seqs[i % N_SEQ].wait()  # the step's results are now in host memory (raises if a device list overflowed); read them
        c = outs[i % N_SEQ][3]
        return int(c[:B, 0].sum()), int(c[:B, 1].sum()), int(c[:, 0].sum()), int(c[:NP, 1].sum())

    seq.upload_ptr(host_frames.data_ptr(), NF)
    for _ in range(max(args.warmup, 3)):
        step_resident()
    ctx.synchronize()
    counts = seq.counts()
    if (counts[:, 3] != 0).any():
        raise RuntimeError(f"device list overflow (status bits {np.unique(counts[:, 3])}); raise capacities")
    n_kp_mean = float(counts[:B, 0].mean())
    n_raw_mean = float(counts[:B, 2].mean())
    n_match_mean = float(counts[:NP, 1].mean())
    cmp_per_step = float((counts[:-1, 0].astype(np.int64) * counts[1:, 0].astype(np.int64)).sum())

    def timed(fn):
        barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        barrier()
        return e0.elapsed_time(e1)

    # ---- timed region 1: resident inputs, exactly as a user runs it (the blurred pyramid overlaps the corner chain) ----
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = ctx.launch_count
    ms = timed(step_resident)
    launches = ctx.launch_count - launches0
    # ---- the same K steps again with per-kernel CUDA events on the launching stream (serialises the two streams) ----
    ctx.profile_enable(True)
    ms_prof = timed(step_resident)
    prof = ctx.profile_read()
    ctx.profile_enable(False)

    # ---- end-to-end leg (host buffers, copies inside the timed region) ----
    n_warm = N_SEQ * max(args.warmup, 3)  # W untimed steps per sequence, results collected like the timed ones
    for i in range(n_warm):
        submit_e2e(i)
        if i >= AHEAD:
            collect_e2e(i - AHEAD)
    for i in range(max(n_warm - AHEAD, 0), n_warm):
        collect_e2e(i)
    ctx.synchronize()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    submit_s[0] = 0.0
    tot = (0, 0, 0, 0)
    for i in range(args.steps):
        submit_e2e(i)
        if i >= AHEAD:
            tot = collect_e2e(i - AHEAD)
    for i in range(max(args.steps - AHEAD, 0), args.steps):
        tot = collect_e2e(i)
    ctx.synchronize()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if sampler else None  # sampled across the timed regions
    g_host = g_counts.cpu().numpy()

    t_res = torch.tensor([ms, e2e_s * 1e3, ms_prof], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_res, op=dist.ReduceOp.MAX)
    ms, e2e_ms, ms_prof = (float(x) for x in t_res.tolist())

    if rank == 0:
        total_frames = world * B
        value = total_frames * args.steps / (ms * 1e-3)
        e2e_value = total_frames * args.steps / (e2e_ms * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        gpopc = ctx.popc_peak()
        kernels = []
        total_k_ms = sum(v[0] for v in prof.values()) or 1.0
        try:
            ncu = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_dram_per_frame.json")))["dram_bytes_per_frame"] if orb and args.workload == "kitti" else {}
        except OSError:
            ncu = {}
        # the penalty-free 256-bit search runs as an integer GEMM on the tensor cores (match_tc.cu); its operand-widening
        # kernel shows up as "match_expand" when it does
        match_on_tensor = "match_expand" in prof
        tensor_peak = 2.0 * float(peaks.get("bf16_tflops", 1664.2))  # u8 dense = 2 x bf16 dense on this part; bf16 is the measured figure
        sm_mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
        for name, (kms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
            bound, work = kernel_model(name, w, NF, n_raw_mean, n_kp_mean, n_match_mean, cmp_per_step, orb)
            if name == "match" and match_on_tensor:
                bound = "tensor"
            lps = cnt / args.steps
            ms_step = kms / args.steps  # all launches of this kernel in one step (one per pyramid level for some)
            ach = work / (ms_step * 1e-3) / 1e9
            k = {"kernel": name, "ms_per_launch": kms / max(cnt, 1), "launches_per_step": lps, "ms_per_step": ms_step, "share": kms / total_k_ms,
                 "achieved": ach, "bound": bound}
            if bound == "hbm":
                k.update(peak=hbm_peak, unit="GB/s", frac=ach / hbm_peak, algorithmic_bytes_per_launch=work / max(lps, 1))
            elif bound == "tensor":
                # one comparison = 256 multiply-adds = 512 operations of the u8 GEMM <q, t>; the epilogue reads one 32-bit
                # accumulator per comparison out of TMEM (64 B / clock / SM, B300_MICROARCH.md) -- the tighter of the two ceilings
                tops = ach * 512.0 / 1e3
                tmem_ceiling = 64.0 / 4.0 * 148 * sm_mhz * 1e6 / 1e9  # G comparisons / s
                k.update(achieved=tops, peak=tensor_peak, unit="TOP/s (u8 dense)", frac=tops / tensor_peak, algorithmic_bytes_per_launch=None,
                         gcmp_s=ach, tmem_read_ceiling_gcmp_s=tmem_ceiling, tmem_read_frac=ach / tmem_ceiling)
            else:
                # POPC instructions per comparison: 4 with the carry-save tree (ORB mode, 256 populated bits); the
                # reference's descriptors populate 46 bits, the kernel skips the all-zero words -> 2
                peak = gpopc / (4.0 if orb else 2.0)
                k.update(peak=peak, unit="Gcmp/s (256-bit)", frac=ach / peak, algorithmic_bytes_per_launch=None)
            per_frame = ncu.get(name)  # dram__bytes_read.sum + dram__bytes_write.sum from the committed ncu --set full capture
            k["traffic"] = per_frame * NF / max(lps, 1) if per_frame else None
            kernels.append(k)

        def roof(k):
            if not k:
                return {}
            src = hbm_src if k["bound"] == "hbm" else (f"integer pipe: POPC issue ceiling measured in this run ({gpopc:.0f} Gpopc/s) / "
                                                       f"{4 if orb else 2} POPC per comparison" + ("" if orb else " (46 populated bits; the float distance penalty, not POPC, limits this path)"))
            if k["bound"] == "tensor":
                src = "2 x MEASURED_PEAKS.json bf16_tflops (u8 dense issues at twice the bf16 rate; bf16 is the measured figure)"
            return {"kernel": k["kernel"], "bound": k["bound"] if k["bound"] in ("hbm", "tensor") else "popc", "achieved": k["achieved"], "peak": k["peak"], "unit": k["unit"],
                    "frac": k["frac"], "traffic": k.get("traffic"), "peak_source": src, "share_of_step": k["share"],
                    "ms_per_launch": k["ms_per_launch"]}
        hbm_kernels = [k for k in kernels if k["bound"] == "hbm"]
        h2d = int(NF * ROWS * COLS)
        d2h = int(tot[2] * (20 + det.descriptor_bytes) + tot[3] * 12 + NF * 16 + NF * 4)
        step_s = e2e_ms * 1e-3 / args.steps
        line = {
            "metric": w["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": config_of(args, w),
            "stats": {"sequence_frames": total_frames, "frames_extracted_per_gpu": NF, "pairs_per_gpu": NP, "keypoints_per_frame_mean": n_kp_mean,
                      "raw_corners_per_frame_mean": n_raw_mean, "matches_per_pair_mean": n_match_mean,
                      "gathered_keypoints_total": int(g_host[:, 0].sum()), "gathered_matches_total": int(g_host[:, 1].sum()),
                      "l2": f"inputs larger than L2: {NF * ROWS * COLS / 1e6:.0f} MB of frames per GPU per step vs 126 MB L2",
                      "timing": "value: CUDA events around K unprofiled steps (blur pyramid on a forked stream), max over ranks; kernels[]: a second pass "
                                "of the same K steps with events around every launch; e2e: chunks alternate between two compute lanes",
                      "ms_per_step_profiled_pass": ms_prof / args.steps, "collective": "all_gather_into_tensor of int32[frames][4] per step, inside both timed regions" if world > 1 else "none at N=1 (device-side count pack only)",
                      "numa": numa},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps, "keypoints_last_step": tot[0], "matches_last_step": tot[1],
                    "host_submit_ms_per_step": submit_s[0] * 1e3 / args.steps,
                    "pipeline_chunk_frames": args.chunk, "sequences_in_rotation": N_SEQ, "steps_in_flight": AHEAD + 1, "outputs": "dense (device compaction into pinned host memory)",
                    "h2d_gbs_per_rank": h2d / step_s / 1e9, "d2h_gbs_per_rank": d2h / step_s / 1e9,
                    "host_link_gbs_all_ranks": world * (h2d + d2h) / step_s / 1e9},
            "gpu_launches": int(launches),
            "roofline": roof(kernels[0] if kernels else None),
            "roofline_hbm": roof(hbm_kernels[0] if hbm_kernels else None),  # the dominant HBM-side kernel
            "kernels": kernels,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            os.sched_setaffinity(0, all_cpus)
            threads = host_threads()
            n_s = max(8, min(NF, 8 * threads))
            line["cpu_baseline"] = cpu_baseline_leg(args, w, frames_np[:n_s], counts)
        if world == 1 and args.extra_legs:
            del seqs, outs
            line.update(extra_legs(args))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def extra_legs(args):
    """Short measured legs for the other configurations the repo claims numbers for (N = 1, outside the headline's timed
    region): each is this same script in a child process, its JSON line attached under its own key."""
    out = {}

    def child(extra, timeout=600):
        cmd = [sys.executable, os.path.abspath(__file__), "--gpus", "1", "--steps", str(max(2, min(args.steps, 10))), "--warmup", "3", "--no-extra-legs", *extra]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
            lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
            if r.returncode == 0 and lines:
                d = json.loads(lines[-1])
                d.pop("kernels", None)
                return d
            return {"error": (r.stderr or r.stdout)[-300:]}
        except Exception as e:
            return {"error": repr(e)[:300]}
    out["mode_reference"] = child(["--mode", "reference"])
    out["config3"] = child(["--workload", "tum"])
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "hamming_sweep.py"), "--quick"], capture_output=True, text=True, timeout=600)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        out["hamming_sweep"] = json.loads(lines[-1]) if r.returncode == 0 and lines else {"error": (r.stderr or r.stdout)[-300:]}
    except Exception as e:
        out["hamming_sweep"] = {"error": repr(e)[:300]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="orb", choices=["orb", "reference"],
                    help="orb: the OpenCV-ORB-compatible path BASELINE.json's headline config names; "
                         "reference: the reference repo's own hand-written detector/matcher")
    ap.add_argument("--workload", default="kitti", choices=sorted(WORKLOADS),
                    help="kitti: BASELINE.json configs[1], the configuration the metric is quoted on (default); tum / 4k: configs[2] / configs[3]")
    ap.add_argument("--frames", type=int, default=None, help="frames per GPU per step (default: the workload's; kitti = 1000 per GPU)")
    ap.add_argument("--max-keypoints", type=int, default=None)
    ap.add_argument("--chunk", type=int, default=None, help="frames per pipeline stage of the end-to-end leg")
    ap.add_argument("--lane-chunk", type=int, default=0, help="frames per chunk of a two-lane resident step (0: one lane, the faster choice with resident inputs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the mode_reference / config3 / hamming_sweep legs (N = 1 default run)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.workload != "kitti" and args.mode != "orb":
        ap.error("--workload tum / 4k are ORB-mode workloads (the reference algorithm has no keypoint budget)")
    explicit = any(a in sys.argv for a in ("--frames", "--mode", "--workload", "--max-keypoints", "--chunk", "--no-cpu-baseline"))
    args.extra_legs = not args.no_extra_legs and not explicit and args.impl == "ours"
    args.frames = args.frames or w["frames"]
    args.max_keypoints = args.max_keypoints or w["max_kp"]
    args.chunk = args.chunk or w["chunk"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, w, rank, world)
    return run_ours(args, w, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
