#!/usr/bin/env python
"""bench.py -- frames/s of the SLAM frontend hot path (extract + consecutive-frame match) on B200.

Contract (one JSON line on stdout from rank 0):
  python bench.py --gpus N --steps K --warmup W            our CUDA path (this repo)
  python bench.py --impl reference ...                     the reference's CPU algorithm on the host cores
For N > 1 launch through torch.distributed.run (one rank per GPU, NCCL).

A "step" is one pass of the hot path over one batch of `--frames` synthetic frames per GPU:
detectAndCompute on every frame, then match(f, f+1) with keypoints for every consecutive pair.
Workload = BASELINE.json configs[1] shape: KITTI-shape synthetic grayscale 1241x376.

  value    : whole-job frames/s with the frames already resident in HBM (device-timed, max over ranks)
  e2e      : same metric through the public host API: pinned host frames -> H2D -> extract -> match ->
             D2H of keypoints, descriptors, matches and counts, every step inside the timed region
  roofline : dominant kernel of the step, per-launch duration from CUDA events recorded on the
             launching stream inside the timed region (slamcu_profile_*), against MEASURED_PEAKS.json
             (HBM-bound kernels) or the popc ceiling measured in this run (the matcher)
  cpu_baseline : the CPU oracle (oracle/ref_frontend.cpp, a port of the reference) on a bounded sample
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

ROWS, COLS = 376, 1241
GRID_PITCH = 14  # synthetic scene density; the achieved keypoint count is reported
METRIC = "frames/s ORB extract+match @1241x376, 2k kp"
UNIT = "frames/s"
WORKLOAD = "KITTI-shape synthetic grayscale sequence 1241x376 (BASELINE.json configs[1]: 1000 frames, 2000 ORB keypoints/frame, 8-level pyramid)"
NFEATURES = 2000
RANSAC_K4 = None  # (fx, fy, cx, cy): also run findEssentialMat(RANSAC) on every consecutive pair
# The default (`kitti`) is the configuration BASELINE.json's metric is quoted on; the other two are BASELINE.json's
# configs[2] / configs[3] at full size, selectable for extra measured lines (ORB mode only).
WORKLOADS = {
    "kitti": dict(rows=376, cols=1241, pitch_px=14, nfeatures=2000, frames=1000, max_kp=2560, chunk=500, metric=METRIC, text=WORKLOAD, k4=None),
    "tum": dict(rows=480, cols=640, pitch_px=17, nfeatures=1000, frames=1000, max_kp=1280, chunk=500,
                metric="frames/s ORB extract+match+findEssentialMat @640x480, 1k kp",
                text="TUM-RGB-D-shape 640x480 synthetic sequence (BASELINE.json configs[2]: 1000 keypoints/frame, consecutive-frame matching + RANSAC essential matrix per pair)",
                k4=(525.0, 525.0, 319.5, 239.5)),
    "4k": dict(rows=2160, cols=3840, pitch_px=28, nfeatures=10000, frames=48, max_kp=12288, chunk=24,
               metric="frames/s ORB extract+match @3840x2160, 10k kp",
               text="4K 3840x2160 synthetic sequence (BASELINE.json configs[3]: 10000 keypoints/frame, sharded by frame range across the GPUs)",
               k4=None),
}


def make_frames(n, seed):
    from slam_cin0051_b200.synth import make_sequence
    out = np.empty((n, ROWS, COLS), np.uint8)
    done, g = 0, 0
    while done < n:  # a new scene every 16 frames (the crop offsets have period 16)
        m = min(16, n - done)
        out[done:done + m] = make_sequence(ROWS, COLS, m, GRID_PITCH, seed=1000 * seed + g)
        done += m
        g += 1
    return out


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_run(frames, threads):
    from oracle import ref_oracle
    ref_oracle.build()
    sec, counts = ref_oracle.frontend_run(frames, with_kp=True, threads=threads)
    return sec, counts


def cpu_orb_run(frames, threads, nfeatures=None, ratio=0.75):
    """OpenCV's own ORB + BFMatcher(k=2) + ratio test (+ findEssentialMat when the workload has it) on the host cores,
    frame-parallel (cv2 releases the GIL).  Returns (seconds, counts[n, 2] = keypoints, matches(f, f+1))."""
    import concurrent.futures as cf

    import cv2
    cv2.setNumThreads(1)
    n = len(frames)
    nfeatures = nfeatures or NFEATURES
    k4 = RANSAC_K4

    def extract(i):
        orb = cv2.ORB_create(nfeatures=nfeatures, scaleFactor=1.2, nlevels=8, edgeThreshold=31, firstLevel=0, WTA_K=2,
                             scoreType=cv2.ORB_HARRIS_SCORE, patchSize=31, fastThreshold=20)
        k, d = orb.detectAndCompute(frames[i], None)
        pts = np.float32([p.pt for p in k]).reshape(-1, 2)
        return (d if d is not None else np.zeros((0, 32), np.uint8)), pts

    def match(i):
        (d1, p1), (d2, p2) = feats[i], feats[i + 1]
        if len(d1) == 0 or len(d2) < 2:
            return 0
        m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(d1, d2, k=2)
        good = [a for a, b in m if not (a.distance >= np.float32(ratio) * np.float32(b.distance))]
        if k4 is not None and len(good) >= 8:  # PoseEstimator::estimate's guard (pose_estimator.cpp:22-26)
            K = np.array([[k4[0], 0, k4[2]], [0, k4[1], k4[3]], [0, 0, 1.0]])
            cv2.findEssentialMat(p1[[a.queryIdx for a in good]], p2[[a.trainIdx for a in good]], K, cv2.RANSAC, 0.999, 1.0, 1000)
        return len(good)

    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(max_workers=threads) as ex:
        feats = list(ex.map(extract, range(n)))
        nm = list(ex.map(match, range(n - 1))) + [0]
    sec = time.perf_counter() - t0
    return sec, np.array([[len(f[0]), m] for f, m in zip(feats, nm)], np.int64)


MODE_TEXT = {
    "orb": "OpenCV-ORB-compatible: 8-level pyramid, FAST-9 + Harris, keypoint budget per frame as the workload names, rBRIEF-256, BF Hamming kNN k=2 + ratio 0.75",
    "reference": "reference algorithm (FAST-12 + SAD NMS + BRIEF + BF Hamming with keypoint penalty)",
}


def run_reference(args, rank, world):
    if rank != 0:
        return 0
    threads = host_threads()
    orb = args.mode == "orb"
    # bounded sample per step: ~0.1-0.15 s of single-core work per frame -> a few seconds per step
    n = max(threads, min(args.frames, 4 * threads))
    frames = make_frames(n, 0)
    run = cpu_orb_run if orb else cpu_reference_run
    for _ in range(args.warmup):
        run(frames[: max(2, n // 4)], threads)
    t = 0.0
    for _ in range(args.steps):
        sec, counts = run(frames, threads)
        t += sec
    value = n * args.steps / t
    if orb:
        import cv2
        kind, how = "reference", f"cv2 {cv2.__version__} ORB_create({NFEATURES}).detectAndCompute + BFMatcher(NORM_HAMMING).knnMatch(k=2) + ratio 0.75, frame-parallel thread pool, cv2.setNumThreads(1) per call"
    else:
        kind, how = "port", "frame-parallel std::thread pool, oracle/ref_frontend.cpp (g++ -O2)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "mode": MODE_TEXT[args.mode],
                   "frames_per_step": n, "keypoints_per_frame_mean": float(counts[:, 0].mean())},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{n} frames/step x {args.steps} steps, {how}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


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


def orb_level_pixels(rows, cols, nlevels=8, sf=1.2):
    out = []
    for l in range(nlevels):
        s = np.float32(np.power(np.float64(np.float32(sf)), float(l)))
        out.append(int(np.rint(np.float32(rows) * (np.float32(1) / s))) * int(np.rint(np.float32(cols) * (np.float32(1) / s))))
    return out


def kernel_model(name, frames, px, pitch_px, n_raw, n_kp, cmp_per_step, orb_stats=None):
    """Algorithmic work per STEP (DESIGN.md section 'Kernels'): bytes for HBM-bound kernels, comparisons for the
    matcher.  Kernels launched once per pyramid level are summed over the levels."""
    mask = px / 8
    if orb_stats is not None:
        lp = orb_level_pixels(ROWS, COLS)
        P = float(sum(lp))
        n_cand, n_sel = orb_stats
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
            "match_finalize": ("hbm", frames * (16 * n_kp + 12 * n_kp)),
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
    }
    return table.get(name, ("hbm", 0.0))


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import slam_cin0051_b200 as S

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: this framework has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B = args.frames
    ctx = S.Context(local_rank)
    stream = torch.cuda.Stream()  # a real (non-null) stream shared by torch's events and the library's launches
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    data = os.path.join(ROOT, "test", "data")
    orb = args.mode == "orb"
    sfx = "_orb" if orb else ""
    with_kp = not orb  # the reference's matcher applies its image-distance penalty; BFMatcher has none
    det_cfg = os.path.join(data, f"feature_detector{sfx}.yml")
    if orb and NFEATURES != 2000:
        from slam_cin0051_b200.config import read_yaml
        det_cfg = {**read_yaml(det_cfg), "MaxFeatures": NFEATURES}
    det = S.FeatureDetector(det_cfg, ctx)
    mat = S.FeatureMatcher(os.path.join(data, f"feature_matcher{sfx}.yml"), ctx)
    max_kp = args.max_keypoints
    seq = S.FrameSequence(ROWS, COLS, B, desc_bytes=det.descriptor_bytes, max_keypoints=max_kp, context=ctx)

    frames_np = make_frames(B, rank)
    host_frames = torch.empty((B, ROWS, COLS), dtype=torch.uint8, pin_memory=True)
    host_frames.numpy()[:] = frames_np
    # pinned result buffers for the end-to-end leg
    h_kps = torch.empty((B, max_kp, 5), dtype=torch.float32, pin_memory=True)
    h_desc = torch.empty((B, max_kp, det.descriptor_bytes), dtype=torch.uint8, pin_memory=True)
    h_matches = torch.empty((B, max_kp, 3), dtype=torch.int32, pin_memory=True)
    h_counts = torch.empty((B, 4), dtype=torch.int32, pin_memory=True)

    def barrier():
        if world > 1:
            dist.barrier()

    def step_resident():
        seq.extract(det, 0, B)
        seq.match_consecutive(mat, 0, B - 1, with_keypoints=with_kp)
        if RANSAC_K4:
            seq.essential(RANSAC_K4, 0, B - 1)

    # end-to-end leg: two sequences double-buffer, so step k+1's H2D runs under step k's kernels (a streaming
    # deployment); every step still uploads its frames from pinned host memory and downloads all its results
    seq2 = S.FrameSequence(ROWS, COLS, B, desc_bytes=det.descriptor_bytes, max_keypoints=max_kp, context=ctx)
    outs = [(h_kps, h_desc, h_matches, h_counts)]
    outs.append(tuple(torch.empty_like(x).pin_memory() for x in outs[0]))
    seqs = [seq, seq2]

    def submit_e2e(i):
        k, d, m, c = outs[i % 2]
        seqs[i % 2].process_ptrs(det, mat, host_frames.data_ptr(), B, chunk=args.chunk, with_keypoints=with_kp,
                                 kps_ptr=k.data_ptr(), desc_ptr=d.data_ptr(), matches_ptr=m.data_ptr(), counts_ptr=c.data_ptr())
        if RANSAC_K4:  # E, inlier masks and counts stay on the device (read per pair with essential_result)
            seqs[i % 2].essential(RANSAC_K4, 0, B - 1)

    def collect_e2e(i):
        seqs[i % 2].wait()  # the step's results are now in host memory; read them
        c = outs[i % 2][3]
        return int(c[:, 0].sum()), int(c[:, 1].sum())

    seq.upload_ptr(host_frames.data_ptr(), B)
    for _ in range(max(args.warmup, 3)):
        step_resident()
    ctx.synchronize()
    counts = seq.counts()
    if (counts[:, 3] != 0).any():
        raise RuntimeError(f"device list overflow (status bits {np.unique(counts[:, 3])}); raise capacities")
    n_kp_mean = float(counts[:, 0].mean())
    n_raw_mean = float(counts[:, 2].mean())
    cmp_per_step = float((counts[:-1, 0].astype(np.int64) * counts[1:, 0].astype(np.int64)).sum())

    # ---- timed region: resident inputs ----
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = ctx.launch_count
    ctx.profile_enable(True)
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
    e1.record(stream)
    torch.cuda.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    launches = ctx.launch_count - launches0

    # ---- end-to-end leg (host buffers, copies inside the timed region) ----
    for i in range(2):
        submit_e2e(i)
    ctx.synchronize()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tot = (0, 0)
    for i in range(args.steps):
        submit_e2e(i)
        if i >= 1:
            tot = collect_e2e(i - 1)
    tot = collect_e2e(args.steps - 1)
    ctx.synchronize()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if sampler else None  # sampled across both timed regions

    t_res = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_res, op=dist.ReduceOp.MAX)
        # the only inter-GPU traffic of the path: per-frame counts gathered over NCCL/NVLink
        gathered = [torch.empty((B, 4), dtype=torch.int32, device="cuda") for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(counts).cuda())
        n_kp_mean = float(torch.stack(gathered)[:, :, 0].float().mean().item())
    ms, e2e_ms = float(t_res[0].item()), float(t_res[1].item())

    if rank == 0:
        value = world * B * args.steps / (ms * 1e-3)
        e2e_value = world * B * args.steps / (e2e_ms * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        gpopc = ctx.popc_peak()
        px = ROWS * COLS
        kernels = []
        total_k_ms = sum(v[0] for v in prof.values()) or 1.0
        for name, (kms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
            bound, work = kernel_model(name, B, px, GRID_PITCH, n_raw_mean, n_kp_mean, cmp_per_step,
                                       (n_raw_mean, 2.0 * n_kp_mean) if orb else None)
            per_launch_ms = kms / max(cnt, 1)
            ms_step = kms / args.steps  # all launches of this kernel in one step (one per pyramid level for some)
            ach = work / (ms_step * 1e-3) / 1e9
            common = {"kernel": name, "ms_per_launch": per_launch_ms, "launches_per_step": cnt / args.steps,
                      "ms_per_step": ms_step, "share": kms / total_k_ms, "achieved": ach}
            if bound == "hbm":
                kernels.append({**common, "bound": "hbm", "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak})
            else:
                # POPC instructions per comparison: 4 with the carry-save tree (ORB mode, 256 populated bits); the
                # reference's descriptors populate 46 bits, the kernel skips the all-zero words -> 2
                peak = gpopc / (4.0 if orb else 2.0)
                kernels.append({**common, "bound": "popc", "peak": peak, "unit": "Gcmp/s (256-bit)", "frac": ach / peak})
        try:
            ncu = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_dram_per_frame.json")))["dram_bytes_per_frame"] if orb and (ROWS, COLS, NFEATURES) == (376, 1241, 2000) else {}
        except OSError:
            ncu = {}
        for k in kernels:  # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu --set full capture
            per_frame = ncu.get(k["kernel"])
            k["traffic"] = per_frame * B / max(k["launches_per_step"], 1) if per_frame else None
            k["algorithmic_bytes_per_launch"] = (kernel_model(k["kernel"], B, px, GRID_PITCH, n_raw_mean, n_kp_mean, cmp_per_step,
                                                              (n_raw_mean, 2.0 * n_kp_mean) if orb else None)[1]
                                                 / max(k["launches_per_step"], 1)) if k["bound"] == "hbm" else None
        def roof(k):
            if not k:
                return {}
            src = hbm_src if k["bound"] == "hbm" else (f"integer pipe: POPC issue ceiling measured in this run ({gpopc:.0f} Gpopc/s) / "
                                                       f"{4 if orb else 2} POPC per comparison" + ("" if orb else " (46 populated bits; the float distance penalty, not POPC, limits this path)"))
            return {"kernel": k["kernel"], "bound": k["bound"], "achieved": k["achieved"], "peak": k["peak"], "unit": k["unit"],
                    "frac": k["frac"], "traffic": k.get("traffic"), "peak_source": src, "share_of_step": k["share"],
                    "ms_per_launch": k["ms_per_launch"]}
        roofline = roof(kernels[0] if kernels else None)
        hbm_kernels = [k for k in kernels if k["bound"] == "hbm"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "mode": MODE_TEXT[args.mode],
                       "frames_per_step_per_gpu": B, "pairs_per_step_per_gpu": B - 1, "keypoints_per_frame_mean": n_kp_mean,
                       "raw_corners_per_frame_mean": n_raw_mean, "parallelism": f"frame-range sharding x{world}",
                       "l2": f"inputs larger than L2: {B * ROWS * COLS / 1e6:.0f} MB of frames per step vs 126 MB L2"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(B * ROWS * COLS),
                    "d2h_bytes_per_step": int(h_kps.nbytes + h_desc.nbytes + h_matches.nbytes + h_counts.nbytes),
                    "ms_per_step": e2e_ms / args.steps, "keypoints_last_step": tot[0], "matches_last_step": tot[1],
                    "pipeline_chunk_frames": args.chunk, "double_buffered_sequences": 2},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "roofline_hbm": roof(hbm_kernels[0] if hbm_kernels else None),  # the dominant HBM-side kernel
            "kernels": kernels,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = host_threads()
            n_s = max(8, min(B, 4 * threads))
            run = cpu_orb_run if orb else cpu_reference_run
            run(frames_np[:4], threads)  # warm
            sec, c_cpu = run(frames_np[:n_s], threads)
            sec1, _ = run(frames_np[:4], 1)
            same = bool((c_cpu[:, 0] == counts[:n_s, 0]).all() and (c_cpu[:-1, 1] == counts[:n_s - 1, 1]).all())
            if orb:
                import cv2
                kind = "reference"
                what = f"cv2 {cv2.__version__} ORB_create({NFEATURES}) + BFMatcher.knnMatch(k=2) + ratio 0.75{' + findEssentialMat(RANSAC) per pair' if RANSAC_K4 else ''} (the OpenCV path BASELINE.json names)"
            else:
                kind, what = "port", "oracle/ref_frontend.cpp (port of the reference's src/frontend, g++ -O2)"
            line["cpu_baseline"] = {"value": n_s / sec, "unit": UNIT, "cores": threads, "kind": kind,
                                    "sample": f"first {n_s} frames of the same batch, frame-parallel over {threads} host threads; {what}",
                                    "single_thread_value": 4 / sec1, "counts_match_gpu": same}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="orb", choices=["orb", "reference"],
                    help="orb: the OpenCV-ORB-compatible path BASELINE.json's headline config names; "
                         "reference: the reference repo's own hand-written detector/matcher")
    ap.add_argument("--workload", default="kitti", choices=sorted(WORKLOADS),
                    help="kitti: BASELINE.json configs[1], the configuration the metric is quoted on (default); tum / 4k: configs[2] / configs[3]")
    ap.add_argument("--frames", type=int, default=None, help="frames per GPU per step (default: the workload's; kitti = a 1000-frame sequence)")
    ap.add_argument("--max-keypoints", type=int, default=None)
    ap.add_argument("--chunk", type=int, default=None, help="frames per pipeline stage of the end-to-end leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    global ROWS, COLS, GRID_PITCH, METRIC, WORKLOAD, NFEATURES, RANSAC_K4
    w = WORKLOADS[args.workload]
    if args.workload != "kitti" and args.mode != "orb":
        ap.error("--workload tum / 4k are ORB-mode workloads (the reference algorithm has no keypoint budget)")
    ROWS, COLS, GRID_PITCH, METRIC, WORKLOAD, NFEATURES, RANSAC_K4 = w["rows"], w["cols"], w["pitch_px"], w["metric"], w["text"], w["nfeatures"], w["k4"]
    args.frames = args.frames or w["frames"]
    args.max_keypoints = args.max_keypoints or w["max_kp"]
    args.chunk = args.chunk or w["chunk"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    return run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
