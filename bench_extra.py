"""bench_extra.py — the two BASELINE.json configurations that are not stream-batched front-end runs:

  C1  mono 640x480, 300-frame textured-plane clip, ref=1, ONE stream on the CPU - the reference's own CPU-runnable case.
      Timed: (a) the reference's OWN decoder + extractor loop (oracle/_ref: src/VideoDecoder.cc + src/MOVExtractor.cc compiled
      unmodified, one core), (b) the oracle port's whole front-end on one core (adds joins / frustum / pose), (c) when a GPU
      is present, the same single stream through the CUDA path (latency mode: S = 1).
  C5  PoseOptimization stress: 20 000 map points per problem, KannalaBrandt8, 128 problems per GPU through
      movfe_pose_optimize; HBM fraction by SURVEY.md 8d's B_pose = I * 20 * P and an FP64 ALU fraction beside it.

Called by bench.py --config C1|C5; prints ONE JSON line in the bench contract's shape.
"""
import os
import sys
import time

import numpy as np

import bench
from movfe import synth, types as T


def run(args):
    if args.config == "C1":
        return run_c1(args)
    return run_c5(args)


# ------------------------------------------------------------------------------------------------------ C1 -----
def run_c1(args):
    from oracle import pyoracle as orc
    W, H, NF = 640, 480, int(os.environ.get("BENCH_C1_FRAMES", 300))
    spec = synth.Spec(W, H, n_frames=NF, refs=1, seed=0x5EED0001)
    recs, off, flags = synth.make_records(spec)
    grey = synth.make_grey(spec)
    cam, pp = spec.camera(), T.pose_params()
    # local map: refreshed every 16 frames from the oracle's own tables, as in the batched configs
    cfg = dict(bench.CONFIGS["C2"], max_ref=0, refs=1, F=16, name="C1")
    clip = orc.Clip(W, H, recs, off, flags, 0)
    prev, cid, tabs = np.zeros(0, T.TRACK), 0, {}
    for f in range(NF):
        prev, _, cid, _ = orc.extract_frame(W, H, flags[f], grey[f], clip.grid(f), clip.hops(f), clip.kps(f), clip.coverage(f), prev, cid,
                                            max_tracks=8192)
        if f == 0 or (f + 1) % 16 == 0:
            tabs[f] = prev
    mp0, nkf0 = bench.local_map(cfg, spec, tabs[0], 0)
    sched = [(f + 1,) + bench.local_map(cfg, spec, tabs[f], f) for f in sorted(tabs) if f > 0 and f + 1 < NF]
    t0 = time.perf_counter()
    res = orc.frontend_run(W, H, recs, off, flags, grey, None, mp0, synth.pose_struct(synth.pose_at(spec, 0)), cam, pp, max_ref=0,
                           max_tracks=8192, n_kf_points=nkf0, map_schedule=sched)
    port_s = time.perf_counter() - t0
    line = {"metric": bench.METRIC, "unit": "frames/s", "n_gpus": 0, "steps": 1, "warmup": 0, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": bench.DTYPE, "data": "synthetic",
            "config": {"workload": "C1: mono 640x480 (TartanAir.yaml intrinsics), %d-frame textured-plane clip, ref=1, 1 stream" % NF},
            "cpu_port_1core": {"value": NF / port_s, "unit": "frames/s", "cores": 1, "kind": "port",
                               "what": "whole front-end (raster, propagation, joins, frustum, 2 x pose) of the oracle port",
                               "mean_inliers": float(np.mean(res["n_inliers"][16:]))}}
    try:
        from oracle import pyref
        if pyref.available("canon"):
            r = pyref.frontend_run(W, H, recs, off, flags, grey, variant="canon")
            same = bool(np.array_equal(r["track_hash"], res["track_hash"]))
            line["cpu_baseline"] = {"value": NF / (r["seconds_decoder"] + r["seconds_extractor"]), "unit": "frames/s", "cores": 1, "kind": "reference",
                                    "sample": "the reference's own VideoDecoder::NextImage (fake libav in front) + MOVExtractor::operator() over the "
                                              "%d frames, one core: %.2f ms decoder (MV loop, cv::Mat fill) + %.2f ms extractor per frame; "
                                              "joins and pose are not in this figure (OpenCV's solvePnPRansac is absent here)" %
                                              (NF, 1e3 * r["seconds_decoder"] / NF, 1e3 * r["seconds_extractor"] / NF),
                                    "tables_equal_the_port": same}
    except Exception as e:  # the GPU box carries the prebuilt libraries; without them the port figure stands alone
        line["cpu_baseline_error"] = str(e)
    value = None
    try:
        from movfe import lib
        import torch
        if torch.cuda.is_available():
            F, LA = 16, 1
            ctx = lib.Context(1, W, H, max_records_per_frame=4800, max_ref=0, window_frames=F, max_tracks=8192, max_map_points=1024,
                              output_grid=False)
            ctx.set_camera(cam, pp, 0.5)
            ctx.set_map_points(0, mp0, nkf0)
            ctx.set_pose(0, synth.pose_struct(synth.pose_at(spec, 0)))
            sd = dict((f, (m, k)) for f, m, k in sched)
            t0 = time.perf_counter()
            pushed = 0
            for first in range(0, NF - NF % F, F):
                want = min(NF, first + F + LA)
                if want > pushed:
                    ctx.push_frames(want - pushed, recs[off[pushed]:off[want]], off[pushed:want + 1] - off[pushed], flags[pushed:want],
                                    grey[None, pushed:want])
                    pushed = want
                if first in sd:
                    ctx.set_map_points(0, sd[first][0], sd[first][1])
                ctx.raster(first, F)
                ctx.extract(first, F)
                ctx.track_poses(first, F)
            P, ninl = ctx.poses(NF - NF % F - F, F)
            gpu_s = time.perf_counter() - t0
            value = (NF - NF % F) / gpu_s
            line["n_gpus"] = 1
            line["single_stream_gpu"] = {"value": value, "unit": "frames/s", "what": "the same stream alone on one B200 through the C-ABI, host buffers, "
                                         "windows of 16 frames (latency mode: the batch dimension is 1)"}
            ctx.close()
    except Exception as e:
        line["gpu_error"] = str(e)
    line["value"] = value if value is not None else line["cpu_port_1core"]["value"]
    line["ms_per_step"] = 1e3 * NF / line["value"]
    bench.emit(line)


# ------------------------------------------------------------------------------------------------------ C5 -----
def run_c5(args):
    import torch
    from movfe import lib
    from oracle import pyoracle as orc
    NP, P = int(os.environ.get("BENCH_C5_PROBLEMS", 128)), int(os.environ.get("BENCH_C5_POINTS", 20000))
    cam = T.camera(190.0, 190.0, 376.0, 240.0, k=(-0.01, 0.002, -0.0005, 0.0001), model=T.CAM_FISHEYE)
    pp = T.pose_params()
    base = [synth.pnp_problem(P, cam, 0x5EED0005 + 13 * b, width=752, height=480) for b in range(8)]
    pts = np.concatenate([base[i % 8][0] for i in range(NP)])
    obs = np.concatenate([base[i % 8][1] for i in range(NP)])
    off = (np.arange(NP + 1) * P).astype(np.int32)
    init = np.array([base[i % 8][3] for i in range(NP)], T.POSE)
    ctx = lib.Context(1, 752, 480, max_records_per_frame=16, max_ref=0, window_frames=1, max_tracks=16, max_map_points=16, has_grey=False)
    for _ in range(max(args.warmup, 1)):
        poses, outl, ninl, stats = ctx.pose_optimize(cam, pp, pts, obs, off, init)
    ctx.profile_enable(True)
    ctx.profile_read(reset=True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        poses, outl, ninl, stats = ctx.pose_optimize(cam, pp, pts, obs, off, init)
    wall = (time.perf_counter() - t0) / args.steps
    ms, _ = ctx.profile_read()
    kern_ms = ms["pose"] / args.steps
    passes = float(np.mean(stats[:, 2]))
    peak, peak_src = bench.peaks()
    b_pose = NP * (passes * 20.0 * P + P / 8.0 + 64)
    flop64 = NP * passes * P * 180.0           # about 180 double-precision operations per correspondence and pass
    # CPU port on the same problems (one per host thread)
    cores = os.cpu_count() or 1
    import threading
    n_cpu = min(NP, cores)
    t0 = time.perf_counter()
    ths = [threading.Thread(target=lambda i=i: orc.pose_optimize(cam, pp, base[i % 8][0], base[i % 8][1], base[i % 8][3])) for i in range(n_cpu)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    cpu_s = time.perf_counter() - t0
    wn, wpose, woutl, wstats = orc.pose_optimize(cam, pp, base[0][0], base[0][1], base[0][3])
    rel = max(float(np.max(np.abs(poses[0][k] - wpose[k])) / max(1.0, float(np.max(np.abs(wpose[k]))))) for k in ("R", "t"))
    line = {"metric": "pose_optimizations_per_s", "value": NP / (kern_ms / 1e3), "unit": "problems/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": kern_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (f32 inputs)", "data": "synthetic",
            "config": {"workload": "C5: Optimizer::PoseOptimization stress, %d problems x %d map points, KannalaBrandt8 (fx=fy=190, k = -0.01, 0.002, -0.0005, 0.0001), "
                                   "sigma 0.5 px, 10 %% gross outliers" % (NP, P), "frames_per_s_equivalent": NP / 2 / (kern_ms / 1e3)},
            "e2e": {"value": NP / wall, "unit": "problems/s", "h2d_bytes_per_step": int(pts.nbytes + obs.nbytes + init.nbytes + off.nbytes),
                    "d2h_bytes_per_step": int(outl.nbytes + poses.nbytes + ninl.nbytes + stats.nbytes)},
            "roofline": {"bound": "hbm", "kernel": "pose_kernel", "achieved": b_pose / 1e9 / (kern_ms / 1e3), "peak": peak, "unit": "GB/s",
                         "frac": b_pose / 1e9 / (kern_ms / 1e3) / peak, "peak_source": peak_src, "traffic": None,
                         "passes_per_problem": passes, "fp64_tflops": flop64 / 1e12 / (kern_ms / 1e3),
                         "fp64_frac_of_40_tflops": flop64 / 1e12 / (kern_ms / 1e3) / 40.0,
                         "note": "SURVEY.md 8d: B_pose = I * 20 * P + P/8 + 64 per problem; arithmetic intensity ~9 flop/B puts the kernel at the FP64 ridge"},
            "parity_check": {"pose_rel": rel, "inliers_equal": bool(int(ninl[0]) == wn), "outliers_equal": bool(np.array_equal(outl[:P], woutl))},
            "cpu_baseline": {"value": n_cpu / cpu_s, "unit": "problems/s", "cores": n_cpu, "kind": "port", "sample": "%d of the problems, one per host thread, %.1fs" % (n_cpu, cpu_s)}}
    ctx.close()
    bench.emit(line)
