"""Dev probe: facts that size the kernels (tracks per frame, kps per frame, GN iterations, H2D bandwidth)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mov-slam_b200", "python"))
import numpy as np
import torch
import bench
from movfe import lib, synth, types as T

S, F, STEPS = 64, 16, 4
LA = bench.MAX_REF + 1
clips = bench.make_clips(F * (STEPS + 1) + LA, n_base=4)
ctx = lib.Context(S, bench.W, bench.H, max_records_per_frame=bench.MAX_RECORDS, max_ref=bench.MAX_REF, window_frames=F,
                  max_tracks=bench.MAX_TRACKS, max_map_points=2048, has_grey=True)
cam = clips[0]["spec"].camera()
ctx.set_camera(cam, T.pose_params(), 0.5)
w = bench.pack_window(clips, S, 0, F + LA, pinned=False)
ctx.push_frames(w["n"], w["recs"].numpy()[:w["n_records"] * 40].view(T.MV_RECORD), w["off"].numpy(), w["flags"].numpy(), w["grey"].numpy())
ctx.raster(0, F); ctx.extract(0, F)
for b in range(len(clips)):
    sp = clips[b]["spec"]
    mp = synth.map_from_tracks(sp, ctx.tracks(b, 0), synth.pose_at(sp, 0))
    for s in range(b, S, len(clips)):
        ctx.set_map_points(s, mp, len(mp) // 2); ctx.set_pose(s, synth.pose_struct(synth.pose_at(sp, 0)))
ctx.track_poses(0, F); ctx.synchronize()
ctx.profile_enable(True)
for k in range(STEPS):
    f0 = F * (k + 1) + LA
    w = bench.pack_window(clips, S, f0, f0 + F, pinned=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.push_frames(F, w["recs"].numpy()[:w["n_records"] * 40].view(T.MV_RECORD), w["off"].numpy(), w["flags"].numpy(), w["grey"].numpy())
    ctx.synchronize(); t_push = time.perf_counter() - t0
    first = F * (k + 1)
    ctx.raster(first, F); ctx.extract(first, F); ctx.track_poses(first, F)
    ms, ln = ctx.profile_read()
    nt = [ctx.track_count(0, first + i)[0] for i in range(F)]
    nk = [ctx.raster_counts(0, first + i)[:2] for i in range(F)]
    m, o = ctx.matches(0, first + F - 1)
    print(json.dumps(dict(step=k, push_ms=round(t_push * 1e3, 2), h2d_GBps=round((w["n_records"] * 40 + w["grey"].numel()) / t_push / 1e9, 1),
                          ms={a: round(b, 3) for a, b in ms.items()}, tracks=nt, hops_kps=nk, matched=int((m >= 0).sum()), inl=int(((m >= 0) & (o == 0)).sum()))), flush=True)
# GN iteration statistics on bench-like pose problems
pp = T.pose_params()
for n in (60, 450, 2000):
    pts, obs, pgt, pin = synth.pnp_problem(n, cam, 7 + n)
    poses, outl, ninl, stats = ctx.pose_optimize(cam, pp, pts, obs, np.array([0, n], np.int32), np.array([pin]))
    print("pose n=%d inliers=%d stats[its,rounds,passes,fail]=%s" % (n, ninl[0], stats[0].tolist()), flush=True)
# pure pinned H2D bandwidth
x = torch.empty(400 << 20, dtype=torch.uint8, pin_memory=True); y = torch.empty_like(x, device="cuda")
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); y.copy_(x, non_blocking=True); torch.cuda.synchronize()
    print("pinned H2D 400 MiB: %.1f GB/s" % (x.numel() / (time.perf_counter() - t0) / 1e9), flush=True)
