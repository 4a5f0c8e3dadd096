"""Probe: wall time of the pose chain of one window (16 frames x 64 streams) run ALONE (nothing else on the GPU), fused vs
split (MOVFE_POSE_SPLIT). env STEPS."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mov-slam_b200", "python"))
import numpy as np
import bench
from movfe import lib, synth, types as T

S, F = bench.S_PER_GPU, bench.F
STEPS = int(os.environ.get("STEPS", 4))
LA = bench.MAX_REF + 1
clips = bench.make_clips(F * (STEPS + 1) + LA, n_base=4)
ctx = lib.Context(S, bench.W, bench.H, max_records_per_frame=bench.MAX_RECORDS, max_ref=bench.MAX_REF, window_frames=F,
                  max_tracks=bench.MAX_TRACKS, max_map_points=2048, has_grey=True)
ctx.set_camera(clips[0]["spec"].camera(), T.pose_params(), 0.5)
w = bench.pack_window(clips, S, 0, F + LA, pinned=False)
ctx.push_frames(w["n"], w["recs"].numpy()[:w["n_records"] * 40].view(T.MV_RECORD), w["off"].numpy(), w["flags"].numpy(), w["grey"].numpy())
ctx.raster(0, F); ctx.extract(0, F)
for b in range(len(clips)):
    sp = clips[b]["spec"]
    mp = synth.map_from_tracks(sp, ctx.tracks(b, 0), synth.pose_at(sp, 0))
    for s in range(b, S, len(clips)):
        ctx.set_map_points(s, mp, len(mp) // 2); ctx.set_pose(s, synth.pose_struct(synth.pose_at(sp, 0)))
ctx.track_poses(0, F); ctx.synchronize()
for k in range(STEPS):
    f0 = F * (k + 1) + LA
    w = bench.pack_window(clips, S, f0, f0 + F, pinned=False)
    ctx.push_frames(F, w["recs"].numpy()[:w["n_records"] * 40].view(T.MV_RECORD), w["off"].numpy(), w["flags"].numpy(), w["grey"].numpy())
    first = F * (k + 1)
    ctx.raster(first, F); ctx.extract(first, F); ctx.synchronize()
    t0 = time.perf_counter()
    ctx.track_poses(first, F); ctx.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps(dict(step=k, pose_chain_ms=round(dt * 1e3, 3), per_frame_us=round(dt * 1e6 / F, 1),
                          median_inliers=float(np.median(ctx.poses(first, F)[1])))), flush=True)
