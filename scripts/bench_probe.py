"""Dev probe: the bench workload (C2, 64 streams x 16 frames per step, max_tracks as bench.py) for ncu launch lists and
captures, without the CPU baseline / e2e arms. env: STEPS (default 3)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mov-slam_b200", "python"))
import numpy as np
import bench
from movfe import lib, synth, types as T

S, F = bench.S_PER_GPU, bench.F
STEPS = int(os.environ.get("STEPS", 3))
LA = bench.MAX_REF + 1
clips = bench.make_clips(F * (STEPS + 1) + LA, n_base=4)
ctx = lib.Context(S, bench.W, bench.H, max_records_per_frame=bench.MAX_RECORDS, max_ref=bench.MAX_REF, window_frames=F,
                  max_tracks=bench.MAX_TRACKS, max_map_points=2048, has_grey=True, serial_raster=bool(os.environ.get("SERIAL_RASTER")))
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
ctx.profile_enable(True)
for k in range(STEPS):
    f0 = F * (k + 1) + LA
    w = bench.pack_window(clips, S, f0, f0 + F, pinned=False)
    ctx.push_frames(F, w["recs"].numpy()[:w["n_records"] * 40].view(T.MV_RECORD), w["off"].numpy(), w["flags"].numpy(), w["grey"].numpy())
    first = F * (k + 1)
    ctx.raster(first, F); ctx.extract(first, F); ctx.track_poses(first, F)
    ms, ln = ctx.profile_read()
    print(json.dumps(dict(step=k, ms={a: round(b, 3) for a, b in ms.items()}, tracks=[ctx.track_count(0, first + i)[0] for i in (0, F - 1)])), flush=True)
