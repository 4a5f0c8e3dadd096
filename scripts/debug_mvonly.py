import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mov-slam_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from movfe import synth, types as T
from oracle import pyoracle as orc
from gpu_util import run_raster_clip, assert_raster_equal, run_frontend_clip, oracle_tracks
specs = [synth.Spec(640, 480, n_frames=9, refs=4, seed=0x5EED0013 + s, start_p=True) for s in range(2)]
streams = [synth.make_records(sp) for sp in specs]
seeds = [synth.seed_tracks_lattice(sp) for sp in specs]
want = oracle_tracks(orc, streams[0], 640, 480, 3, seeds=seeds[0])
clip = orc.Clip(640, 480, *streams[0], 3)
for window in (9,):
    got, _, ctx = run_frontend_clip(streams, 640, 480, 9, window, 3, seeds=seeds)
    print("window", window, [(len(got[(0, f)]), len(want[f])) for f in range(9)])
    if window == 9:
        f = 4
        g, w = got[(0, f)], want[f]
        gi = set(g["track_id"].tolist()); wi = set(w["track_id"].tolist())
        miss = sorted(wi - gi)
        prev = want[f - 1]
        G = ctx.grid(0, f); H = ctx.hops(0, f)
        for tid in miss[:3]:
            p = prev[prev["track_id"] == tid][0]
            x, y = int(p["pt_x"]), int(p["pt_y"])
            sl = G[y, x]; osl = clip.grid(f)[y, x]
            print(" tid", tid, "pt", p["pt_x"], p["pt_y"], "slots gpu", sl, "oracle", osl, "hop", H[sl[0]], clip.hops(f)[osl[0]])
            d = H[sl[0]]["d_indx"]
            # who else goes to d
            users = []
            for q in prev:
                xx, yy = int(q["pt_x"]), int(q["pt_y"])
                s0 = clip.grid(f)[yy, xx][0]
                if s0 >= 0 and clip.hops(f)[s0]["d_indx"] == d: users.append((int(q["track_id"]), float(q["pt_x"]), float(q["pt_y"])))
            print("   d_indx", d, "users", users)
    ctx.close()
