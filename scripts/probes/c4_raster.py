"""Probe: raster stage on config C4 (1920x1080, one 4x4 record per cell = 129 600 records / frame, ref = 0, MV-only mode):
stage times of ingest / hop lists / slot grid and the slot-grid kernel's HBM rate. env: S (streams, default 16), F (8)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mov-slam_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from movfe import lib, synth, types as T
from gpu_util import pack_streams

S, F = int(os.environ.get("S", 16)), int(os.environ.get("F", 8))
W, H, NW = 1920, 1080, 3
M = (W // 4) * (H // 4)
n_frames = F * NW + 1
specs = [synth.Spec(W, H, n_frames=n_frames, refs=1, seed=0x5EED0400 + b, fx=960.0, fy=960.0, dense4x4=True, start_p=True, phase=0.2 * b)
         for b in range(2)]
clips = [synth.make_records(sp) for sp in specs]
per_stream = [clips[s % 2] for s in range(S)]
ctx = lib.Context(S, W, H, max_records_per_frame=M, max_ref=0, window_frames=F, max_tracks=8192, has_grey=False, serial_raster=True)
for s in range(S):
    seeds = synth.seed_tracks_lattice(specs[s % 2])
    ctx.set_tracks(s, seeds, int(seeds["track_id"].max()))
r, o, fl = pack_streams(per_stream, n_frames, 0, F + 1)
ctx.push_frames(F + 1, r, o, fl, None)
ctx.profile_enable(True)
for k in range(NW):
    first = F * k
    ctx.raster(first, F)
    ctx.extract(first, F)
    if k + 1 < NW:
        r, o, fl = pack_streams(per_stream, n_frames, F * (k + 1) + 1, F * (k + 2) + 1)
        ctx.push_frames(F, r, o, fl, None)
    ms, ln = ctx.profile_read()
    gb = S * F * W * H * 16 / 1e9
    print(json.dumps(dict(window=k, streams=S, frames=F, ms={a: round(b, 3) for a, b in ms.items()},
                          grid_GBs=round(gb / (ms["grid"] / 1e3), 1), frames_per_s_raster_extract=round(S * F / ((ms["hops"] + ms["grid"] + ms["extract"]) / 1e3)),
                          tracks=int(ctx.track_count(0, first + F - 1)[0]))), flush=True)
