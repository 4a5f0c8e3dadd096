#!/usr/bin/env python
"""Wall time of movfe_search_by_projection for a batch of 64 frames x 4000 keypoints x 1024 projected map points (host arrays in,
host arrays out, as the operator is called).  python scripts/probes/search_timing.py   (one JSON line)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "mov-slam_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from movfe import lib, types as T  # noqa: E402
from search_util import make_frame  # noqa: E402

rng = np.random.Generator(np.random.PCG64(7))
S, NF, NP = 64, 4000, 1024
base = [make_frame(rng, 640, 480, NF, NP) for _ in range(4)]
frames = [base[i % 4] for i in range(S)]
cat = lambda i: np.concatenate([f[i] for f in frames])
feat, pts, proj, desc = cat(0), cat(1), cat(2), cat(3)
foff, poff = np.arange(S + 1, dtype=np.int32) * NF, np.arange(S + 1, dtype=np.int32) * NP
prm = np.zeros(1, T.PROJECTION_SEARCH)
prm["th"], prm["th_high"], prm["nn_ratio"] = 1.0, 60, 0.8
if os.environ.get("PROBE_PINNED", "1") != "0":   # page-locked inputs (what a caller that cares about the copies would hand over)
    import torch
    pin = lambda a: torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).pin_memory().numpy().view(a.dtype).reshape(a.shape)
    feat, pts, proj, desc = pin(feat), pin(pts), pin(proj), pin(desc)
ctx = lib.Context(1, 640, 480, max_records_per_frame=64, max_ref=0, window_frames=1, max_tracks=64, max_map_points=16, has_grey=False)
ctx.profile_enable(False)
for _ in range(3):
    out = ctx.search_by_projection(feat, foff, pts, proj, desc, poff, prm)
t0 = time.perf_counter()
R = 20
for _ in range(R):
    out = ctx.search_by_projection(feat, foff, pts, proj, desc, poff, prm)
dt = (time.perf_counter() - t0) / R
print(json.dumps({"probe": "movfe_search_by_projection, %d frames x %d keypoints x %d map points, host arrays in and out" % (S, NF, NP),
                  "ms_per_call": round(dt * 1e3, 3), "frames_per_s": round(S / dt), "map_points_per_s": round(S * NP / dt), "matches": int(out[3].sum()),
                  "pinned_inputs": os.environ.get("PROBE_PINNED", "1") != "0", "h2d_mb": round((feat.nbytes + pts.nbytes + proj.nbytes + desc.nbytes) / 1e6, 1)}), flush=True)
