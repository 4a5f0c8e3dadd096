#!/usr/bin/env python
"""Regenerates tests/golden/frontend_small.npz — frozen outputs of the CPU oracle on small seeded inputs.

The reference ships no golden vectors and cannot be built in this image, so these fixtures do not pin the
oracle to the reference; they freeze the oracle (itself pinned to the hand-derived KATs of SURVEY.md App. B) so that a
later change to the oracle OR to the CUDA path is caught, and so the GPU box can check the CUDA path without
/root/reference. Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "mov-slam_b200", "python"))
from movfe import synth, types as T  # noqa: E402
from oracle import pyoracle as orc  # noqa: E402

CASES = {
    # name: (Spec kwargs, max_ref, with_grey)
    "textured_ref3": (dict(width=160, height=96, n_frames=7, refs=3, seed=0x5EEDA001, fx=80.0, fy=80.0), 2, True),
    "flat_ref2": (dict(width=128, height=64, n_frames=5, refs=2, seed=0x5EEDA002, fx=64.0, fy=64.0, start_p=True), 1, False),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_case(kw, max_ref, with_grey, max_tracks=1024):
    sp = synth.Spec(**kw)
    recs, off, flags = synth.make_records(sp)
    grey = synth.make_grey(sp) if with_grey else None
    clip = orc.Clip(sp.W, sp.H, recs, off, flags, max_ref)
    out = {"recs": recs, "off": off, "flags": flags}
    if grey is not None:
        out["grey"] = grey
    flat = np.full((sp.H, sp.W), 128, np.uint8)
    prev = np.zeros(0, T.TRACK) if with_grey else synth.seed_tracks_lattice(sp)
    out["seed_tracks"] = prev
    cid = int(prev["track_id"].max()) if len(prev) else 0
    for f in range(sp.n_frames):
        out["hops_%d" % f] = clip.hops(f)
        out["kps_%d" % f] = clip.kps(f)
        out["cov_%d" % f] = np.float64(clip.coverage(f))
        out["grid_sha_%d" % f] = np.array(sha(clip.grid(f)))
        img = grey[f] if grey is not None else flat
        t, _, cid, _ = orc.extract_frame(sp.W, sp.H, flags[f], img, clip.grid(f), clip.hops(f), clip.kps(f), clip.coverage(f),
                                         prev, cid, max_tracks=max_tracks)
        out["tracks_%d" % f] = t
        prev = t
    return out


def pose_case():
    cam = T.camera(320.0, 320.0, 320.0, 240.0)
    pp = T.pose_params()
    pts, obs, pgt, pin = synth.pnp_problem(200, cam, 0x5EEDA003)
    n, pose, outl, stats = orc.pose_optimize(cam, pp, pts, obs, pin)
    return {"pose_pts": pts, "pose_obs": obs, "pose_init": pin, "pose_out": pose, "pose_outlier": outl,
            "pose_inliers": np.int32(n), "pose_stats": np.array(stats, np.int32)}


def main():
    blob = {}
    for name, (kw, max_ref, with_grey) in CASES.items():
        for k, v in run_case(kw, max_ref, with_grey).items():
            blob["%s/%s" % (name, k)] = v
    for k, v in pose_case().items():
        blob["pose/%s" % k] = v
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "frontend_small.npz")
    np.savez_compressed(path, **blob)
    print("wrote %s (%d arrays, %d bytes)" % (path, len(blob), os.path.getsize(path)))


if __name__ == "__main__":
    main()
