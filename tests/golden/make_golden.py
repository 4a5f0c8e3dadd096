#!/usr/bin/env python
"""Regenerates tests/golden/frontend_small.npz — outputs of THE REFERENCE'S OWN CODE on small seeded inputs.

The reference ships no golden vectors. Cases marked "ref" are produced by running the reference's
unmodified src/VideoDecoder.cc + src/MOVExtractor.cc + include/EXPRESS.h (oracle/_ref, built by `make -C oracle ref`
against stand-in headers; `canon` build = ties of the prev->mvVF sort ordered stably) in this container, and main()
refuses to write them unless the oracle reproduces every byte. The GPU box has no /root/reference: there the CUDA path is
checked against these committed arrays. The "oracle" case starts in mid-stream (its first records reference frames before
the clip, which under-runs the reference's decoder queue - undefined behaviour there), so it can only freeze the oracle's
documented window semantics; the pose case freezes the oracle's GN/Huber solver (parity unpinned at the OpenCV boundary).
Run from the repo root:  python tests/golden/make_golden.py
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
    "textured_ref4_long": (dict(width=176, height=112, n_frames=15, refs=4, seed=0x5EEDA004, fx=88.0, fy=88.0), 3, True),
    "flat_ref2": (dict(width=128, height=64, n_frames=5, refs=2, seed=0x5EEDA002, fx=64.0, fy=64.0, start_p=True), 1, False),
}
ENGINE = {"textured_ref3": "ref", "textured_ref4_long": "ref", "flat_ref2": "oracle"}   # who produced the committed arrays


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_case(kw, max_ref, with_grey, max_tracks=1024, engine="oracle"):
    sp = synth.Spec(**kw)
    recs, off, flags = synth.make_records(sp)
    grey = synth.make_grey(sp) if with_grey else None
    if engine == "ref":
        from oracle import pyref
        clip = pyref.Clip(sp.W, sp.H, recs, off, flags, grey=grey)
    else:
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
        # a frame without side data leaves VideoImage::coverageArea unassigned in the reference (VideoDecoder.cc:350)
        out["cov_%d" % f] = np.float64(clip.coverage(f) if flags[f] & T.FRAME_MV else 0.0)
        out["grid_sha_%d" % f] = np.array(sha(clip.grid(f)))
        img = grey[f] if grey is not None else flat
        if engine == "ref":
            r = pyref.extract_frame(sp.W, sp.H, flags[f], img, clip.grid(f), clip.hops(f), clip.kps(f), clip.coverage(f), prev, cid,
                                    has_prev=f > 0 or len(prev) > 0, variant="canon")
            assert r["consistent"] and len(r["tracks"]) <= max_tracks
            t, cid = r["tracks"], r["current_id"]
        else:
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
        got = run_case(kw, max_ref, with_grey, engine=ENGINE[name])
        if ENGINE[name] == "ref":   # the oracle must reproduce the reference's output byte for byte
            chk = run_case(kw, max_ref, with_grey, engine="oracle")
            for k in got:
                assert np.asarray(got[k]).tobytes() == np.asarray(chk[k]).tobytes(), (name, k)
        for k, v in got.items():
            blob["%s/%s" % (name, k)] = v
        print("%-20s produced by %s: %d tracks in the last table" % (name, ENGINE[name], len(got["tracks_%d" % (kw["n_frames"] - 1)])))
    for k, v in pose_case().items():
        blob["pose/%s" % k] = v
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "frontend_small.npz")
    np.savez_compressed(path, **blob)
    print("wrote %s (%d arrays, %d bytes)" % (path, len(blob), os.path.getsize(path)))


if __name__ == "__main__":
    main()
