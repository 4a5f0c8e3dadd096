"""Committed golden fixtures (tests/golden/frontend_small.npz, made by tests/golden/make_golden.py):
 - not gpu: the oracle still reproduces them (a change of the oracle's behaviour is caught);
 - gpu: the CUDA path reproduces them through the C-ABI without the oracle in the loop."""
import hashlib
import os
import sys

import numpy as np
import pytest

from movfe import types as T

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "frontend_small.npz"))


def _case(name):
    pre = name + "/"
    return {k[len(pre):]: GOLD[k] for k in GOLD.files if k.startswith(pre)}


@pytest.mark.parametrize("name", list(make_golden.CASES))
def test_oracle_reproduces_golden(orc, name):
    kw, max_ref, with_grey = make_golden.CASES[name]
    got = make_golden.run_case(kw, max_ref, with_grey)
    want = _case(name)
    assert set(got) == set(want)
    for k in want:
        assert np.asarray(got[k]).tobytes() == np.asarray(want[k]).tobytes(), (name, k)


def test_oracle_reproduces_golden_pose(orc):
    got, want = make_golden.pose_case(), _case("pose")
    for k in ("pose_outlier", "pose_inliers", "pose_stats"):
        assert np.asarray(got[k]).tobytes() == np.asarray(want[k]).tobytes(), k
    for f in ("R", "t"):
        assert np.allclose(got["pose_out"][f], want["pose_out"][f], rtol=0, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(make_golden.CASES))
def test_cuda_reproduces_golden(name):
    from movfe import lib
    kw, max_ref, with_grey = make_golden.CASES[name]
    g = _case(name)
    W, H, NF = kw["width"], kw["height"], kw["n_frames"]
    ctx = lib.Context(1, W, H, max_records_per_frame=1024, max_ref=max_ref, window_frames=NF, max_tracks=1024, has_grey=with_grey)
    if len(g["seed_tracks"]):
        ctx.set_tracks(0, g["seed_tracks"], int(g["seed_tracks"]["track_id"].max()))
    ctx.push_frames(NF, g["recs"], g["off"], g["flags"], g["grey"] if with_grey else None)
    ctx.raster(0, NF)
    ctx.extract(0, NF)
    for f in range(NF):
        assert ctx.hops(0, f).tobytes() == g["hops_%d" % f].tobytes(), (name, f, "hops")
        assert ctx.kps(0, f).tobytes() == g["kps_%d" % f].tobytes(), (name, f, "kps")
        assert ctx.raster_counts(0, f)[2] == float(g["cov_%d" % f]), (name, f, "coverage")
        assert hashlib.sha256(ctx.grid(0, f).tobytes()).hexdigest() == str(g["grid_sha_%d" % f]), (name, f, "grid")
        assert ctx.tracks(0, f).tobytes() == g["tracks_%d" % f].tobytes(), (name, f, "tracks")
    ctx.close()


@pytest.mark.gpu
def test_cuda_reproduces_golden_pose():
    from movfe import lib
    g = _case("pose")
    ctx = lib.Context(1, 640, 480, has_grey=False)
    cam = T.camera(320.0, 320.0, 320.0, 240.0)
    poses, outl, ninl, stats = ctx.pose_optimize(cam, T.pose_params(), g["pose_pts"], g["pose_obs"],
                                                 np.array([0, len(g["pose_pts"])], np.int32), np.array([g["pose_init"]]))
    assert int(ninl[0]) == int(g["pose_inliers"])
    assert outl.tobytes() == g["pose_outlier"].tobytes()
    # north_star tolerance: pose updates within 1e-5 relative of the reference restatement
    for f in ("R", "t"):
        ref = g["pose_out"][f]
        assert np.max(np.abs(poses[0][f] - ref)) <= 1e-5 * max(1.0, float(np.max(np.abs(ref))))
    ctx.close()
