"""BASELINE.json configs C3 / C4 / C5 at their full frame sizes (fewer streams than the headline counts, so the oracle
finishes in seconds): bit-exact hop lists, kps, slot grids and track tables, poses within 1e-5 relative."""
import numpy as np
import pytest

from movfe import lib, synth, types as T

from gpu_util import assert_raster_equal, assert_tracks_equal, oracle_tracks, run_frontend_clip, run_raster_clip

pytestmark = pytest.mark.gpu
POSE_RTOL = 1e-5     # north_star: pose updates within 1e-5 relative


def _rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(1.0, float(np.max(np.abs(b)))))


def test_c3_euroc_stereo_frame_packed(orc):
    """752x480, frame-packed stereo (odd frames are right views without MVs), ref=2."""
    W, H, NF = 752, 480, 10
    specs = [synth.Spec(W, H, n_frames=NF, refs=2, seed=0x5EED0300 + s, fx=458.654, fy=457.296, cx=367.215, cy=248.375,
                        stereo=True, phase=0.3 * s) for s in range(3)]
    streams = [synth.make_records(sp) for sp in specs]
    grey = [synth.make_grey(sp) for sp in specs]
    got, _, ctx = run_frontend_clip(streams, W, H, NF, 4, 1, grey=grey, max_records=5640)
    for s in range(len(specs)):
        want = oracle_tracks(orc, streams[s], W, H, 1, grey=grey[s])
        for f in range(NF):
            assert_tracks_equal(got[(s, f)], want[f], ("C3", s, f))
    ctx.close()


def test_c4_1080p_dense_4x4(orc):
    """1920x1080, one 4x4 record per cell (129 600 records / frame), MV-only mode with seeded tracks."""
    W, H, NF = 1920, 1080, 3
    M = (W // 4) * (H // 4)
    specs = [synth.Spec(W, H, n_frames=NF, refs=1, seed=0x5EED0400 + s, fx=960.0, fy=960.0, dense4x4=True, start_p=True, phase=0.2 * s)
             for s in range(2)]
    streams = [synth.make_records(sp) for sp in specs]
    assert all(len(r) == NF * M for r, _, _ in streams)
    seeds = [synth.seed_tracks_lattice(sp) for sp in specs]
    got, ctx = run_raster_clip(streams, W, H, NF, window=2, max_ref=0, max_records=M)
    for s in range(len(specs)):
        clip = orc.Clip(W, H, *streams[s], 0)
        for f in range(NF):
            assert_raster_equal(clip, got, s, f)
    ctx.close()
    tracks, _, ctx = run_frontend_clip(streams, W, H, NF, 2, 0, seeds=seeds, max_records=M, max_tracks=8192)
    for s in range(len(specs)):
        want = oracle_tracks(orc, streams[s], W, H, 0, seeds=seeds[s], max_tracks=8192)
        for f in range(NF):
            assert_tracks_equal(tracks[(s, f)], want[f], ("C4", s, f))
        assert len(want[NF - 1]) > 5000
    ctx.close()


@pytest.mark.parametrize("cluster", [None, "1", "4"])
def test_c5_pose_stress_fisheye(orc, cluster, monkeypatch):
    """20 000 correspondences per problem, KannalaBrandt8 (k from SURVEY.md 8d), sigma 0.5 px + 10 % gross outliers. Default: a
    cluster of two CTAs per problem adding its partial sums through distributed shared memory (pose_solve_cluster); MOVFE_POSE_CLUSTER=1
    one CTA per problem, =4 four."""
    if cluster:
        monkeypatch.setenv("MOVFE_POSE_CLUSTER", cluster)
    cam = T.camera(190.0, 190.0, 376.0, 240.0, k=(-0.01, 0.002, -0.0005, 0.0001), model=T.CAM_FISHEYE)
    pp = T.pose_params()
    ctx = lib.Context(1, 752, 480, has_grey=False)
    probs = [synth.pnp_problem(20000, cam, 0x5EED0500 + i, width=752, height=480) for i in range(6)]
    off = np.arange(len(probs) + 1, dtype=np.int32) * 20000
    poses, outl, ninl, stats = ctx.pose_optimize(cam, pp, np.concatenate([p[0] for p in probs]), np.concatenate([p[1] for p in probs]),
                                                 off, np.array([p[3] for p in probs]))
    for i, (pts, obs, gt, init) in enumerate(probs):
        n, pose, wout, wstats = orc.pose_optimize(cam, pp, pts, obs, init)
        assert int(ninl[i]) == n
        assert np.array_equal(outl[off[i]:off[i + 1]], wout)
        assert max(_rel(poses[i]["R"], pose["R"]), _rel(poses[i]["t"], pose["t"])) <= POSE_RTOL
        assert n > 17000
        assert np.linalg.norm(pose["t"] - gt["t"]) < 0.01
    ctx.close()
