"""GPU parity of the grid-bucketed search by projection (movfe_search_by_projection, north_star subsystem 3) through the C-ABI:
match indices and distances bit-exact against the oracle, whose results a brute-force restatement pins on the CPU
(tests/test_search_oracle.py). Projections come from movfe_frustum through both camera models."""
import numpy as np
import pytest

from movfe import lib, synth, types as T

from search_util import brute_force, make_frame

pytestmark = pytest.mark.gpu


def params(th=1.0, far=0, th_far=20.0, th_high=100, ratio=0.8):
    p = np.zeros(1, T.PROJECTION_SEARCH)
    p["th"], p["far_points"], p["th_far"], p["th_high"], p["nn_ratio"] = th, far, th_far, th_high, ratio
    return p


@pytest.fixture(scope="module")
def ctx():
    c = lib.Context(1, 640, 480, max_records_per_frame=64, max_ref=0, window_frames=1, max_tracks=64, max_map_points=16, has_grey=False)
    yield c
    c.close()


@pytest.mark.parametrize("prm", [params(), params(th=3.0, ratio=0.9), params(far=1, th_far=10.0, th_high=40, ratio=0.6)])
def test_synthetic_projections_batch(orc, ctx, prm):
    """Ragged batch: empty keypoint set, empty point list, clustered keypoints (long cells, many conflicts), 9000 keypoints."""
    rng = np.random.Generator(np.random.PCG64(0x5EED0041))
    frames = [make_frame(rng, 640, 480, nf, npt, clustered=cl) for nf, npt, cl in
              ((900, 400, False), (0, 40, False), (50, 0, False), (700, 900, True), (9000, 3000, False), (1, 1, False))]
    foff = np.cumsum([0] + [len(f[0]) for f in frames]).astype(np.int32)
    poff = np.cumsum([0] + [len(f[1]) for f in frames]).astype(np.int32)
    cat = lambda i: np.concatenate([f[i] for f in frames])
    for with_taken in (False, True):
        taken = (rng.integers(0, 4, foff[-1]) == 0).astype(np.uint8) if with_taken else None
        fm, pm, pd, nm = ctx.search_by_projection(cat(0), foff, cat(1), cat(2), cat(3), poff, prm, taken)
        total = 0
        for i, (feat, pts, proj, desc) in enumerate(frames):
            tk = None if taken is None else taken[foff[i]:foff[i + 1]]
            wfm, wpm, wpd, wn = orc.search_by_projection(feat, 640, 480, pts, proj, desc, prm, tk)
            assert np.array_equal(pd[poff[i]:poff[i + 1]], wpd), i
            assert np.array_equal(pm[poff[i]:poff[i + 1]], wpm), i
            assert np.array_equal(fm[foff[i]:foff[i + 1]], wfm), i
            assert nm[i] == wn, i
            total += wn
        assert total > 100
    # the brute-force restatement on one of the frames, straight against the GPU
    feat, pts, proj, desc = frames[3]
    g = ctx.search_by_projection(feat, [0, len(feat)], pts, proj, desc, [0, len(pts)], prm)
    w = brute_force(feat, 640, 480, pts, proj, desc, prm[0])
    assert np.array_equal(g[0], w[0]) and np.array_equal(g[1], w[1]) and np.array_equal(g[2], w[2])


@pytest.mark.parametrize("model", ["pinhole", "fisheye"])
def test_projected_map_points(orc, ctx, model):
    """Map points -> movfe_frustum (Pinhole / KannalaBrandt8) -> search: keypoints sit at the projections of a subset of the points
    (plus noise) and carry their descriptors with a few bits flipped; most of those must be found, bit-exact against the oracle."""
    rng = np.random.Generator(np.random.PCG64(0x5EED0042))
    cam = T.camera(320, 320, 320, 240) if model == "pinhole" else \
        T.camera(190, 190, 320, 240, k=(-0.01, 0.002, -0.0005, 0.0001), model=T.CAM_FISHEYE)
    ctx.set_camera(cam, T.pose_params(), 0.5)
    poses = [synth.pose_struct(synth.pose_at(synth.Spec(phase=0.2 * i), 4 * i)) for i in range(3)]
    frames = []
    for pose in poses:
        n = 2500
        mp = np.zeros(n, T.MAP_POINT)
        mp["pos"] = rng.uniform([-8, -6, 2], [8, 6, 25], (n, 3))
        Ow = -np.asarray(pose["R"]).reshape(3, 3).T @ np.asarray(pose["t"])
        d = mp["pos"].astype(np.float64) - Ow
        dist = np.linalg.norm(d, axis=1)
        mp["normal"] = (d / dist[:, None]).astype(np.float32)
        mp["min_dist"], mp["max_dist"] = dist * 0.5, dist * 2.0
        mp["flags"] = rng.choice([0, 0, 0, 0, 0, T.MP_BAD, T.MP_SKIP], n)
        pr = orc.frustum(pose, cam, 640, 480, 0.5, mp)
        desc = rng.integers(0, 2 ** 32, (n, 8), dtype=np.uint64).astype(np.uint32)
        seen = np.nonzero(pr["in_view"])[0]
        pick = seen[rng.random(len(seen)) < 0.7]
        feat = np.zeros(len(pick) + 300, T.TRACK)
        feat["pt_x"][:len(pick)] = pr["u"][pick] + rng.normal(0, 0.7, len(pick))
        feat["pt_y"][:len(pick)] = pr["v"][pick] + rng.normal(0, 0.7, len(pick))
        fd = desc[pick].copy()
        for row in fd:
            for b in rng.integers(0, 256, 10):
                row[b >> 5] ^= np.uint32(1 << (b & 31))
        feat["desc"][:len(pick)] = fd
        feat["pt_x"][len(pick):] = rng.uniform(0, 640, 300)
        feat["pt_y"][len(pick):] = rng.uniform(0, 480, 300)
        feat["desc"][len(pick):] = rng.integers(0, 2 ** 32, (300, 8), dtype=np.uint64).astype(np.uint32)
        perm = rng.permutation(len(feat))
        frames.append((feat[perm], mp, desc, pose, pick, np.argsort(perm)))
    foff = np.cumsum([0] + [len(f[0]) for f in frames]).astype(np.int32)
    poff = np.cumsum([0] + [len(f[1]) for f in frames]).astype(np.int32)
    pts = np.concatenate([f[1] for f in frames])
    proj = ctx.frustum(np.array(poses, T.POSE), pts, poff)
    prm = params(th=1.0, th_high=50, ratio=0.8)
    fm, pm, pd, nm = ctx.search_by_projection(np.concatenate([f[0] for f in frames]), foff, pts, proj, np.concatenate([f[2] for f in frames]), poff, prm)
    for i, (feat, mp, desc, pose, pick, inv) in enumerate(frames):
        pr = orc.frustum(pose, cam, 640, 480, 0.5, mp)
        assert pr.tobytes() == proj[poff[i]:poff[i + 1]].tobytes()
        wfm, wpm, wpd, wn = orc.search_by_projection(feat, 640, 480, mp, pr, desc, prm)
        assert np.array_equal(pm[poff[i]:poff[i + 1]], wpm) and np.array_equal(pd[poff[i]:poff[i + 1]], wpd)
        assert np.array_equal(fm[foff[i]:foff[i + 1]], wfm) and nm[i] == wn
        # the planted correspondences: keypoint inv[j] was made from map point pick[j]
        found = (wpm[pick] == inv[:len(pick)]).mean()
        assert found > 0.9, (model, i, found)


def test_properties_at_batch_scale(orc, ctx):
    """64 frames x 4000 keypoints x 1024 map points (the shape of one C2 window's last frames): properties that do not need the
    oracle on every frame - matches are mutual and unique, within the radius and the distance bound, independent of where a frame
    sits in the batch - and three frames against the oracle."""
    rng = np.random.Generator(np.random.PCG64(0x5EED0043))
    S = 64
    base = [make_frame(rng, 640, 480, 4000, 1024) for _ in range(4)]
    frames = [base[i % 4] for i in range(S)]
    foff = np.arange(S + 1, dtype=np.int32) * 4000
    poff = np.arange(S + 1, dtype=np.int32) * 1024
    cat = lambda i: np.concatenate([f[i] for f in frames])
    prm = params(th=1.0, th_high=60, ratio=0.8)
    feat, pts, proj, desc = cat(0), cat(1), cat(2), cat(3)
    fm, pm, pd, nm = ctx.search_by_projection(feat, foff, pts, proj, desc, poff, prm)
    assert nm.sum() > 64 * 150
    for i in range(S):
        f, p_, d = fm[foff[i]:foff[i + 1]], pm[poff[i]:poff[i + 1]], pd[poff[i]:poff[i + 1]]
        k = np.nonzero(p_ >= 0)[0]
        assert nm[i] == len(k) == (f >= 0).sum()
        assert np.array_equal(f[p_[k]], k)                          # mutual
        assert len(np.unique(p_[k])) == len(k)                      # a keypoint holds at most one map point
        assert (d[k] >= 0).all() and (d[k] <= 60).all() and (d[p_ < 0] <= 60).all()
        ft, pr = frames[i][0], frames[i][2]
        r = np.where(pr["view_cos"][k] > np.float32(0.998), 2.5, 4.0)
        assert (np.abs(ft["pt_x"][p_[k]] - pr["u"][k]) < r).all() and (np.abs(ft["pt_y"][p_[k]] - pr["v"][k]) < r).all()
        j = i % 4                                                   # the same frame elsewhere in the batch: same answer
        assert np.array_equal(p_, pm[poff[j]:poff[j + 1]]) and np.array_equal(f, fm[foff[j]:foff[j + 1]])
    for i in (0, 1, 2):
        w = orc.search_by_projection(frames[i][0], 640, 480, frames[i][1], frames[i][2], frames[i][3], prm)
        assert np.array_equal(fm[foff[i]:foff[i + 1]], w[0]) and np.array_equal(pm[poff[i]:poff[i + 1]], w[1])
        assert np.array_equal(pd[poff[i]:poff[i + 1]], w[2]) and nm[i] == w[3]


def test_argument_errors(ctx):
    feat = np.zeros(2, T.TRACK)
    pts, proj, desc = np.zeros(1, T.MAP_POINT), np.zeros(1, T.PROJECTION), np.zeros((1, 8), np.uint32)
    with pytest.raises(lib.MovfeError, match="offsets"):
        ctx.search_by_projection(feat, [0, 3, 2], pts, proj, desc, [0, 1, 1], params())
    big = np.zeros(16385, T.TRACK)
    with pytest.raises(lib.MovfeError, match="limit"):
        ctx.search_by_projection(big, [0, 16385], pts, proj, desc, [0, 1], params())
