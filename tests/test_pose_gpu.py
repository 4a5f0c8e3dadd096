"""GPU parity for frustum / joins / pose optimisation and the whole per-frame tracking chain, through the C-ABI.

Bars (BASELINE.json north_star): match indices bit-exact; poses within 1e-5 relative of the oracle."""
import numpy as np
import pytest

from movfe import lib, synth, types as T

from gpu_util import oracle_tracks, run_frontend_clip

pytestmark = pytest.mark.gpu
POSE_RTOL = 1e-5


def _rel(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / max(np.linalg.norm(b), 1e-12)


@pytest.fixture(scope="module")
def ctx():
    c = lib.Context(1, 640, 480, max_records_per_frame=64, max_ref=0, window_frames=1, max_tracks=64, max_map_points=16,
                    has_grey=False)
    yield c
    c.close()


def _random_map(rng, n, pose):
    mp = np.zeros(n, T.MAP_POINT)
    mp["pos"] = rng.uniform([-8, -6, 1], [8, 6, 25], (n, 3))
    Ow = -np.asarray(pose["R"]).reshape(3, 3).T @ np.asarray(pose["t"])
    d = mp["pos"].astype(np.float64) - Ow
    dist = np.linalg.norm(d, axis=1)
    nrm = d / dist[:, None] + rng.normal(0, 0.4, (n, 3))
    mp["normal"] = nrm / np.linalg.norm(nrm, axis=1)[:, None]
    mp["min_dist"], mp["max_dist"] = dist * rng.uniform(0.5, 1.3, n), dist * rng.uniform(0.8, 2.0, n)
    mp["track_id"] = rng.integers(1, 4000, n)
    mp["flags"] = rng.choice([0, 0, 0, 0, T.MP_BAD, T.MP_SKIP], n)
    return mp


@pytest.mark.parametrize("model", ["pinhole", "fisheye"])
def test_frustum_parity(orc, ctx, model):
    rng = np.random.Generator(np.random.PCG64(0x5EED0020))
    cam = T.camera(320, 320, 320, 240) if model == "pinhole" else \
        T.camera(190, 190, 376, 240, k=(-0.01, 0.002, -0.0005, 0.0001), model=T.CAM_FISHEYE)
    ctx.set_camera(cam, T.pose_params(), 0.5)
    poses = [synth.pose_struct(synth.pose_at(synth.Spec(phase=0.3 * i), 5 * i)) for i in range(3)]
    sets = [_random_map(rng, n, p) for n, p in zip((1000, 1, 20000), poses)]
    off = np.cumsum([0] + [len(s) for s in sets]).astype(np.int32)
    got = ctx.frustum(np.array(poses, T.POSE), np.concatenate(sets), off)
    for i, (mp, pose) in enumerate(zip(sets, poses)):
        want = orc.frustum(pose, cam, 640, 480, 0.5, mp)
        g = got[off[i]:off[i + 1]]
        assert g.tobytes() == want.tobytes(), (model, i, int((g["in_view"] != want["in_view"]).sum()))
        if len(mp) > 100:
            assert 0 < want["in_view"].sum() < len(mp)


def test_join_parity(orc, ctx):
    rng = np.random.Generator(np.random.PCG64(0x5EED0021))
    probs = []
    for n_t, n_p in ((4, 3), (3000, 1500), (1, 0), (0, 5), (4096, 20000)):
        tr = np.zeros(n_t, T.TRACK)
        tr["track_id"] = rng.integers(1, max(2, n_t // 2 + 2), n_t)           # duplicates: first index must win
        mp = np.zeros(n_p, T.MAP_POINT)
        mp["track_id"] = rng.integers(0, max(2, n_t // 2 + 50), n_p)          # duplicates: last probe must win
        proj = np.zeros(n_p, T.PROJECTION)
        proj["in_view"] = rng.random(n_p) < 0.7
        mp["flags"] = np.where(rng.random(n_p) < 0.1, T.MP_BAD, 0)
        init = np.where(rng.random(n_t) < 0.2, 7, -1).astype(np.int32)         # pre-existing matches stay unless hit
        probs.append((tr, mp, proj, init))
    probs[0][0]["track_id"][:] = [7, 3, 7, 9]                                  # KAT-8
    probs[0][1]["track_id"][:] = [9, 7, 7]
    probs[0][1]["flags"][:] = 0
    probs[0][2]["in_view"][:] = 1
    probs[0][3][:] = -1
    toff = np.cumsum([0] + [len(p[0]) for p in probs]).astype(np.int32)
    poff = np.cumsum([0] + [len(p[1]) for p in probs]).astype(np.int32)
    valid = np.concatenate([(p[2]["in_view"] != 0) & ((p[1]["flags"] & T.MP_BAD) == 0) for p in probs]).astype(np.uint8)
    match, n = ctx.join(np.concatenate([p[0]["track_id"] for p in probs]), toff,
                        np.concatenate([p[1]["track_id"] for p in probs]), valid, poff,
                        np.concatenate([p[3] for p in probs]))
    for i, (tr, mp, proj, init) in enumerate(probs):
        wn, wm = orc.search_by_video_feature(tr, mp, proj, init)
        assert n[i] == wn, (i, n[i], wn)
        assert np.array_equal(match[toff[i]:toff[i + 1]], wm), i
    assert list(match[:4]) == [2, -1, -1, 0] and n[0] == 3


@pytest.mark.parametrize("model", ["pinhole", "fisheye"])
def test_pose_optimize_parity(orc, ctx, model):
    cam = T.camera(320, 320, 320, 240) if model == "pinhole" else \
        T.camera(190, 190, 376, 240, k=(-0.01, 0.002, -0.0005, 0.0001), model=T.CAM_FISHEYE)
    pp = T.pose_params()
    sizes = (3, 4, 12, 500, 1000, 20000)
    probs = [synth.pnp_problem(n, cam, seed=0x5EED0030 + i, width=752 if model == "fisheye" else 640) for i, n in enumerate(sizes)]
    off = np.cumsum([0] + list(sizes)).astype(np.int32)
    poses, outl, ninl, stats = ctx.pose_optimize(cam, pp, np.concatenate([p[0] for p in probs]),
                                                 np.concatenate([p[1] for p in probs]), off, np.array([p[3] for p in probs], T.POSE))
    for i, (pts, obs, gt, init) in enumerate(probs):
        wn, wpose, woutl, wstats = orc.pose_optimize(cam, pp, pts, obs, init)
        assert ninl[i] == wn, (i, ninl[i], wn)
        assert np.array_equal(outl[off[i]:off[i + 1]], woutl), i
        assert _rel(poses[i]["R"], wpose["R"]) <= POSE_RTOL and _rel(poses[i]["t"], wpose["t"]) <= POSE_RTOL, i
        assert list(stats[i]) == list(wstats), (i, stats[i], wstats)
    assert ninl[0] == 0 and poses[0].tobytes() == probs[0][3].tobytes()      # < 4 points: untouched


def test_pose_optimize_lost_threshold(orc, ctx):
    cam = T.camera(458.654, 457.296, 367.215, 248.375)
    pp = T.pose_params(is_lost=True, iteration_count=20, reprojection_error=3.0, reprojection_error_lost=8.0)
    pts, obs, gt, init = synth.pnp_problem(800, cam, seed=0x5EED0040, sigma=1.5, width=752)
    poses, outl, ninl, _ = ctx.pose_optimize(cam, pp, pts, obs, np.array([0, 800], np.int32), np.array([init], T.POSE))
    wn, wpose, woutl, _ = orc.pose_optimize(cam, pp, pts, obs, init)
    assert ninl[0] == wn and np.array_equal(outl, woutl) and _rel(poses[0]["t"], wpose["t"]) <= POSE_RTOL


def _pipeline_case(orc, specs, window, max_ref, with_grey, seeds=None):
    W, H, NF = specs[0].W, specs[0].H, specs[0].n_frames
    streams = [synth.make_records(sp) for sp in specs]
    grey = [synth.make_grey(sp) for sp in specs] if with_grey else None
    cam = specs[0].camera()
    pp = T.pose_params()
    maps, pose0 = [], []
    for s, sp in enumerate(specs):
        t0 = oracle_tracks(orc, streams[s], W, H, max_ref, grey=None if grey is None else grey[s],
                           seeds=None if seeds is None else seeds[s])[0]
        maps.append(synth.map_from_tracks(sp, t0, synth.pose_at(sp, 0)))
        # start from a slightly wrong pose so the first optimisation has work to do
        R0, t0p = synth.pose_at(sp, 0)
        pose0.append(T.pose(R0, t0p + np.array([0.01, -0.02, 0.015])))

    def hook(ctx):
        ctx.set_camera(cam, pp, 0.5)
        for s in range(len(specs)):
            ctx.set_map_points(s, maps[s], len(maps[s]) // 2)
            ctx.set_pose(s, pose0[s])

    tracks, extra, ctx = run_frontend_clip(streams, W, H, NF, window, max_ref, grey=grey, seeds=seeds, poses=True,
                                           ctx_hook=hook, )
    worst = 0.0
    for s, sp in enumerate(specs):
        ref = orc.frontend_run(W, H, *streams[s], None if grey is None else grey[s], None if seeds is None else seeds[s],
                               maps[s], pose0[s], cam, pp, max_ref=max_ref, n_kf_points=len(maps[s]) // 2)
        for f in range(NF):
            P, n_inl = extra[(s, f)]
            assert len(tracks[(s, f)]) == ref["n_tracks"][f]
            assert n_inl == ref["n_inliers"][f], (s, f, n_inl, ref["n_inliers"][f])
            e = max(_rel(P["R"], ref["poses"][f]["R"]), _rel(P["t"], ref["poses"][f]["t"]))
            worst = max(worst, e)
            assert e <= POSE_RTOL, (s, f, e)
        gt = synth.pose_at(sp, NF - 1)
        assert np.linalg.norm(ref["poses"][NF - 1]["t"] - gt[1]) < 0.05       # the tracker really tracks
        assert ref["n_inliers"][NF - 1] > 50
    ctx.close()
    return worst


def test_pipeline_textured(orc):
    specs = [synth.Spec(640, 480, n_frames=10, refs=4, seed=0x5EED0050 + s, phase=0.5 * s) for s in range(2)]
    _pipeline_case(orc, specs, window=4, max_ref=3, with_grey=True)


def test_pipeline_mv_only_seeded(orc):
    specs = [synth.Spec(640, 480, n_frames=8, refs=2, seed=0x5EED0051, start_p=True)]
    seeds = [synth.seed_tracks_lattice(sp) for sp in specs]
    _pipeline_case(orc, specs, window=8, max_ref=1, with_grey=False, seeds=seeds)


def test_bucket_grid_parity(orc, ctx):
    """Frame::AssignFeaturesToGrid / GetFeaturesInArea (Frame.cc:356-388, 602-680): same CSR lists, same query results in
    the same order. Sets: empty, one point, clustered (many per cell), points on / outside the frame border, 9000 points."""
    rng = np.random.Generator(np.random.PCG64(0x5EED0031))
    sets = [np.zeros((0, 2), np.float32), np.array([[639.9, 479.9]], np.float32),
            rng.normal([320, 240], [30, 20], (700, 2)).astype(np.float32),
            np.concatenate([rng.uniform([-20, -20], [660, 500], (500, 2)), [[0, 0], [640, 480], [635.1, 475.1], [634.9, 474.9], [-0.4, -0.4]]]).astype(np.float32),
            rng.uniform([0, 0], [640, 480], (9000, 2)).astype(np.float32)]
    off = np.cumsum([0] + [len(s) for s in sets]).astype(np.int32)
    pts = np.concatenate(sets)
    start, items = ctx.assign_features_to_grid(pts, off)
    want = []
    for i, s in enumerate(sets):
        tr = np.zeros(len(s), T.TRACK)
        tr["pt_x"], tr["pt_y"] = s[:, 0], s[:, 1]
        ws, wi = orc.assign_features_to_grid(tr, 640, 480)
        want.append((tr, ws, wi))
        assert np.array_equal(start[i], ws), i
        nv = int(ws[-1])
        assert np.array_equal(items[off[i]:off[i] + nv], wi[:nv]), i
        assert (items[off[i] + nv:off[i + 1]] == -1).all()
    assert 0 < want[3][1][-1] < len(sets[3])          # some points of set 3 fall outside the grid
    q = []
    for i in (0, 1, 2, 3, 4):
        for _ in range(40):
            q.append((i, rng.uniform(-30, 670), rng.uniform(-30, 510), rng.choice([0.5, 5.0, 15.0, 60.0, 400.0])))
    q += [(4, 320.0, 240.0, 1000.0), (2, 320.0, 240.0, 0.0), (1, 639.0, 479.0, 2.0)]
    q = np.array(q, T.AREA_QUERY)
    cap = 9000
    out, counts = ctx.features_in_area(pts, off, start, items, q, cap)
    hit = 0
    for k, (i, x, y, r) in enumerate(q.tolist()):
        tr, ws, wi = want[i]
        w = orc.get_features_in_area(tr, 640, 480, ws, wi, np.float32(x), np.float32(y), np.float32(r))
        assert counts[k] == len(w), (k, counts[k], len(w))
        assert np.array_equal(out[k, :len(w)], w), k
        hit += len(w)
    assert hit > 1000
    # truncation: the full count is still reported
    out2, counts2 = ctx.features_in_area(pts, off, start, items, q[-3:], 10)
    assert counts2[0] == counts[-3] and np.array_equal(out2[0], out[-3, :10])


def test_keyframe_join_and_initialization_search_parity(orc, ctx):
    """MOVMatcher::SearchByVideoFeature(KeyFrame*, Frame&, out) (MOVMatcher.h:70-103: output starts all-NULL, null / bad
    entries of the keyframe's list are skipped) and SearchForInitialization (:105-137) as calls of movfe_join, the way the
    shim makes them."""
    rng = np.random.Generator(np.random.PCG64(0x5EED0023))
    tr = np.zeros(2500, T.TRACK)
    tr["track_id"] = rng.integers(1, 1800, len(tr))
    kf = np.zeros(900, T.MAP_POINT)
    kf["track_id"] = rng.integers(1, 2200, len(kf))
    kf["flags"] = rng.choice([0, 0, 0, T.MP_BAD, T.MP_NULL], len(kf))
    valid = ((kf["flags"] & (T.MP_BAD | T.MP_NULL)) == 0).astype(np.uint8)
    match, n = ctx.join(tr["track_id"], [0, len(tr)], kf["track_id"], valid, [0, len(kf)], np.full(len(tr), -1, np.int32))
    wn, wm = orc.search_by_keyframe(tr, kf)
    assert n[0] == wn and np.array_equal(match, wm) and 0 < wn < len(tr)
    # initialisation: F1's tracks probed with F2's ids -> vnMatches12[i1] = i2; vbPrevMatched from F2's keypoints
    f1, f2 = np.zeros(1200, T.TRACK), np.zeros(1500, T.TRACK)
    f1["track_id"], f2["track_id"] = rng.integers(1, 1400, len(f1)), rng.integers(1, 1400, len(f2))
    f2["pt_x"], f2["pt_y"] = rng.uniform(0, 640, len(f2)), rng.uniform(0, 480, len(f2))
    m12, n12 = ctx.join(f1["track_id"], [0, len(f1)], f2["track_id"], np.ones(len(f2), np.uint8), [0, len(f2)],
                        np.full(len(f1), -1, np.int32))
    prev = np.full((len(f1), 2), -9, np.float32)
    wn, wm12, wprev = orc.search_for_initialization(f1, f2, prev)
    assert n12[0] == wn and np.array_equal(m12, wm12)
    got_prev = prev.copy()
    hit = m12 >= 0
    got_prev[hit, 0], got_prev[hit, 1] = f2["pt_x"][m12[hit]], f2["pt_y"][m12[hit]]      # MOVMatcher.h:131-134, host side
    assert np.array_equal(got_prev, wprev)


def test_update_local_points_on_device(orc):
    """Tracking::UpdateLocalPoints (src/Tracking.cc:1171-1198) built on the device from index lists into a resident point store
    (movfe_set_map_store / movfe_update_local_points): the installed local maps equal the oracle's walk - NULL entries, bad points,
    duplicates across keyframes, a list longer than the local-map capacity, an empty list - and a pose chain run on them gives
    the poses of a context that received the same maps through movfe_set_map_points."""
    rng = np.random.default_rng(77)
    S, W, H, NF, K, CAP, NSTORE = 3, 320, 240, 5, 2, 60, 1500
    specs = [synth.Spec(W, H, n_frames=NF, refs=K + 1, seed=0x5EED0A10 + s, fx=160.0, fy=160.0) for s in range(S)]
    streams = [synth.make_records(sp) for sp in specs]
    greys = [synth.make_grey(sp) for sp in specs]
    want0 = [oracle_tracks(orc, streams[s], W, H, K, grey=greys[s])[0] for s in range(S)]
    stores, lists, n_first, maps = [], [], [], []
    for s in range(S):
        base = synth.map_from_tracks(specs[s], want0[s], synth.pose_at(specs[s], 0))   # real points first, padding after
        store = np.zeros(NSTORE, T.MAP_POINT)
        store[:len(base)] = base[:NSTORE]
        store["flags"][len(base):] = T.MP_BAD
        bad = rng.choice(len(base), len(base) // 10, replace=False)
        store["flags"][bad] |= T.MP_BAD
        n_list = [45, 0, 2500][s]                                  # stream 1: nothing; stream 2: more survivors than CAP
        idx = rng.integers(-1, len(base) + 20, n_list).astype(np.int32)    # -1 = NULL, >= len(base): culled padding
        stores.append(store)
        lists.append(idx)
        n_first.append(min(n_list, 20))
        maps.append(orc.update_local_points(store, idx, n_first[s], CAP))
    off = np.cumsum([0] + [len(l) for l in lists]).astype(np.int64)
    cam, pp = specs[0].camera(), T.pose_params()

    def run(install):
        ctx = lib.Context(S, W, H, max_records_per_frame=4800, max_ref=K, window_frames=NF, max_tracks=4096, max_map_points=CAP, has_grey=True)
        ctx.set_camera(cam, pp, 0.5)
        install(ctx)
        got = [ctx.map_points(s) for s in range(S)]
        for s in range(S):
            ctx.set_pose(s, synth.pose_struct(synth.pose_at(specs[s], 0)))
        from gpu_util import pack_streams
        r, o, fl = pack_streams(streams, NF, 0, NF)
        ctx.push_frames(NF, r, o, fl, np.stack([g[:NF] for g in greys]))
        ctx.raster(0, NF)
        ctx.extract(0, NF)
        ctx.track_poses(0, NF)
        P, ninl = ctx.poses(0, NF)
        ctx.close()
        return got, P, ninl

    def by_lists(ctx):
        ctx.reserve_map_store(NSTORE)
        for s in range(S):
            ctx.set_map_store(s, 0, stores[s][:700])
            ctx.set_map_store(s, 700, stores[s][700:])       # patched in two pieces
        ctx.update_local_points(np.concatenate(lists), off, n_first)

    def by_points(ctx):
        for s in range(S):
            ctx.set_map_points(s, maps[s][0], maps[s][1])

    got, P1, n1 = run(by_lists)
    for s in range(S):
        assert got[s][0].tobytes() == maps[s][0].tobytes(), (s, len(got[s][0]), len(maps[s][0]))
        assert got[s][1] == maps[s][1], (s, got[s][1], maps[s][1])
    assert len(maps[2][0]) == CAP and len(maps[1][0]) == 0 and 0 < len(maps[0][0]) < CAP
    _, P2, n2 = run(by_points)
    assert np.array_equal(n1, n2) and P1.tobytes() == P2.tobytes()
    assert n1[2].max() > 20      # the chain really ran on these maps
