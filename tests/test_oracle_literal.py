"""The C++ oracle against a second, independently written restatement of the reference (oracle/literal.py: plain Python,
one float32 rounding per operation, literal containers) on random inputs. The reference ships no test vectors and cannot
be built here, so agreement of two separate transcriptions is what pins the oracle beyond the hand-derived KATs."""
import numpy as np
import pytest

from movfe import synth, types as T
from oracle import literal as lit


def _records(rng, W, H, n, max_ref, sizes, frame_index):
    r = np.zeros(n, T.MV_RECORD)
    r["source"] = rng.choice([-1, -1, -1, -1, 0, 1], n)
    r["w"], r["h"] = rng.choice(sizes, n), rng.choice(sizes, n)
    r["dst_x"], r["dst_y"] = rng.integers(-6, W + 6, n), rng.integers(-6, H + 6, n)
    r["src_x"] = r["dst_x"] + rng.integers(-40, 41, n)
    r["src_y"] = r["dst_y"] + rng.integers(-40, 41, n)
    r["ref"] = rng.integers(0, min(max_ref, max(frame_index - 1, 0)) + 1, n)
    return r


@pytest.mark.parametrize("seed,W,H,max_ref", [(1, 64, 48, 3), (2, 97, 61, 2), (3, 40, 40, 0), (4, 128, 72, 5)])
def test_raster_matches_literal_transcription(orc, seed, W, H, max_ref):
    rng = np.random.Generator(np.random.PCG64(0x11E0 + seed))
    NF = 8
    frames, flags = [], []
    for f in range(NF):
        mv_on = f > 0 and rng.random() > 0.15
        n = int(rng.integers(0, 70)) if f > 0 else 0
        frames.append(_records(rng, W, H, n, max_ref, [4, 8, 16], f))
        flags.append((T.FRAME_P if f > 0 else 0) | (T.FRAME_MV if mv_on else 0))
    off = np.cumsum([0] + [len(fr) for fr in frames]).astype(np.int64)
    recs = np.concatenate(frames) if off[-1] else np.zeros(0, T.MV_RECORD)
    clip = orc.Clip(W, H, recs, off, np.array(flags, np.uint8), max_ref)
    vqueue = []
    for f in range(NF):
        rl = [dict(source=int(r["source"]), w=int(r["w"]), h=int(r["h"]), src_x=int(r["src_x"]), src_y=int(r["src_y"]),
                   dst_x=int(r["dst_x"]), dst_y=int(r["dst_y"]), ref=int(r["ref"])) for r in frames[f]]
        lit.next_image_mv_loop(W, H, vqueue, rl, mv_enabled=bool(flags[f] & T.FRAME_MV))
    for f in range(NF):
        want = vqueue[f]
        hops, kps = clip.hops(f), clip.kps(f)
        assert len(hops) == len(want.mvs) and len(kps) == len(want.kps), (f, len(hops), len(want.mvs), len(kps), len(want.kps))
        for i, (mx, my, d) in enumerate(want.mvs):
            assert (hops["mv_x"][i], hops["mv_y"][i], hops["d_indx"][i]) == (mx, my, d), (f, i)
        for i, k in enumerate(want.kps):
            assert (kps["x"][i], kps["y"][i], kps["w"][i], kps["h"][i]) == k, (f, i)
        assert np.array_equal(clip.grid(f), want.mvi), f
        assert clip.coverage(f) == want.coverageArea, (f, clip.coverage(f), want.coverageArea)
    assert sum(len(v.mvs) for v in vqueue) > 100


@pytest.mark.parametrize("shape", [(16, 16), (8, 8), (16, 8), (8, 16)])
def test_express_matches_literal_transcription(orc, shape):
    rows, cols = shape
    rng = np.random.Generator(np.random.PCG64(0x11F0 + rows * 3 + cols))
    spec = synth.Spec(160, 120, n_frames=2, refs=1, seed=0x5EED0061)
    textured = synth.make_grey(spec)[1]
    noisy = rng.integers(0, 256, (120, 160)).astype(np.uint8)
    steps = (np.add.outer(np.arange(120) // 9, np.arange(160) // 7) % 2 * 180 + 20).astype(np.uint8)
    smooth = (np.add.outer(np.arange(120), np.arange(160)) % 256).astype(np.uint8)   # exercises the uint8 wrap of the band
    n_true = 0
    for img in (textured, noisy, steps, smooth):
        for _ in range(60):
            x0, y0 = int(rng.integers(0, 160 - cols - 1)), int(rng.integers(0, 120 - rows))
            thr = int(rng.choice([5, 25, 60, 140]))
            m = lit.Roi(img, x0, y0, cols, rows)
            assert orc.express_center(img, x0, y0, cols, rows) == lit.compute_center(m)
            d = orc.express_descriptor(img, x0, y0, cols, rows, thr)
            want = lit.compute_descriptor(m, thr)
            got = sum(int(w) << (32 * i) for i, w in enumerate(d))
            assert got == want, (x0, y0, thr)
            e = bool(orc.express_test(img, x0, y0, cols, rows, thr))
            assert e == lit.compute_express(m, thr), (x0, y0, thr)
            n_true += e
    assert 0 < n_true < 240


def _tracks_of(vfs):
    t = np.zeros(len(vfs), T.TRACK)
    for i, v in enumerate(vfs):
        t[i]["pt_x"], t[i]["pt_y"] = v.pt
        t[i]["mb"] = v.mb
        t[i]["track_id"], t[i]["age"], t[i]["q_indx"] = v.trackId, v.age, v.qIndx
        t[i]["flags"] = T.TRACK_COVERAGE if v.coverage else 0
        t[i]["desc"] = [(v.desc >> (32 * k)) & 0xffffffff for k in range(8)]
    return t


@pytest.mark.parametrize("seed,cov_thr", [(0, 0.20), (1, 0.95)])
def test_extractor_matches_literal_transcription(orc, seed, cov_thr):
    """MOVExtractor::operator() frame after frame (I-frame seeding, propagation with candidate choice, claims, descriptor
    gate, births, lattice back-fill), both restatements fed their own previous table."""
    W, H, NF, K = 160, 112, 6, 2
    spec = synth.Spec(W, H, n_frames=NF, refs=K + 1, seed=0x5EED0070 + seed, fx=80.0, fy=80.0)
    recs, off, flags = synth.make_records(spec)
    grey = synth.make_grey(spec)
    clip = orc.Clip(W, H, recs, off, flags, K)
    vqueue = []
    for f in range(NF):
        rl = [dict(source=int(r["source"]), w=int(r["w"]), h=int(r["h"]), src_x=int(r["src_x"]), src_y=int(r["src_y"]),
                   dst_x=int(r["dst_x"]), dst_y=int(r["dst_y"]), ref=int(r["ref"])) for r in recs[off[f]:off[f + 1]]]
        lit.next_image_mv_loop(W, H, vqueue, rl, mv_enabled=bool(flags[f] & T.FRAME_MV))
    prev_o, cid_o = np.zeros(0, T.TRACK), 0
    prev_l, cid_l = [], 0
    seen_multi = births = 0
    for f in range(NF):
        got, _, cid_o, _ = orc.extract_frame(W, H, flags[f], grey[f], clip.grid(f), clip.hops(f), clip.kps(f), clip.coverage(f),
                                             prev_o, cid_o, threshold=25, coverage_threshold=cov_thr, max_tracks=4096)
        want_vf, cid_l = lit.extractor(vqueue[f], grey[f], bool(flags[f] & T.FRAME_P), prev_l, cid_l, 25, cov_thr)
        want = _tracks_of(want_vf)
        assert cid_o == cid_l, (f, cid_o, cid_l)
        assert len(got) == len(want), (f, len(got), len(want))
        for name in got.dtype.names:
            assert got[name].tobytes() == want[name].tobytes(), (f, name)
        births += int((got["q_indx"] < 0).sum()) if f > 0 else 0
        prev_o, prev_l = got, want_vf
        if f > 0:
            g = clip.grid(f).reshape(H, W, 4)
            seen_multi += int(sum(g[int(v.pt[1]), int(v.pt[0]), 1] >= 0 for v in want_vf if v.qIndx >= 0))
    assert len(prev_o) > 20 and births > 0 and seen_multi >= 0


@pytest.mark.parametrize("model", ["pinhole", "fisheye"])
def test_pose_building_blocks_against_independent_numerics(orc, model):
    """SE3 exponential against scipy's matrix exponential of the twist; the analytic 2x6 residual Jacobian against central
    differences of the projection; the projection itself against a separate restatement. No formula is shared."""
    cam = T.camera(320, 320, 320, 240) if model == "pinhole" else \
        T.camera(190, 190, 376, 240, k=(-0.01, 0.002, -0.0005, 0.0001), model=T.CAM_FISHEYE)
    rng = np.random.Generator(np.random.PCG64(0x11A0))
    for _ in range(30):
        dx = rng.normal(0, [0.3, 0.3, 0.3, 1, 1, 1])
        R, t = orc.se3_exp(dx)
        Rw, tw = lit.se3_exp_expm(dx)
        assert np.abs(R - Rw).max() < 1e-12 and np.abs(t - tw).max() < 1e-12
        Xc = rng.uniform([-3, -2, 1], [3, 2, 20])
        assert np.abs(orc.project(cam, Xc) - lit._project(cam, Xc)).max() < 1e-9
        J, Jfd = orc.pose_jacobian(cam, Xc), lit.residual_jacobian_fd(cam, Xc)
        assert np.abs(J - Jfd).max() <= 2e-6 * max(1.0, np.abs(Jfd).max()), (Xc, J, Jfd)
    R0, t0 = orc.se3_exp(np.array([1e-9, 0, 0, 0.5, 0, 0]))            # small-angle branch
    Rw, tw = lit.se3_exp_expm(np.array([1e-9, 0, 0, 0.5, 0, 0]))
    assert np.abs(R0 - Rw).max() < 1e-14 and np.abs(t0 - tw).max() < 1e-14


@pytest.mark.parametrize("model,n,seed", [("pinhole", 60, 1), ("pinhole", 300, 2), ("fisheye", 200, 3)])
def test_pose_solver_matches_second_restatement(orc, model, n, seed):
    """Optimizer::PoseOptimization's Gauss-Newton / Huber schedule, noisy observations with 10 % gross outliers: the C++
    oracle (analytic Jacobian, Cholesky, closed-form exponential) and the numpy restatement (LU solve, scipy expm) - run
    once with the oracle's Jacobian and once with finite differences - classify every point alike and agree on the pose."""
    cam = T.camera(320, 320, 320, 240) if model == "pinhole" else \
        T.camera(190, 190, 376, 240, k=(-0.01, 0.002, -0.0005, 0.0001), model=T.CAM_FISHEYE)
    pts, obs, gt, init = synth.pnp_problem(n, cam, seed=0x5EED0090 + seed, width=752 if model == "fisheye" else 640)
    pp = T.pose_params()
    wn, wpose, woutl, _ = orc.pose_optimize(cam, pp, pts, obs, init)
    R0, t0 = np.asarray(init["R"]).reshape(3, 3), np.asarray(init["t"])
    for jac, tol in ((lambda c, X: orc.pose_jacobian(c, X), 1e-9), (lit.residual_jacobian_fd, 1e-6)):
        R, t, outl, ninl = lit.pose_optimize_ref(cam, pts, obs, R0.copy(), t0.copy(), float(pp["reprojection_error"]),
                                                 int(pp["iteration_count"]), jac)
        assert ninl == wn and np.array_equal(outl.astype(np.uint8), woutl), (ninl, wn)
        assert np.abs(R - np.asarray(wpose["R"]).reshape(3, 3)).max() < tol and np.abs(t - np.asarray(wpose["t"])).max() < tol * 10
    assert 0.7 * n < wn < n                                               # the gross outliers were rejected, the rest kept


def test_frustum_joins_and_bucket_grid_match_literal(orc):
    rng = np.random.Generator(np.random.PCG64(0x11B0))
    spec = synth.Spec(640, 480, n_frames=2, refs=1, seed=1)
    Rcw, tcw = synth.pose_at(spec, 7)
    pose, cam = synth.pose_struct((Rcw, tcw)), T.camera(320, 320, 320, 240)
    n = 400
    mp = np.zeros(n, T.MAP_POINT)
    mp["pos"] = rng.uniform([-8, -6, -2], [8, 6, 25], (n, 3))
    Ow = -np.asarray(Rcw).T @ np.asarray(tcw)
    d = mp["pos"].astype(np.float64) - Ow
    dist = np.linalg.norm(d, axis=1)
    nrm = d / dist[:, None] + rng.normal(0, 0.5, (n, 3))
    mp["normal"] = nrm / np.linalg.norm(nrm, axis=1)[:, None]
    mp["min_dist"], mp["max_dist"] = dist * rng.uniform(0.5, 1.3, n), dist * rng.uniform(0.8, 2.0, n)
    got = orc.frustum(pose, cam, 640, 480, 0.5, mp)
    for k in range(n):
        w = lit.is_in_frustum(np.asarray(Rcw), np.asarray(tcw), cam, 640, 480, 0.5, mp["pos"][k], mp["normal"][k],
                              mp["min_dist"][k], mp["max_dist"][k])
        g = got[k]
        assert (int(g["in_view"]), g["u"], g["v"], g["depth"], g["view_cos"]) == tuple(w), (k, g, w)
    assert 20 < got["in_view"].sum() < n
    # joins
    tr = np.zeros(300, T.TRACK)
    tr["track_id"] = rng.integers(1, 200, len(tr))
    mp["track_id"] = rng.integers(1, 260, n)
    proj = got
    init = np.where(rng.random(len(tr)) < 0.2, 7, -1).astype(np.int32)
    wn, wm = orc.search_by_video_feature(tr, mp, proj, init)
    m = init.copy()
    assert lit.search_by_video_feature(tr["track_id"], mp["track_id"], proj["in_view"] != 0, m) == wn and np.array_equal(m, wm)
    # bucket grid
    tr["pt_x"], tr["pt_y"] = rng.uniform(-10, 650, len(tr)), rng.uniform(-10, 490, len(tr))
    start, items = orc.assign_features_to_grid(tr, 640, 480)
    want = lit.assign_features_to_grid(tr["pt_x"], tr["pt_y"], 640, 480)
    for ix in range(64):
        for iy in range(48):
            c = ix * 48 + iy
            assert list(items[start[c]:start[c + 1]]) == want.get((ix, iy), []), (ix, iy)
