"""The drop-in C++ shims (mov-slam_b200/shim): same class names and signatures as the reference's MOVExtractor,
MOVMatcher and Optimizer::PoseOptimization, plus the RasterQueue that stands where VideoDecoder::NextImage's MV loop
was. not gpu: they compile against the stand-in headers and link with libmovfe.so. gpu: a C++ driver runs them the way
Tracking.cc drives the reference classes and every output is compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

from movfe import synth, types as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "mov-slam_b200", "shim")


def _build():
    subprocess.run(["make", "-C", os.path.join(ROOT, "mov-slam_b200")], check=True, capture_output=True)
    r = subprocess.run(["make", "-C", SHIM], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_shims_compile_and_link():
    _build()
    assert os.path.exists(os.path.join(SHIM, "test_shim")) and os.path.exists(os.path.join(SHIM, "batched_frontend"))
    # the reference's signatures are what the shim headers declare
    h = open(os.path.join(SHIM, "MOVExtractor_movfe.h")).read()
    assert "MOVExtractor(int threshold = 20, double coverageThreshold = 0.60, double relocalizationDistance = 0.25)" in h
    assert "int operator()(const shared_ptr<MotionVectorImage> &_smv, std::vector<cv::KeyPoint> &_keypoints" in h
    m = open(os.path.join(SHIM, "MOVMatcher_movfe.h")).read()
    for sig in ("static int SearchByVideoFeature(Frame &F, const vector<MapPoint *> &vpMapPoints, const bool bFarPoints, const float thFarPoints)",
                "static int SearchByVideoFeature(KeyFrame *pKF, Frame &F, vector<MapPoint *> &vpMapPointMatches)",
                "static int SearchForInitialization(Frame &F1, Frame &F2, vector<cv::Point2f> &vbPrevMatched, vector<int> &vnMatches12, int windowSize)"):
        assert sig in m


def test_kannala_brandt8_camera_class_behind_the_reference_interface(orc):
    """shim/KannalaBrandt8_movfe.h overrides every pure virtual of GeometricCamera (GeometricCamera.h:61-101; the check
    program would not compile otherwise) and its projection / Jacobian equal the oracle's camera model, which the CUDA
    solver is tested against."""
    _build()
    cam = T.camera(190, 190, 376, 240, k=(-0.01, 0.002, -0.0005, 0.0001), model=T.CAM_FISHEYE)
    rng = np.random.Generator(np.random.PCG64(0x5A))
    for _ in range(20):
        X = rng.uniform([-3, -2, 0.5], [3, 2, 20])
        r = subprocess.run([os.path.join(SHIM, "camera_iface_check"), "190", "190", "376", "240", "-0.01", "0.002", "-0.0005", "0.0001"] +
                           ["%.17g" % v for v in X], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        v = [float(t) for t in r.stdout.split()]
        assert np.abs(np.array(v[:2]) - orc.project(cam, X)).max() < 1e-9
        assert np.abs(np.array(v[2:8]).reshape(2, 3) - orc.project_jac(cam, X)).max() < 1e-9
        assert abs(v[8] - X[0] / X[2]) < 1e-4 and abs(v[9] - X[1] / X[2]) < 1e-4 and int(v[10]) == 1     # unproject, CAM_FISHEYE


def test_shim_sources_compile_against_the_reference_headers():
    """-DMOVFE_IN_TREE: the shim sources include the reference's OWN Frame.h / MOVExtractor.h / VideoDecoder.h (through the
    oracle/_ref symlinks, third-party headers replaced by stand-ins) instead of standin/mov_slam_min.h."""
    ref = os.path.join(ROOT, "oracle", "_ref", "inc")
    if not os.path.isdir("/root/reference/include"):
        pytest.skip("/root/reference absent")
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "ref"], check=True)
    for src in ("MOVExtractor_movfe.cc", "VideoDecoder_movfe.cc", "shim_common.cc"):
        r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-w", "-DMOVFE_IN_TREE", "-I" + ref, "-I" + os.path.join(ROOT, "oracle", "ref_standin"),
                            "-I" + os.path.join(ROOT, "include"), "-I" + SHIM, os.path.join(SHIM, src)], capture_output=True, text=True)
        assert r.returncode == 0, (src, r.stderr[-3000:])


def pseudo_lk(pts_xy):
    """mirror of test_shim.cc::pseudo_lk, the deterministic stand-in for cv::calcOpticalFlowPyrLK"""
    pts = np.asarray(pts_xy, np.float32).reshape(-1, 2)
    out = (pts + np.array([0.5, -0.25], np.float32)).astype(np.float32)
    status = (((pts[:, 0].astype(np.int32) * 7 + pts[:, 1].astype(np.int32) * 3) % 5) != 0).astype(np.uint8)
    return status, out


def _stable_order(tr):
    pc = np.array([sum(bin(int(w)).count("1") for w in t["desc"]) for t in tr], np.int64)
    return sorted(range(len(tr)), key=lambda i: (-int(tr["age"][i]), -int(pc[i])))


@pytest.mark.gpu
@pytest.mark.parametrize("cov_thr,iframe_at", [(0.20, None), (0.95, 4)])
def test_shims_against_oracle(orc, tmp_path, cov_thr, iframe_at):
    """VideoDecoder shim (GPU raster behind the reference's class, fake libav in front) -> MOVExtractor shim (host LK hook,
    GPU merge) -> MOVMatcher / PoseOptimization shims, against the oracle frame by frame. The second case turns on the
    coverage back-fill (coverage tracks carried by LK from then on) and puts an intra picture in mid-stream (every track
    carried by LK): the drop-in keeps its track ids where round 1 silently lost them."""
    _build()
    W, H, NF, K, thr = 320, 240, 7, 2, 25
    sp = synth.Spec(W, H, n_frames=NF, refs=K + 1, seed=0x5EED0020, fx=160.0, fy=160.0)
    recs, off, flags = synth.make_records(sp)
    flags = flags.copy()
    if iframe_at is not None:
        flags[iframe_at] &= ~np.uint8(T.FRAME_P)
    grey = synth.make_grey(sp)
    clip = orc.Clip(W, H, recs, off, flags, 10)     # the shim's raster context accepts the reference's full ref range
    # oracle chain: tracks per frame, LK results from the same deterministic hook
    prev, cid, tracks, n_carried = np.zeros(0, T.TRACK), 0, [], 0
    for f in range(NF):
        lk = None
        if len(prev) and not (flags[f] & T.FRAME_P):
            lk = pseudo_lk(np.stack([prev["pt_x"], prev["pt_y"]], 1))
        elif len(prev):
            cov = [i for i in _stable_order(prev) if prev["flags"][i] & T.TRACK_COVERAGE]
            if cov:
                lk = pseudo_lk(np.stack([prev["pt_x"][cov], prev["pt_y"][cov]], 1))
        t, _, cid, _ = orc.extract_frame(W, H, flags[f], grey[f], clip.grid(f), clip.hops(f), clip.kps(f), clip.coverage(f), prev, cid,
                                         threshold=thr, coverage_threshold=cov_thr, max_tracks=8192,
                                         lk_status=None if lk is None else lk[0], lk_pts=None if lk is None else lk[1])
        n_carried += int(((t["q_indx"] >= 0) & (((t["flags"] & T.TRACK_COVERAGE) != 0) | (not (flags[f] & T.FRAME_P)))).sum())
        tracks.append(t)
        prev = t
    assert n_carried > 0 or iframe_at is None
    mp = synth.map_from_tracks(sp, tracks[0], synth.pose_at(sp, 0))
    mp["flags"][3] = T.MP_BAD
    mp["flags"][5] = T.MP_NULL
    n_kf = len(mp) - 7
    pose0 = synth.pose_struct(synth.pose_at(sp, 0))
    cam = sp.camera()
    d = str(tmp_path)
    np.array([W, H, NF, 10, thr, len(mp), n_kf, int(round(cov_thr * 1000))], np.int32).tofile(d + "/meta.bin")
    np.ascontiguousarray(recs, T.MV_RECORD).tofile(d + "/recs.bin")   # the 40-byte layout (numpy packs concatenated records)
    off.tofile(d + "/off.bin"); flags.tofile(d + "/flags.bin"); grey.tofile(d + "/grey.bin")
    mp.tofile(d + "/map.bin")
    np.concatenate([pose0["R"], pose0["t"]]).astype(np.float64).tofile(d + "/pose0.bin")
    np.array([cam["fx"], cam["fy"], cam["cx"], cam["cy"]], np.float32).tofile(d + "/cam.bin")
    r = subprocess.run([os.path.join(SHIM, "test_shim"), d], capture_output=True, text=True, )
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 carried tracks dropped" in r.stdout, r.stdout        # every carried track got its LK result
    # MOVMatcher::SearchByProjection (the shim's addition) on the last frame: its keypoints as map points projecting onto themselves
    import re
    msp = re.search(r"search_by_projection (\d+) matches, (\d+) identities of (\d+) keypoints", r.stdout)
    assert msp, r.stdout
    lt = tracks[NF - 1]
    sp_pts = np.zeros(len(lt), T.MAP_POINT)
    sp_proj = np.zeros(len(lt), T.PROJECTION)
    sp_proj["u"], sp_proj["v"], sp_proj["view_cos"], sp_proj["depth"], sp_proj["in_view"] = lt["pt_x"], lt["pt_y"], 0.9, 1.0, 1
    sp_prm = np.zeros(1, T.PROJECTION_SEARCH)
    sp_prm["th"], sp_prm["th_high"], sp_prm["nn_ratio"] = 1.0, 100, 1.0
    _, sp_pm, _, sp_n = orc.search_by_projection(lt, W, H, sp_pts, sp_proj, lt["desc"], sp_prm)
    assert [int(v) for v in msp.groups()] == [sp_n, int((sp_pm == np.arange(len(lt))).sum()), len(lt)], (msp.groups(), sp_n)
    assert sp_n > 0.9 * len(lt) > 0

    def f32(p):  # the Frame stores Sophus::SE3f: the pose is rounded to float between calls
        q = p.copy()
        q["R"] = q["R"].astype(np.float32)
        q["t"] = q["t"].astype(np.float32)
        return q

    pp = T.pose_params()
    last = f32(pose0)
    for f in range(NF):
        # the VideoDecoder shim: what VideoDecoder::NextImage's MV loop would have left in the VideoImage
        assert np.fromfile(d + "/out_hops_%d.bin" % f, T.HOP).tobytes() == clip.hops(f).tobytes(), ("hops", f)
        assert np.fromfile(d + "/out_kps_%d.bin" % f, T.RECT).tobytes() == clip.kps(f).tobytes(), ("kps", f)
        assert np.array_equal(np.fromfile(d + "/out_grid_%d.bin" % f, np.int32).reshape(H, W, 4), clip.grid(f)), ("grid", f)
        cov, frame_no, is_p = np.fromfile(d + "/out_cov_%d.bin" % f, np.float64)
        assert cov == clip.coverage(f) and int(is_p) == int(bool(flags[f] & T.FRAME_P)), ("coverage / type", f)
        got_t = np.fromfile(d + "/out_tracks_%d.bin" % f, T.TRACK)
        assert got_t.tobytes() == tracks[f].tobytes(), ("tracks", f, len(got_t), len(tracks[f]), r.stderr[-400:])
        # mDescriptors: the reference pushes none for back-fill features (MOVExtractor.cc:421 shadows the argument)
        n_backfill = int((((tracks[f]["flags"] & T.TRACK_COVERAGE) != 0) & (tracks[f]["q_indx"] < 0)).sum())
        assert int(np.fromfile(d + "/out_ndesc_%d.bin" % f, np.int32)[0]) == len(tracks[f]) - n_backfill, ("descriptors", f)
        n_m, want_m = orc.search_by_keyframe(tracks[f], mp[:n_kf])
        got_m = np.fromfile(d + "/out_match_%d.bin" % f, np.int32)
        assert np.array_equal(got_m, want_m), ("match", f)
        got_p = np.fromfile(d + "/out_pose_%d.bin" % f, np.float64)
        assert int(got_p[12]) == n_m
        sel = np.nonzero(want_m >= 0)[0]
        got_o = np.fromfile(d + "/out_outlier_%d.bin" % f, np.uint8)
        if len(sel) < 4:
            assert int(got_p[13]) == 0
            continue
        pts = mp["pos"][want_m[sel]]
        obs = np.stack([tracks[f]["pt_x"][sel], tracks[f]["pt_y"][sel]], 1)
        n, pose, outl, _ = orc.pose_optimize(cam, pp, pts, obs, last)
        assert int(got_p[13]) == n, ("inliers", f, got_p[13], n)
        want_o = np.ones(len(tracks[f]), np.uint8)
        want_o[sel] = outl
        assert np.array_equal(got_o, want_o), ("outlier", f)
        ref = np.concatenate([pose["R"], pose["t"]])
        assert np.max(np.abs(got_p[:12] - ref)) <= 1e-5 * max(1.0, float(np.max(np.abs(ref)))), ("pose", f)
        if n > 0:
            last = f32(pose)


@pytest.mark.gpu
def test_extractor_shim_with_gpu_lk(orc, tmp_path):
    """The MOVExtractor shim with its own LK provider (movfe_lk, the GPU tracker) at the reference's call sites: an intra picture in
    mid-stream carries every track by Lucas-Kanade (src/MOVExtractor.cc:81-120). Against the oracle chain fed with the numpy LK
    oracle's results: the same tracks survive, in the same order, with the same ids, ages, blocks and descriptors; carried
    positions within 5e-3 px (the tolerance of the GPU tracker against OpenCV)."""
    from oracle import lk as olk
    _build()
    W, H, NF, K, thr, iframe_at = 320, 240, 6, 2, 25, 4
    sp = synth.Spec(W, H, n_frames=NF, refs=K + 1, seed=0x5EED0021, fx=160.0, fy=160.0)
    recs, off, flags = synth.make_records(sp)
    flags = flags.copy()
    flags[iframe_at] &= ~np.uint8(T.FRAME_P)
    grey = synth.make_grey(sp)
    clip = orc.Clip(W, H, recs, off, flags, 10)
    prev, cid, tracks = np.zeros(0, T.TRACK), 0, []
    for f in range(iframe_at + 1):
        lk = None
        if f == iframe_at:
            out, st, _ = olk.track(grey[f - 1], grey[f], np.stack([prev["pt_x"], prev["pt_y"]], 1))
            lk = (st, out)
        t, _, cid, _ = orc.extract_frame(W, H, flags[f], grey[f], clip.grid(f), clip.hops(f), clip.kps(f), clip.coverage(f), prev, cid,
                                         threshold=thr, coverage_threshold=0.20, max_tracks=8192,
                                         lk_status=None if lk is None else lk[0], lk_pts=None if lk is None else lk[1])
        tracks.append(t)
        prev = t
    mp = synth.map_from_tracks(sp, tracks[0], synth.pose_at(sp, 0))
    pose0, cam = synth.pose_struct(synth.pose_at(sp, 0)), sp.camera()
    d = str(tmp_path)
    np.array([W, H, NF, 10, thr, len(mp), len(mp) // 2, 200], np.int32).tofile(d + "/meta.bin")
    np.ascontiguousarray(recs, T.MV_RECORD).tofile(d + "/recs.bin")
    off.tofile(d + "/off.bin"); flags.tofile(d + "/flags.bin"); grey.tofile(d + "/grey.bin"); mp.tofile(d + "/map.bin")
    np.concatenate([pose0["R"], pose0["t"]]).astype(np.float64).tofile(d + "/pose0.bin")
    np.array([cam["fx"], cam["fy"], cam["cx"], cam["cy"]], np.float32).tofile(d + "/cam.bin")
    r = subprocess.run([os.path.join(SHIM, "test_shim"), d], capture_output=True, text=True, env=dict(os.environ, SHIM_LK="gpu"))
    assert r.returncode == 0, r.stdout + r.stderr
    for f in range(iframe_at):       # before the intra picture nothing is carried: bit-exact
        assert np.fromfile(d + "/out_tracks_%d.bin" % f, T.TRACK).tobytes() == tracks[f].tobytes(), f
    got, want = np.fromfile(d + "/out_tracks_%d.bin" % iframe_at, T.TRACK), tracks[iframe_at]
    assert len(got) == len(want) and len(want) > 50
    for name in ("track_id", "age", "q_indx", "flags", "mb", "desc"):
        assert got[name].tobytes() == want[name].tobytes(), name
    assert np.max(np.abs(got["pt_x"] - want["pt_x"])) <= 5e-3 and np.max(np.abs(got["pt_y"] - want["pt_y"])) <= 5e-3


@pytest.mark.gpu
def test_batched_cpp_driver(orc, tmp_path):
    """The batched front-end driven from C++ through the C ABI alone (shim/batched_frontend.cc, the loop of INTEGRATION.md
    section 3, nothing but movfe_download_poses waits for the GPU): last window's track tables bit-exact, poses within 1e-5."""
    from gpu_util import oracle_tracks
    _build()
    W, H, F, K, NW, S, thr = 320, 240, 4, 2, 4, 5, 25
    NF = F * NW + K + 1
    sp = synth.Spec(W, H, n_frames=NF, refs=K + 1, seed=0x5EED0027, fx=160.0, fy=160.0)
    recs, off, flags = synth.make_records(sp)
    grey = synth.make_grey(sp)
    want = oracle_tracks(orc, (recs, off, flags), W, H, K, grey=grey, max_tracks=8192)
    mp = synth.map_from_tracks(sp, want[0], synth.pose_at(sp, 0))
    pose0, cam, pp = synth.pose_struct(synth.pose_at(sp, 0)), sp.camera(), T.pose_params()
    d = str(tmp_path)
    np.array([W, H, NF, K, thr, len(mp), len(mp) // 2, S, F], np.int32).tofile(d + "/meta.bin")
    np.ascontiguousarray(recs, T.MV_RECORD).tofile(d + "/recs.bin")
    off.tofile(d + "/off.bin"); flags.tofile(d + "/flags.bin"); grey.tofile(d + "/grey.bin"); mp.tofile(d + "/map.bin")
    np.concatenate([pose0["R"], pose0["t"]]).astype(np.float64).tofile(d + "/pose0.bin")
    np.array([cam["fx"], cam["fy"], cam["cx"], cam["cy"]], np.float32).tofile(d + "/cam.bin")
    r = subprocess.run([os.path.join(SHIM, "batched_frontend"), d], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rejected records 0" in r.stdout
    last = F * (NW - 1)
    for s in (0, S - 1):
        for f in range(last, last + F):
            got = np.fromfile(d + "/b_tracks_%d_%d.bin" % (s, f), T.TRACK)
            assert got.tobytes() == want[f].tobytes(), (s, f, len(got), len(want[f]))
    ref = orc.frontend_run(W, H, recs, off, flags, grey, None, mp, pose0, cam, pp, max_ref=K, max_tracks=8192, n_kf_points=len(mp) // 2)
    poses = np.fromfile(d + "/b_poses.bin", T.POSE).reshape(S, F)
    ninl = np.fromfile(d + "/b_ninl.bin", np.int32).reshape(S, F)
    for s in range(S):
        for k in range(F):
            assert ninl[s, k] == ref["n_inliers"][last + k], (s, k)
            for name in ("R", "t"):
                w = ref["poses"][last + k][name]
                assert np.max(np.abs(poses[s, k][name] - w)) <= 1e-5 * max(1.0, float(np.max(np.abs(w)))), (s, k, name)
