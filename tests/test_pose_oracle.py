"""CPU checks of the oracle's pose-only Gauss-Newton/Huber solver (the checker must itself be sane)."""
import numpy as np

from movfe import synth, types as T


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-12)


def test_recovers_ground_truth_on_clean_data(orc):
    cam = T.camera(320, 320, 320, 240)
    pts, obs, gt, init = synth.pnp_problem(500, cam, seed=1, sigma=0.0, outlier_frac=0.0)
    n, pose, outl, stats = orc.pose_optimize(cam, T.pose_params(), pts, obs, init)
    assert n == 500 and outl.sum() == 0
    assert _rel(pose["R"], gt["R"]) < 1e-6 and _rel(pose["t"], gt["t"]) < 1e-5
    assert stats[0] <= 48 and stats[3] == 0


def test_robust_to_gross_outliers(orc):
    cam = T.camera(320, 320, 320, 240)
    pts, obs, gt, init = synth.pnp_problem(1000, cam, seed=2, sigma=0.5, outlier_frac=0.10)
    n, pose, outl, _ = orc.pose_optimize(cam, T.pose_params(), pts, obs, init)
    assert 850 <= n <= 930
    assert _rel(pose["t"], gt["t"]) < 0.02 and _rel(pose["R"], gt["R"]) < 1e-3


def test_fisheye_model(orc):
    cam = T.camera(190, 190, 376, 240, k=(-0.01, 0.002, -0.0005, 0.0001), model=T.CAM_FISHEYE)
    pts, obs, gt, init = synth.pnp_problem(2000, cam, seed=3, sigma=0.0, outlier_frac=0.0)
    n, pose, _, _ = orc.pose_optimize(cam, T.pose_params(), pts, obs, init)
    assert n == 2000 and _rel(pose["t"], gt["t"]) < 1e-5
    # analytic Jacobian vs central differences
    X = np.array([0.4, -0.3, 1.7])
    J = orc.project_jac(cam, X)
    for k in range(3):
        d = np.zeros(3); d[k] = 1e-6
        num = (orc.project(cam, X + d) - orc.project(cam, X - d)) / 2e-6
        assert np.allclose(J[:, k], num, rtol=1e-6, atol=1e-6)


def test_fewer_than_four_points_returns_zero_and_keeps_pose(orc):
    cam = T.camera(320, 320, 320, 240)
    pts, obs, gt, init = synth.pnp_problem(3, cam, seed=4)
    n, pose, _, _ = orc.pose_optimize(cam, T.pose_params(), pts, obs, init)
    assert n == 0 and pose.tobytes() == init.tobytes()


def test_is_lost_uses_the_lost_threshold(orc):
    cam = T.camera(320, 320, 320, 240)
    pts, obs, gt, init = synth.pnp_problem(400, cam, seed=5, sigma=2.0, outlier_frac=0.0)
    n5, _, _, _ = orc.pose_optimize(cam, T.pose_params(reprojection_error=3.0, reprojection_error_lost=8.0), pts, obs, init)
    n8, _, _, _ = orc.pose_optimize(cam, T.pose_params(is_lost=True, reprojection_error=3.0, reprojection_error_lost=8.0), pts, obs, init)
    assert n8 > n5


def test_kat9_cross_check_against_opencv_on_clean_data(orc):
    """SURVEY KAT-9: on noise-free data the shipped solver (cv::solvePnPRansac, USAC_MAGSAC = 38) and the GN/Huber
    restatement must agree to ~1e-6. cv2 is the only piece of the reference's third-party arithmetic available here."""
    cv2 = __import__("pytest").importorskip("cv2")
    cam = T.camera(320, 320, 320, 240)
    pts, obs, gt, init = synth.pnp_problem(500, cam, seed=6, sigma=0.0, outlier_frac=0.0)
    K = np.array([[320, 0, 320], [0, 320, 240], [0, 0, 1]], np.float64)
    ok, rvec, tvec, inl = cv2.solvePnPRansac(pts.astype(np.float64), obs.astype(np.float64), K, np.zeros(4), iterationsCount=50,
                                             reprojectionError=5.0, confidence=0.95, flags=38)
    assert ok and len(inl) == 500
    Rcv, _ = cv2.Rodrigues(rvec)
    n, pose, _, _ = orc.pose_optimize(cam, T.pose_params(), pts, obs, init)
    assert np.abs(Rcv.ravel() - pose["R"]).max() < 1e-5 and np.abs(tvec.ravel() - pose["t"]).max() < 1e-4


def test_frustum_basic(orc):
    cam = T.camera(320, 320, 320, 240)
    mp = np.zeros(5, T.MAP_POINT)
    mp["pos"] = [[0, 0, 5], [0, 0, -5], [100, 0, 5], [0, 0, 5], [0, 0, 5]]
    mp["normal"] = [[0, 0, 1], [0, 0, 1], [0, 0, 1], [0, 0, -1], [0, 0, 1]]
    mp["min_dist"], mp["max_dist"] = 1, 10
    mp["flags"][4] = T.MP_BAD
    pr = orc.frustum(T.pose(), cam, 640, 480, 0.5, mp)
    assert list(pr["in_view"]) == [1, 0, 0, 0, 0]
    assert (pr["u"][0], pr["v"][0], pr["depth"][0], pr["view_cos"][0]) == (320, 240, 5, 1)
    assert (pr["u"][3], pr["v"][3]) == (320, 240) and (pr["u"][2], pr["v"][2]) == (-1, -1)


def test_bucket_grid(orc):
    rng = np.random.Generator(np.random.PCG64(7))
    tr = np.zeros(300, T.TRACK)
    tr["pt_x"], tr["pt_y"] = rng.uniform(0, 640, 300), rng.uniform(0, 480, 300)
    start, items = orc.assign_features_to_grid(tr, 640, 480)
    assert start[-1] <= 300 and len(set(items.tolist())) == len(items)
    got = orc.get_features_in_area(tr, 640, 480, start, items, 320.0, 240.0, 60.0)
    brute = [i for i in range(300) if abs(tr["pt_x"][i] - 320) < 60 and abs(tr["pt_y"][i] - 240) < 60 and i in set(items.tolist())]
    assert sorted(got.tolist()) == sorted(brute)
