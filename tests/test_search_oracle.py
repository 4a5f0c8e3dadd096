"""oracle/match.cc: orc_search_by_projection against a brute-force numpy restatement of the documented operator (no bucket
grid, candidates ordered by an explicit sort). The reference has no such function: parity unpinned,
the two restatements pin each other."""
import numpy as np
import pytest

from movfe import types as T
from oracle import pyoracle as orc
from search_util import brute_force, make_frame


def params(th=1.0, far=0, th_far=20.0, th_high=100, ratio=0.8):
    p = np.zeros(1, T.PROJECTION_SEARCH)
    p["th"], p["far_points"], p["th_far"], p["th_high"], p["nn_ratio"] = th, far, th_far, th_high, ratio
    return p


@pytest.mark.parametrize("W,H,n_feat,n_pts,clustered,prm", [
    (640, 480, 900, 400, False, params()),
    (640, 480, 700, 500, True, params(th=3.0, ratio=0.9)),          # many keypoints per cell, wide radius: conflicts
    (752, 480, 300, 300, False, params(th=1.0, far=1, th_far=10.0, th_high=40, ratio=0.6)),
    (100, 52, 60, 80, False, params(th=2.0)),
    (640, 480, 0, 50, False, params()), (640, 480, 50, 0, False, params()),
])
def test_oracle_equals_brute_force(W, H, n_feat, n_pts, clustered, prm):
    rng = np.random.Generator(np.random.PCG64(0x5EED0040 + n_feat + n_pts))
    feat, pts, proj, desc = make_frame(rng, W, H, n_feat, n_pts, clustered=clustered)
    for taken in (None, (rng.integers(0, 4, n_feat) == 0).astype(np.uint8)):
        fm, pm, pd, n = orc.search_by_projection(feat, W, H, pts, proj, desc, prm, taken)
        wfm, wpm, wpd = brute_force(feat, W, H, pts, proj, desc, prm[0], taken)
        assert np.array_equal(pd, wpd)
        assert np.array_equal(pm, wpm)
        assert np.array_equal(fm, wfm)
        assert n == int((wpm >= 0).sum())
        if n_feat >= 300 and n_pts >= 300:
            assert n > 20 and (wpd >= 0).sum() > n        # matches exist, and some proposals lost their keypoint


def test_single_candidate_skips_the_ratio_test():
    """One keypoint in range: bestLevel2 stays -1, the ratio test does not apply; a second one at distance 256 changes nothing."""
    feat = np.zeros(2, T.TRACK)
    feat["pt_x"], feat["pt_y"] = [100.0, 101.0], [100.0, 100.0]
    feat["desc"][0] = 0
    feat["desc"][1] = 0xffffffff
    pts = np.zeros(1, T.MAP_POINT)
    proj = np.zeros(1, T.PROJECTION)
    proj["u"], proj["v"], proj["view_cos"], proj["in_view"] = 100.5, 100.0, 0.9, 1
    desc = np.zeros((1, 8), np.uint32)
    desc[0, 0] = 0x7          # distance 3 to keypoint 0, 253 to keypoint 1
    fm, pm, pd, n = orc.search_by_projection(feat, 640, 480, pts, proj, desc, params(ratio=0.001))
    assert n == 0 and pd[0] == -1                      # 3 > 0.001 * 253: rejected by the ratio test
    desc[0, 0] = 0            # distances 1 / 256: `dist < bestDist2` (256) is false, the second candidate never registers
    feat["desc"][0, 0] = 1
    fm, pm, pd, n = orc.search_by_projection(feat, 640, 480, pts, proj, desc, params(ratio=0.001))
    assert n == 1 and pm[0] == 0 and pd[0] == 1 and fm[0] == 0
