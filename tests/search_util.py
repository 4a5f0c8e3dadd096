"""Shared by the search-by-projection tests: scene generation and a brute-force numpy restatement (no bucket grid) of the
operator include/movfe.h documents, written independently of oracle/match.cc."""
import numpy as np

from movfe import types as T


def popcount256(a, b):
    return int(np.unpackbits((a ^ b).view(np.uint8)).sum())


def make_frame(rng, W, H, n_feat, n_pts, noise=1.5, flips=12, clustered=False):
    """Keypoints with random descriptors; map points that project near some of them (descriptor = the keypoint's with a few
    bits flipped), near others with unrelated descriptors, and some far from any keypoint. Projections are synthetic."""
    feat = np.zeros(n_feat, T.TRACK)
    if clustered:
        feat["pt_x"] = np.clip(rng.normal(W / 2, 25, n_feat), -5, W + 5).astype(np.float32)
        feat["pt_y"] = np.clip(rng.normal(H / 2, 18, n_feat), -5, H + 5).astype(np.float32)
    else:
        feat["pt_x"] = rng.uniform(-4, W + 4, n_feat).astype(np.float32)
        feat["pt_y"] = rng.uniform(-4, H + 4, n_feat).astype(np.float32)
    feat["desc"] = rng.integers(0, 2 ** 32, (n_feat, 8), dtype=np.uint64).astype(np.uint32)
    pts = np.zeros(n_pts, T.MAP_POINT)
    proj = np.zeros(n_pts, T.PROJECTION)
    desc = rng.integers(0, 2 ** 32, (n_pts, 8), dtype=np.uint64).astype(np.uint32)
    for k in range(n_pts):
        kind = rng.integers(0, 10)
        if n_feat and kind < 7:                      # near a keypoint
            i = int(rng.integers(0, n_feat))
            proj["u"][k] = feat["pt_x"][i] + rng.normal(0, noise)
            proj["v"][k] = feat["pt_y"][i] + rng.normal(0, noise)
            if kind < 5:                             # with its descriptor, a few bits flipped
                d = feat["desc"][i].copy()
                for b in rng.integers(0, 256, int(rng.integers(0, flips + 1))):
                    d[b >> 5] ^= np.uint32(1 << (b & 31))
                desc[k] = d
        else:
            proj["u"][k] = rng.uniform(-10, W + 10)
            proj["v"][k] = rng.uniform(-10, H + 10)
        proj["view_cos"][k] = rng.choice([0.9, 0.998, 0.9981, 0.9999, 0.5])
        proj["depth"][k] = rng.uniform(0.5, 30.0)
        proj["in_view"][k] = rng.integers(0, 8) != 0
        pts["flags"][k] = rng.choice([0, 0, 0, 0, 0, 0, T.MP_BAD, T.MP_SKIP, T.MP_NULL])
    # duplicates: several map points with the SAME descriptor and projection compete for one keypoint
    for k in range(0, n_pts - 1, 9):
        desc[k + 1] = desc[k]
        proj[k + 1] = proj[k]
        pts["flags"][k + 1] = pts["flags"][k]
    return feat, pts, proj, desc


def brute_force(feat, W, H, pts, proj, desc, prm, taken=None):
    n, m = len(feat), len(pts)
    w_inv, h_inv = np.float32(64) / np.float32(W), np.float32(48) / np.float32(H)
    vx = (feat["pt_x"] * w_inv).astype(np.float32).astype(np.float64)
    vy = (feat["pt_y"] * h_inv).astype(np.float32).astype(np.float64)
    px = np.trunc(vx + np.copysign(0.5, vx)).astype(np.int64)     # C round(): half away from zero
    py = np.trunc(vy + np.copysign(0.5, vy)).astype(np.int64)
    in_grid = (px >= 0) & (px < 64) & (py >= 0) & (py < 48)
    order = np.lexsort((np.arange(n), py, px))                    # the reference's (ix, iy, insertion) walk
    order = order[in_grid[order]]
    prop = np.full(m, -1)
    dist = np.full(m, -1)
    th, far, th_far, th_high, ratio = prm["th"], prm["far_points"], prm["th_far"], prm["th_high"], prm["nn_ratio"]
    for k in range(m):
        if not proj["in_view"][k] or (far and proj["depth"][k] > th_far) or pts["flags"][k] & 7:
            continue
        r = np.float32(np.float32(2.5 if proj["view_cos"][k] > np.float32(0.998) else 4.0) * np.float32(th))
        dx = np.abs((feat["pt_x"][order] - proj["u"][k]).astype(np.float32))
        dy = np.abs((feat["pt_y"][order] - proj["v"][k]).astype(np.float32))
        cand = order[(dx < r) & (dy < r)]
        if taken is not None:
            cand = cand[taken[cand] == 0]
        if len(cand) == 0:
            continue
        d = np.array([popcount256(desc[k], feat["desc"][i]) for i in cand])
        b = int(np.argmin(d))                                     # first of the smallest
        if d[b] >= 256 or d[b] > th_high:
            continue
        rest = np.delete(d, b)
        if len(rest) and rest.min() < 256 and np.float32(d[b]) > np.float32(ratio) * np.float32(rest.min()):
            continue
        prop[k], dist[k] = cand[b], d[b]
    feat_match = np.full(n, -1)
    pt_match = np.full(m, -1)
    for f in np.unique(prop[prop >= 0]):
        ks = np.nonzero(prop == f)[0]
        w = ks[np.lexsort((ks, dist[ks]))[0]]                     # smallest (distance, point index)
        feat_match[f], pt_match[w] = w, f
    return feat_match, pt_match, dist
