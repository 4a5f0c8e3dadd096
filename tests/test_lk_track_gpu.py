"""GPU parity of the pyramidal Lucas-Kanade tracker (movfe_lk, csrc/lk.cu) - SURVEY.md 8f item 3: cv::calcOpticalFlowPyrLK as the
reference calls it (src/MOVExtractor.cc:91-92,196-197,347-348; src/Frame.cc:305). Against OpenCV's own outputs
(tests/golden/lk_golden.npz, made with cv2 by tests/golden/make_lk_golden.py): status flags identical, positions within 5e-3 px,
min-eigenvalues within 1e-6; against the numpy oracle (oracle/lk.py) on fresh image pairs: status identical, 5e-3 px. Tolerances
are written here because the window sums are formed in a different order than OpenCV's (float32, last bits)."""
import os

import numpy as np
import pytest

from movfe import lib

from oracle import lk as olk

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lk_golden.npz"))
POS_TOL = 5e-3


def test_against_opencv_golden():
    for k in range(3):
        prev, nxt, pts, win = G["prev%d" % k], G["next%d" % k], G["pts%d" % k], int(G["win%d" % k])
        H, W = prev.shape
        ctx = lib.Context(1, W, H, has_grey=False)
        out, st, err = ctx.lk(prev, nxt, pts, [0, len(pts)], win=win)
        ctx.close()
        assert np.array_equal(st, G["status%d" % k]), (k, np.nonzero(st != G["status%d" % k])[0])
        ok = st == 1
        assert np.max(np.linalg.norm(out[ok] - G["out%d" % k][ok], axis=1)) <= POS_TOL, k
        assert np.max(np.abs(err - G["err%d" % k])) <= 1e-6, k


def test_batched_pairs_against_oracle():
    """several image pairs and point lists of different length in one call (an I-frame carry-over hands in every track of a stream)"""
    rng = np.random.default_rng(21)
    W, H, NP = 256, 192, 3
    prevs, nxts, pts, off = [], [], [], [0]
    for p in range(NP):
        img = rng.random((H + 16, W + 16))
        for _ in range(3):          # smooth texture: box blur passes
            img = (img + np.roll(img, 1, 0) + np.roll(img, -1, 0) + np.roll(img, 1, 1) + np.roll(img, -1, 1)) / 5.0
        img = ((img - img.min()) / (img.max() - img.min()) * 255).astype(np.uint8)
        dx, dy = int(rng.integers(-4, 5)), int(rng.integers(-3, 4))
        prevs.append(img[8:8 + H, 8:8 + W])
        nxts.append(img[8 + dy:8 + dy + H, 8 + dx:8 + dx + W])
        n = [40, 1, 25][p]
        pts.append(np.stack([rng.uniform(0, W - 1, n), rng.uniform(0, H - 1, n)], 1).astype(np.float32))
        off.append(off[-1] + n)
    ctx = lib.Context(1, W, H, has_grey=False)
    out, st, err = ctx.lk(np.stack(prevs), np.stack(nxts), np.concatenate(pts), off)
    ctx.close()
    for p in range(NP):
        want, wst, werr = olk.track(prevs[p], nxts[p], pts[p])
        sl = slice(off[p], off[p + 1])
        assert np.array_equal(st[sl], wst), p
        ok = wst == 1
        if ok.any():
            assert np.max(np.linalg.norm(out[sl][ok] - want[ok], axis=1)) <= POS_TOL, p
        assert np.max(np.abs(err[sl] - werr)) <= 1e-6, p
    assert st.sum() > 40
