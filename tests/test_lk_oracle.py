"""The pyramidal Lucas-Kanade oracle (oracle/lk.py, a numpy restatement of cv::calcOpticalFlowPyrLK as the reference calls it at
src/MOVExtractor.cc:91-92,196-197,347-348 and src/Frame.cc:305) against OpenCV's own outputs (tests/golden/lk_golden.npz, generated
by tests/golden/make_lk_golden.py with cv2 in the build container): pyramid level and Scharr derivatives bit-exact, status flags
identical, positions within 2e-3 px, min-eigenvalues within 1e-6."""
import os

import numpy as np

from oracle import lk

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lk_golden.npz"))


def test_pyramid_and_derivatives_bit_exact():
    assert np.array_equal(lk.pyr_down(G["prev0"]), G["pyr1"])
    dx, dy = lk.scharr(G["prev0"])
    assert np.array_equal(dx, G["dx"]) and np.array_equal(dy, G["dy"])
    odd = G["prev1"]                                            # odd width and height: (w+1)/2 x (h+1)/2
    assert lk.pyr_down(odd).shape == ((odd.shape[0] + 1) // 2, (odd.shape[1] + 1) // 2)


def test_tracks_equal_opencv():
    for k in range(3):
        out, st, err = lk.track(G["prev%d" % k], G["next%d" % k], G["pts%d" % k], win=int(G["win%d" % k]))
        assert np.array_equal(st, G["status%d" % k]), k
        assert 0 < int(st.sum()) < len(st)                      # tracked points and rejected ones (flat patch) both occur
        ok = st == 1
        assert np.max(np.linalg.norm(out[ok] - G["out%d" % k][ok], axis=1)) <= 2e-3, k
        assert np.max(np.abs(err - G["err%d" % k])) <= 1e-6, k
