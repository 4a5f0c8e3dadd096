"""Generates tests/golden/lk_golden.npz: outputs of OpenCV's own cv2.calcOpticalFlowPyrLK (the build container has OpenCV 4.13 as a
Python package; the reference pins 4.6.0, Dockerfile:143) with the reference's parameters (src/MOVExtractor.cc:69,91-92) on seeded
image pairs - a smooth random texture, the same texture under a small affine motion; points in the interior, on the borders and in
flat regions (status 0). Also one pyramid level and the Scharr derivatives of the first image.   python tests/golden/make_lk_golden.py"""
import os

import cv2
import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lk_golden.npz")


def pair(seed, W, H, flat_box):
    rng = np.random.default_rng(seed)
    base = cv2.GaussianBlur((rng.random((H + 40, W + 40)) * 255).astype(np.uint8), (0, 0), 2.0)
    base = cv2.normalize(base, None, 0, 255, cv2.NORM_MINMAX)
    prev = base[20:20 + H, 20:20 + W].copy()
    x0, y0, x1, y1 = flat_box
    prev[y0:y1, x0:x1] = 90                                      # a flat patch: min eigenvalue below the threshold
    M = np.float32([[1.0 + 0.01 * rng.normal(), 0.01 * rng.normal(), 4 * rng.normal()], [0.01 * rng.normal(), 1.0 + 0.01 * rng.normal(), 3 * rng.normal()]])
    nxt = cv2.warpAffine(prev, M, (W, H), borderMode=cv2.BORDER_REFLECT_101)
    n = 80
    pts = np.stack([rng.uniform(1, W - 1, n), rng.uniform(1, H - 1, n)], 1).astype(np.float32)
    pts[:6] = [[0.5, 0.5], [W - 1.2, 3.0], [5.0, H - 1.0], [W - 0.6, H - 0.7], [16.0, 16.0], [(x0 + x1) / 2, (y0 + y1) / 2]]
    return prev, nxt, pts


def main():
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 20, 0.01)
    out = {"opencv_version": np.array(cv2.__version__)}
    for k, (seed, W, H, win, box) in enumerate([(11, 320, 240, 31, (100, 60, 180, 140)), (12, 213, 157, 31, (20, 20, 90, 90)), (13, 320, 240, 21, (200, 100, 260, 160))]):
        prev, nxt, pts = pair(seed, W, H, box)
        ref, st, er = cv2.calcOpticalFlowPyrLK(prev, nxt, pts, None, winSize=(win, win), maxLevel=3, criteria=crit, flags=cv2.OPTFLOW_LK_GET_MIN_EIGENVALS,
                                               minEigThreshold=1e-4)
        out.update({"prev%d" % k: prev, "next%d" % k: nxt, "pts%d" % k: pts, "out%d" % k: ref, "status%d" % k: st.ravel(), "err%d" % k: er.ravel(),
                    "win%d" % k: np.array(win)})
        if k == 0:
            out["pyr1"] = cv2.pyrDown(prev)
            out["dx"], out["dy"] = cv2.Scharr(prev, cv2.CV_16S, 1, 0), cv2.Scharr(prev, cv2.CV_16S, 0, 1)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: int(np.sum(out["status%d" % k])) for k in range(3)})


if __name__ == "__main__":
    main()
