"""oracle/lk.py — TEST INFRASTRUCTURE. numpy restatement of the sparse pyramidal Lucas-Kanade tracker the reference calls at
src/MOVExtractor.cc:91-92,196-197,347-348 (winSize 31x31, maxLevel 3, 20 iterations / eps 0.01, OPTFLOW_LK_GET_MIN_EIGENVALS,
minEigThreshold 1e-4) and src/Frame.cc:305 (21x21): cv::calcOpticalFlowPyrLK.

The algorithm lives in a THIRD-PARTY dependency that is absent from /root/reference: OpenCV 4.6.0 (Dockerfile:143),
modules/video/src/lkpyramid.cpp (cv::calcOpticalFlowPyrLK -> cv::buildOpticalFlowPyramid, calcSharrDeriv,
cv::detail::LKTrackerInvoker::operator()) and modules/imgproc/src/pyramids.cpp (cv::pyrDown). It is restated here from the
published algorithm: the 5-tap [1 4 6 4 1] pyramid with (sum + 128) >> 8 rounding, Scharr derivatives as int16, BORDER_REFLECT_101
image borders and zero derivative borders of winSize pixels, 14-bit fixed-point bilinear weights, the float normal equations
scaled by 2^-20, the min-eigenvalue test, the coarse-to-fine loop with the half-step oscillation stop.
PINNED against the real thing: tests/golden/lk_golden.npz holds outputs of cv2.calcOpticalFlowPyrLK (OpenCV 4.13, run in the
build container by tests/golden/make_lk_golden.py) on seeded image pairs; tests/test_lk_oracle.py requires identical status
flags, positions within 2e-3 px and min-eigenvalues within 1e-6 (OpenCV's SIMD path sums the window in a different order, so
the last float bits differ), pyramid levels and derivative images bit-exact. Only tests/ may import this module."""
import numpy as np

W_BITS = 14
FLT_SCALE = np.float32(1.0 / (1 << 20))


def _reflect101(idx, n):
    """BORDER_REFLECT_101 source index for every entry of idx (any integers) into an axis of length n."""
    idx = np.asarray(idx, np.int64)
    if n == 1:
        return np.zeros_like(idx)
    period = 2 * (n - 1)
    m = np.mod(idx, period)
    return np.where(m >= n, period - m, m)


def pyr_down(img):
    """cv::pyrDown for uint8: separable [1 4 6 4 1], BORDER_REFLECT_101, exact integer sum, (sum + 128) >> 8; (w+1)/2 x (h+1)/2."""
    h, w = img.shape
    oh, ow = (h + 1) // 2, (w + 1) // 2
    k = np.array([1, 4, 6, 4, 1], np.int32)
    xs = _reflect101(2 * np.arange(ow)[:, None] + np.arange(5)[None, :] - 2, w)
    ys = _reflect101(2 * np.arange(oh)[:, None] + np.arange(5)[None, :] - 2, h)
    a = img.astype(np.int32)
    rows = (a[:, xs] * k).sum(-1)
    out = (rows[ys, :] * k[None, :, None]).sum(1)
    return ((out + 128) >> 8).astype(np.uint8)


def scharr(img):
    """calcSharrDeriv: (dx, dy) as int16, 3/10/3 smoothing across, central difference along, BORDER_REFLECT_101."""
    h, w = img.shape
    a = img.astype(np.int32)
    ym, yp = _reflect101(np.arange(h) - 1, h), _reflect101(np.arange(h) + 1, h)
    xm, xp = _reflect101(np.arange(w) - 1, w), _reflect101(np.arange(w) + 1, w)
    t0 = (a[ym] + a[yp]) * 3 + a * 10
    t1 = a[yp] - a[ym]
    dx = t0[:, xp] - t0[:, xm]
    dy = (t1[:, xp] + t1[:, xm]) * 3 + t1 * 10
    return dx.astype(np.int16), dy.astype(np.int16)


def build_pyramid(img, max_level, win):
    """cv::buildOpticalFlowPyramid: levels stop before one would be no larger than the window in either direction."""
    levels = [np.ascontiguousarray(img, np.uint8)]
    for _ in range(max_level):
        h, w = levels[-1].shape
        if (w + 1) // 2 <= win or (h + 1) // 2 <= win:
            break
        levels.append(pyr_down(levels[-1]))
    return levels


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _weights(a, b):
    s = np.float32(1 << W_BITS)
    one = np.float32(1.0)
    iw00 = int(np.rint((one - a) * (one - b) * s))   # cvRound: round half to even
    iw01 = int(np.rint(a * (one - b) * s))
    iw10 = int(np.rint((one - a) * b * s))
    return iw00, iw01, iw10, (1 << W_BITS) - iw00 - iw01 - iw10


def _interp(A, w, n):
    return _descale(A[:-1, :-1] * w[0] + A[:-1, 1:] * w[1] + A[1:, :-1] * w[2] + A[1:, 1:] * w[3], n)


def track(prev, nxt, pts, win=31, max_level=3, max_count=20, eps=0.01, min_eig_threshold=1e-4):
    """-> (next points float32 [n, 2], status uint8 [n], err float32 [n] = min eigenvalue at level 0)."""
    P, N = build_pyramid(prev, max_level, win), build_pyramid(nxt, max_level, win)
    max_level = len(P) - 1
    D = [scharr(p) for p in P]
    pts = np.asarray(pts, np.float32).reshape(-1, 2)
    half = np.float32((win - 1) * 0.5)
    out = np.zeros((len(pts), 2), np.float32)
    status = np.ones(len(pts), np.uint8)
    err = np.zeros(len(pts), np.float32)
    eps2 = float(eps) * float(eps)
    f32eps = np.finfo(np.float32).eps
    for pi in range(len(pts)):
        next_pt = None
        for level in range(max_level, -1, -1):
            I, J = P[level], N[level]
            dx, dy = D[level]
            h, w = I.shape
            prev_pt = pts[pi] * np.float32(1.0 / (1 << level))
            next_pt = prev_pt.copy() if level == max_level else next_pt * np.float32(2.0)
            out[pi] = next_pt
            prev_pt = prev_pt - half
            ip = np.floor(prev_pt).astype(np.int64)
            if ip[0] < -win or ip[0] >= w or ip[1] < -win or ip[1] >= h:
                if level == 0:
                    status[pi] = 0
                    err[pi] = 0
                continue
            wts = _weights(np.float32(prev_pt[0] - ip[0]), np.float32(prev_pt[1] - ip[1]))
            ys, xs = np.arange(win + 1) + ip[1], np.arange(win + 1) + ip[0]
            yi, xi = _reflect101(ys, h), _reflect101(xs, w)
            inside = ((ys >= 0) & (ys < h))[:, None] & ((xs >= 0) & (xs < w))[None, :]   # derivative border: zeros
            Iw = _interp(I[np.ix_(yi, xi)].astype(np.int64), wts, W_BITS - 5)
            Ix = _interp(np.where(inside, dx[np.ix_(yi, xi)].astype(np.int64), 0), wts, W_BITS)
            Iy = _interp(np.where(inside, dy[np.ix_(yi, xi)].astype(np.int64), 0), wts, W_BITS)
            A11 = (Ix * Ix).astype(np.float32).sum(dtype=np.float32) * FLT_SCALE
            A12 = (Ix * Iy).astype(np.float32).sum(dtype=np.float32) * FLT_SCALE
            A22 = (Iy * Iy).astype(np.float32).sum(dtype=np.float32) * FLT_SCALE
            det = np.float32(A11 * A22 - A12 * A12)
            min_eig = np.float32((A22 + A11 - np.sqrt(np.float32((A11 - A22) * (A11 - A22) + np.float32(4.0) * A12 * A12))) / np.float32(2 * win * win))
            if level == 0:
                err[pi] = min_eig
            if min_eig < min_eig_threshold or det < f32eps:
                if level == 0:
                    status[pi] = 0
                continue
            inv = np.float32(1.0) / det
            next_pt = next_pt - half
            prev_delta = np.zeros(2, np.float32)
            for j in range(max_count):
                inp = np.floor(next_pt).astype(np.int64)
                if inp[0] < -win or inp[0] >= w or inp[1] < -win or inp[1] >= h:
                    if level == 0:
                        status[pi] = 0
                    break
                wj = _weights(np.float32(next_pt[0] - inp[0]), np.float32(next_pt[1] - inp[1]))
                yi, xi = _reflect101(np.arange(win + 1) + inp[1], h), _reflect101(np.arange(win + 1) + inp[0], w)
                diff = _interp(J[np.ix_(yi, xi)].astype(np.int64), wj, W_BITS - 5) - Iw
                b1 = (diff * Ix).astype(np.float32).sum(dtype=np.float32) * FLT_SCALE
                b2 = (diff * Iy).astype(np.float32).sum(dtype=np.float32) * FLT_SCALE
                delta = np.array([(A12 * b2 - A22 * b1) * inv, (A12 * b1 - A11 * b2) * inv], np.float32)
                next_pt = next_pt + delta
                out[pi] = next_pt + half
                if float(delta[0]) * float(delta[0]) + float(delta[1]) * float(delta[1]) <= eps2:
                    break
                if j > 0 and abs(float(delta[0] + prev_delta[0])) < 0.01 and abs(float(delta[1] + prev_delta[1])) < 0.01:
                    out[pi] = out[pi] - delta * np.float32(0.5)
                    break
                prev_delta = delta
            next_pt = out[pi].copy()
    return out, status, err
