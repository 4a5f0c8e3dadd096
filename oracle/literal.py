"""oracle/literal.py — TEST INFRASTRUCTURE. A second, independently written restatement of the two pieces of the reference
on which every bit-exact claim rests, in plain Python (ints + numpy.float32 scalars, one rounding per operation, no
vectorisation), used only to cross-check the C++ oracle on random inputs (tests/test_oracle_literal.py):

  * the motion-vector loop of VideoDecoder::NextImage   (src/VideoDecoder.cc:202-350)
  * EXPRESS: compute_center / compute_descriptor / compute_express / diagonal   (include/EXPRESS.h:20-192)

It models the reference's containers literally: a deque of frames, cv::Mat ROIs as (image, x0, y0, cols, rows) with the
parent image's row stride, uint8 counters that wrap. The one deliberate deviation is shared with oracle/raster.cc and
DESIGN.md §4: `vqueue[(size-1)-ref]` with a negative index (undefined behaviour in the reference) drops the entry.
"""
import numpy as np

F32 = np.float32


def f32(x):
    return F32(x)


class VideoImage:
    """include/Frame.h:109-156: mvi = Mat(H, W, CV_32SC4, Scalar(-1,-1,-1,-1)), kps, mvs, coverageArea."""

    def __init__(self, width, height):
        self.mvi = np.full((height, width, 4), -1, np.int32)
        self.kps = []        # (x, y, w, h)
        self.mvs = []        # (mv_x float32, mv_y float32, dIndx)
        self.coverageArea = 0.0


def next_image_mv_loop(width, height, vqueue, records, mv_enabled=True):
    """One pass of src/VideoDecoder.cc:190-353 for a decoded frame whose side data is `records` (dicts with source, w, h,
    src_x, src_y, dst_x, dst_y, ref). `vqueue` is the list of earlier frames; the new frame is appended and returned."""
    smv = VideoImage(width, height)
    if mv_enabled:                                              # :200 if (sd && mv)
        coverage = f32(0)                                       # :204 float coverage = 0
        for mv in records:                                      # :209
            mb_h = f32(mv["h"] * 1)                             # :213-216
            mb_w = f32(mv["w"] * 1)
            mb_h_half = f32(mb_h / f32(2))
            mb_w_half = f32(mb_w / f32(2))
            mv_x = f32(int(mv["dst_x"]) - int(mv["src_x"]))     # :218-219 int arithmetic, then int -> float
            mv_y = f32(int(mv["dst_y"]) - int(mv["src_y"]))
            mv_x = f32(mv_x / f32(mv["ref"] + 1))               # :221-222 float / int
            mv_y = f32(mv_y / f32(mv["ref"] + 1))
            chained = mv["ref"] > 0 and mv["source"] < 0
            dst_x = f32(mv["src_x"] if chained else mv["dst_x"])  # :225-226
            dst_y = f32(mv["src_y"] if chained else mv["dst_y"])
            d_x_top = f32(dst_x - mb_w_half)                    # :228-239
            if d_x_top < 0:
                d_x_top = f32(0)
            d_y_top = f32(dst_y - mb_h_half)
            if d_y_top < 0:
                d_y_top = f32(0)
            d_x_bottom = f32(dst_x + mb_w_half)
            if d_x_bottom >= width:
                continue
            d_y_bottom = f32(dst_y + mb_h_half)
            if d_y_bottom >= height:
                continue
            dIndx = -1                                          # :241
            dMB = (int(d_x_top), int(d_y_top), int(mb_w), int(mb_h))   # cv::Rect(float...) truncates
            if chained:                                         # :243-251
                qi = (len(vqueue) - 1) - mv["ref"]
                if qi >= 0:                                     # negative index: undefined in the reference, dropped here
                    vqueue[qi].kps.append(dMB)
            else:
                smv.kps.append(dMB)
                dIndx = len(smv.kps) - 1
            if mv["source"] > 0:                                # :253-284 B frames feed bmap, which nothing reads
                continue
            for j in range(mv["ref"] + 1, 0, -1):               # :288
                src_x = f32(f32(mv["dst_x"]) + f32(f32(mv_x * f32(j)) * f32(-1)))   # :290-291
                src_y = f32(f32(mv["dst_y"]) + f32(f32(mv_y * f32(j)) * f32(-1)))
                s_x_top = f32(src_x - mb_w_half)                # :294-305
                if s_x_top < 0:
                    s_x_top = f32(0)
                s_y_top = f32(src_y - mb_h_half)
                if s_y_top < 0:
                    s_y_top = f32(0)
                s_x_bottom = f32(src_x + mb_w_half)
                if s_x_bottom >= width:
                    s_x_bottom = f32(width - 1)
                s_y_bottom = f32(src_y + mb_h_half)
                if s_y_bottom >= height:
                    s_y_bottom = f32(height - 1)
                if j == 1:                                      # :314-322
                    sp = smv
                else:
                    qi = len(vqueue) - (j - 1)
                    sp = vqueue[qi] if qi >= 0 else None        # before the clip start: dropped (see module docstring)
                if sp is not None:
                    sp.mvs.append((mv_x, mv_y, dIndx))          # :324
                    sMB_size = len(sp.mvs) - 1                  # :326
                    h = int(s_y_top)                            # :329 for (int h = s_y_top; h <= s_y_bottom; h++)
                    while h <= s_y_bottom:                      #      int compared with float
                        w = int(s_x_top)
                        while w <= s_x_bottom:
                            v = sp.mvi[h, w]
                            if v[0] == -1:                      # :335-342
                                v[0] = sMB_size
                            elif v[1] == -1:
                                v[1] = sMB_size
                            elif v[2] == -1:
                                v[2] = sMB_size
                            else:
                                v[3] = sMB_size
                            w += 1
                        h += 1
            coverage = f32(coverage + f32(dMB[2] * dMB[3]))     # :346 coverage += dMB.area()
        smv.coverageArea = float(coverage) / float(width * height)   # :349 float / double
    vqueue.append(smv)                                          # :351
    return smv


# ---- EXPRESS (include/EXPRESS.h) ------------------------------------------------------------------------------------
def _tables(rows, cols):
    """The literal tables of EXPRESS.h:20-38 for the four supported shapes: length, start row, start column per direction."""
    T = {
        (8, 8): ([1, 2, 3, 4, 5, 6, 7, 8, 7, 6, 5, 4, 3, 2, 1], [7, 6, 5, 4, 3, 2, 1, 0, 0, 0, 0, 0, 0, 0, 0],
                 [[7, 7, 7, 7, 7, 7, 7, 7, 6, 5, 4, 3, 2, 1, 0], [0, 0, 0, 0, 0, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7]]),
        (16, 8): ([1, 2, 3, 4, 5, 6, 7, 8, 8, 8, 8, 8, 8, 8, 8, 8, 7, 6, 5, 4, 3, 2, 1],
                  [15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0, 0, 0, 0, 0, 0, 0, 0],
                  [[7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 7, 6, 5, 4, 3, 2, 1, 0],
                   [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7]]),
        (8, 16): ([1, 2, 3, 4, 5, 6, 7, 8, 8, 8, 8, 8, 8, 8, 8, 8, 7, 6, 5, 4, 3, 2, 1],
                  [7, 6, 5, 4, 3, 2, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
                  [[15, 15, 15, 15, 15, 15, 15, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0],
                   [0, 0, 0, 0, 0, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15]]),
        (16, 16): ([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1],
                   [15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
                   [[15] * 16 + [14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0],
                    [0] * 16 + [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15]]),
    }
    return T[(rows, cols)]


class Roi:
    """cv::Mat header over a parent uint8 image: data pointer (flat offset), rows, cols, step = parent row stride."""

    def __init__(self, img, x0, y0, cols, rows):
        self.buf = np.ascontiguousarray(img, np.uint8).reshape(-1)
        self.step = img.shape[1]
        self.data = y0 * self.step + x0
        self.rows, self.cols = rows, cols

    def at(self, row, col):          # img.at<uint8_t>(row, col)
        return int(self.buf[self.data + row * self.step + col])


def compute_center(m):
    center_row = (m.rows // 2) & 0xff                        # :81-82 uint8_t
    center_col = (m.cols // 2) & 0xff
    return ((m.at(center_col, center_row) + m.at(center_col - 1, center_row - 1) + m.at(center_col, center_row - 1) +
             m.at(center_col - 1, center_row)) // 4) & 0xff   # :83-87: at(row = center_col, col = center_row), int sum / 4 -> uint8


def _bounds(m, threshold):
    center = compute_center(m)
    return (center - threshold) & 0xff, (center + threshold) & 0xff   # :92-93 uint8 wrap


def compute_descriptor(m, threshold):
    """:90-110 -> 256-bit descriptor as a Python int (bit i = desc[i])."""
    low, high = _bounds(m, threshold)
    desc = 0
    for y in range(m.rows):
        p = m.data + y * m.step                               # uchar *p = img.ptr(y)
        for x in range(m.cols):
            p += 1                                            # p++ BEFORE the read
            v = int(m.buf[p])
            if low > v or high < v:
                desc |= 1 << (y * m.rows + x)                 # desc.set(y * img.rows + x)
    return desc


def compute_express(m, threshold):
    """:117-192."""
    low, high = _bounds(m, threshold)
    precheck = int(m.rows * m.cols * .125) & 0xff
    f = 0
    for row in range(m.rows):
        p = m.data + row * m.step
        for col in range(m.cols):
            p += 1
            v = int(m.buf[p])
            if low > v or high < v:
                f = (f + 1) & 0xff                            # uint8_t f
        if f >= precheck:
            break
    if f < precheck:
        return False
    slices = (m.rows + m.cols - 1) & 0xff
    rounds = int(np.floor(slices * .25 + .5)) & 0xff          # round(): half away from zero, argument positive
    u_rounds = (slices - rounds) & 0xff
    L, S, R = _tables(m.rows, m.cols)
    for a in range(2):
        direction = 1 if a == 0 else 0                        # diagonal(img, i, a == 0)
        wins = losses = 0
        for i in range(slices):
            length = L[i]                                     # diagonal(): :40-77
            ptr = m.data + m.step * S[i] + R[direction][i]
            stride = m.step + (1 if direction else -1)
            win = loss = 0
            for r in range(length):
                v = int(m.buf[ptr + r * stride])
                if low > v or high < v:
                    win += 1
                else:
                    loss += 1
            if wins < rounds:
                wins = wins + 1 if win >= loss else 0
            if losses < rounds:
                losses = losses + 1 if loss > win else 0
            if i > u_rounds and (wins == 0 or losses == 0):
                break
        if wins >= rounds and losses >= rounds:
            return True
    return False


# ---- MOVExtractor::operator() (src/MOVExtractor.cc:63-455), LK-carried features dropped (lk_status == NULL mode) -------
class VideoFeature:
    """include/Frame.h:79-107."""

    def __init__(self, trackId, qIndx, pt, mb, age, desc, coverage=False):
        self.trackId, self.qIndx, self.pt, self.mb, self.age, self.desc, self.coverage = trackId, qIndx, pt, mb, age, desc, coverage


def _popcount(x):
    return bin(x).count("1")


def _rect_from(ptx, pty, w, h):
    # cv::Rect mb(pt.x - (mb.width / 2), pt.y - (mb.height / 2), w, h): float - int -> float, truncated by the Rect ctor
    return (int(f32(ptx - f32(w // 2))), int(f32(pty - f32(h // 2))), w, h)


def _inside(mb, cols, rows):
    return mb[0] >= 0 and mb[1] >= 0 and (mb[0] + mb[2]) < cols and (mb[1] + mb[3]) < rows


def extract_moves(img, threshold):
    """:39-61 -> [(pt, desc)] on the 16-px lattice."""
    rows, cols = img.shape
    out = []
    for y in range(8, rows - 8, 16):
        for x in range(8, cols - 8, 16):
            mb = (x - 8, y - 8, 16, 16)
            if _inside(mb, cols, rows):
                m = Roi(img, mb[0], mb[1], 16, 16)
                if compute_express(m, threshold):
                    out.append(((f32(x), f32(y)), compute_descriptor(m, threshold)))
    return out


def extractor(smv, img, is_p_frame, prev_vf, current_id, threshold, coverage_threshold):
    """Returns (new mvVF list, mCurrentId). prev_vf is sorted in place like prev->mvVF (:249-252, as a stable sort).
    Features the reference hands to cv::calcOpticalFlowPyrLK (coverage features, every feature on an I frame) are dropped."""
    rows, cols = img.shape
    vf_out = []
    lbFound = [False] * len(smv.kps)                                     # :68
    mov_cnt = 0
    if not is_p_frame:                                                   # :78 I frame
        if prev_vf:                                                      # :80-120 LK carry-over only: host work, dropped
            return vf_out, current_id
        for pt, desc in extract_moves(img, threshold):                   # :123-157 (same lattice walk)
            current_id += 1
            vf_out.append(VideoFeature(current_id, -1, pt, (int(pt[0]) - 8, int(pt[1]) - 8, 16, 16), 0, desc))
        return vf_out, current_id
    if prev_vf:
        prev_vf.sort(key=lambda a: (-a.age, -_popcount(a.desc)))         # :249-252, Python's sort is stable
        for i, pvf in enumerate(prev_vf):
            if pvf.coverage:                                             # :258-262 -> covFeat -> LK (dropped)
                continue
            x, y = int(pvf.pt[0]), int(pvf.pt[1])                        # :264
            cell = smv.mvi[y, x]
            if cell[0] == -1:                                            # :265
                continue
            indx = int(cell[0])                                          # :270
            if cell[1] >= 0:                                             # :272
                bestDesc = 256
                for j in range(4):
                    if cell[j] == -1:
                        break
                    mvx, mvy, _ = smv.mvs[int(cell[j])]
                    ptx, pty = f32(pvf.pt[0] + mvx), f32(pvf.pt[1] + mvy)   # :283
                    mb = _rect_from(ptx, pty, pvf.mb[2], pvf.mb[3])
                    if _inside(mb, cols, rows):                          # :286
                        dist = _popcount(pvf.desc ^ compute_descriptor(Roi(img, *mb), threshold))
                        if dist < bestDesc:                              # :292 strict
                            bestDesc = dist
                            indx = int(cell[j])
            mvx, mvy, dIndx = smv.mvs[indx]                              # :301
            ptx, pty = f32(pvf.pt[0] + mvx), f32(pvf.pt[1] + mvy)
            mb = _rect_from(ptx, pty, pvf.mb[2], pvf.mb[3])
            if (dIndx == -1 or not lbFound[dIndx]) and _inside(mb, cols, rows):   # :306 (imageCols/Rows == image size)
                if dIndx >= 0:
                    lbFound[dIndx] = True
                desc = compute_descriptor(Roi(img, *mb), threshold)
                if _popcount(pvf.desc ^ desc) <= 40:                     # :316
                    vf_out.append(VideoFeature(pvf.trackId, i, (ptx, pty), mb, pvf.age + 1, desc))
    for i, mb in enumerate(smv.kps):                                     # :379-416 new features
        if lbFound[i]:
            continue
        # (mb.br() + mb.tl()) * 0.5: integer points summed, scaled in double, saturate_cast<int> (cvRound) back to
        # Point_<int>, then converted to Point2f
        ptx = f32(int(np.rint((mb[0] + mb[2] + mb[0]) * 0.5)))
        pty = f32(int(np.rint((mb[1] + mb[3] + mb[1]) * 0.5)))
        if _inside(mb, cols, rows):
            m = Roi(img, *mb)
            if compute_express(m, threshold):
                current_id += 1
                vf_out.append(VideoFeature(current_id, -1, (ptx, pty), mb, 0, compute_descriptor(m, threshold)))
                mov_cnt += 1
    if smv.coverageArea < coverage_threshold or mov_cnt < 60:            # :418-451 back-fill from the lattice
        for pt, desc in extract_moves(img, threshold):
            mb = (int(f32(pt[0] - f32(8))), int(f32(pt[1] - f32(8))), 16, 16)
            if _inside(mb, cols, rows):
                if smv.mvi[int(pt[1]), int(pt[0])][0] >= 0:
                    continue
                current_id += 1
                vf_out.append(VideoFeature(current_id, -1, pt, mb, 0, desc, coverage=True))
    return vf_out, current_id


# ---- pose-only Gauss-Newton / Huber, second restatement -------------------------------------
# Independent of oracle/pose.cc in its building blocks: the SE3 exponential is scipy's matrix exponential of the 4x4 twist,
# the Jacobian of the residual is taken by central differences of the projection (no analytic formula), the normal
# equations are solved by numpy's LU solver, Huber is restated from its definition.
def _project(cam, Xc):
    x, y, z = Xc
    if cam["model"] == 0:                                    # Pinhole.cpp:45-52
        return np.array([cam["fx"] * x / z + cam["cx"], cam["fy"] * y / z + cam["cy"]])
    r = np.hypot(x, y)                                       # KannalaBrandt8 (App. A.6)
    if r < 1e-12:
        return np.array([float(cam["cx"]), float(cam["cy"])])
    th = np.arctan2(r, z)
    k = [float(v) for v in cam["k"]]
    thd = th * (1 + k[0] * th ** 2 + k[1] * th ** 4 + k[2] * th ** 6 + k[3] * th ** 8)
    return np.array([cam["fx"] * thd * x / r + cam["cx"], cam["fy"] * thd * y / r + cam["cy"]])


def se3_exp_expm(dx):
    """exp of the twist [omega, upsilon] (rotation first, g2o SE3Quat::exp) through scipy.linalg.expm."""
    from scipy.linalg import expm
    w, v = dx[:3], dx[3:]
    M = np.zeros((4, 4))
    M[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
    M[:3, 3] = v
    E = expm(M)
    return E[:3, :3], E[:3, 3]


def residual_jacobian_fd(cam, Xc, h=1e-6):
    """d e / d dx at dx = 0 for e = obs - pi(exp(dx) * Xc), by central differences: 2 x 6."""
    J = np.zeros((2, 6))
    for k in range(6):
        d = np.zeros(6)
        d[k] = h
        Rp, tp = se3_exp_expm(d)
        Rm, tm = se3_exp_expm(-d)
        J[:, k] = -(_project(cam, Rp @ Xc + tp) - _project(cam, Rm @ Xc + tm)) / (2 * h)
    return J


def pose_optimize_ref(cam, pts, obs, R, t, rep_error, iteration_count, jac):
    """The schedule of App. A.5: 4 rounds of iteration_count/4 Gauss-Newton steps, Huber (delta = rep_error) in the first
    three, re-classification (chi2 > rep_error^2 -> outlier, excluded from the next round) after every round, early exit
    at |dx|_inf < 1e-10 and below 3 inliers. `jac(cam, Xc)` supplies the 2x6 residual Jacobian. Returns (R, t, outlier)."""
    pts, obs = np.asarray(pts, np.float64), np.asarray(obs, np.float64)
    n = len(pts)
    outlier = np.zeros(n, bool)
    if n < 4:
        return R, t, outlier, 0
    delta = float(np.float32(rep_error))
    its = max(iteration_count // 4, 1)
    for rnd in range(4):
        robust = rnd < 3
        for _ in range(its):
            H, b = np.zeros((6, 6)), np.zeros(6)
            for i in range(n):
                if outlier[i]:
                    continue
                Xc = R @ pts[i] + t
                if not Xc[2] > 0:
                    continue
                e = obs[i] - _project(cam, Xc)
                chi2 = float(e @ e)
                w = 1.0
                if robust and chi2 > delta * delta:           # Huber: rho'(chi2) = delta / sqrt(chi2) beyond delta^2
                    w = delta / np.sqrt(chi2)
                J = jac(cam, Xc)
                H += w * J.T @ J
                b -= w * J.T @ e
            try:
                dx = np.linalg.solve(H, b)
            except np.linalg.LinAlgError:
                break
            dR, dt = se3_exp_expm(dx)
            R, t = dR @ R, dR @ t + dt
            if np.max(np.abs(dx)) < 1e-10:
                break
        bad = 0
        for i in range(n):
            Xc = R @ pts[i] + t
            o = True
            if Xc[2] > 0:
                e = obs[i] - _project(cam, Xc)
                o = float(e @ e) > delta * delta
            outlier[i] = o
            bad += o
        if n - bad < 3:
            break
    return R, t, outlier, n - int(outlier.sum())


# ---- Frame::isInFrustum, mono branch (src/Frame.cc:456-519) + the joins (include/MOVMatcher.h:35-137) -------------------
def _dot3(a, b):
    # float32, no contraction, Eigen 3.4's unrolled 3-term reduction order c0 + (c1 + c2) (see oracle/match.cc's header)
    return f32(f32(a[0] * b[0]) + f32(f32(a[1] * b[1]) + f32(a[2] * b[2])))


def is_in_frustum(Rcw, tcw, cam, width, height, cos_limit, pos, normal, min_dist, max_dist):
    """Pinhole only. Returns (in_view, u, v, depth, view_cos) as the reference leaves them in the MapPoint."""
    R = [[f32(Rcw[i][j]) for j in range(3)] for i in range(3)]
    t = [f32(v) for v in tcw]
    Ow = [f32(-(Rcw[0][i] * tcw[0] + Rcw[1][i] * tcw[1] + Rcw[2][i] * tcw[2])) for i in range(3)]   # -R^T t in double
    P = [f32(v) for v in pos]
    out = [0, f32(-1), f32(-1), f32(0), f32(0)]                           # :460-462
    Pc = [f32(_dot3(R[i], P) + t[i]) for i in range(3)]                   # :468
    Pc_dist = f32(np.sqrt(_dot3(Pc, Pc)))                                 # :469
    if Pc[2] < f32(0):                                                    # :474
        return out
    u = f32(f32(f32(f32(cam["fx"]) * Pc[0]) / Pc[2]) + f32(cam["cx"]))    # Pinhole.cpp:48-49
    v = f32(f32(f32(f32(cam["fy"]) * Pc[1]) / Pc[2]) + f32(cam["cy"]))
    if u < f32(0) or u > f32(width) or v < f32(0) or v > f32(height):     # :479-482 with mnMin/Max of an undistorted frame
        return out
    out[1], out[2] = u, v                                                 # :484-485
    maxD, minD = f32(f32(1.2) * f32(max_dist)), f32(f32(0.8) * f32(min_dist))   # MapPoint.cc:443-453
    PO = [f32(P[i] - Ow[i]) for i in range(3)]                            # :490
    dist = f32(np.sqrt(_dot3(PO, PO)))
    if dist < minD or dist > maxD:                                        # :493
        return out
    view_cos = f32(_dot3(PO, [f32(x) for x in normal]) / dist)            # :499
    if view_cos < f32(cos_limit):                                         # :501
        return out
    out[0], out[3], out[4] = 1, Pc_dist, view_cos                         # :508-516
    return out


def search_by_video_feature(track_ids, mp_track_ids, mp_ok, match):
    """MOVMatcher.h:35-68: vfmap = first index per id (std::map::insert never overwrites); every eligible map point, in
    order, overwrites the match of its track (last wins). Returns nmatches; `match` is updated in place."""
    vfmap = {}
    for i, tid in enumerate(track_ids):
        vfmap.setdefault(int(tid), i)
    n = 0
    for k, tid in enumerate(mp_track_ids):
        if not mp_ok[k]:
            continue
        if int(tid) in vfmap:
            match[vfmap[int(tid)]] = k
            n += 1
    return n


def assign_features_to_grid(xs, ys, width, height):
    """Frame.cc:356-388 with PosInGrid (:670-680, round()) -> {(ix, iy): [indices in insertion order]}."""
    w_inv, h_inv = f32(f32(64) / f32(width)), f32(f32(48) / f32(height))
    grid = {}
    for i, (x, y) in enumerate(zip(xs, ys)):
        fx, fy = f32(f32(f32(x) - f32(0)) * w_inv), f32(f32(f32(y) - f32(0)) * h_inv)
        px = int(np.floor(abs(fx) + f32(0.5))) * (1 if fx >= 0 else -1)   # C round(): half away from zero
        py = int(np.floor(abs(fy) + f32(0.5))) * (1 if fy >= 0 else -1)
        if px < 0 or px >= 64 or py < 0 or py >= 48:
            continue
        grid.setdefault((px, py), []).append(i)
    return grid
