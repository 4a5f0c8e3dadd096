"""Seeded synthetic inputs of the shapes BASELINE.json names.

Scene: a textured plane at Z = 5 m facing the camera; the camera sweeps sideways sinusoidally (<= 0.03 m/frame)
with a small yaw (<= 0.2 deg/frame). MV records mimic ffmpeg's H.264 exporter (ffmpeg-ref-patch.patch:15-98):
block centres on the macroblock lattice, w,h in {16, 8}, quarter-pel motion, src = dst + motion/scale with C
truncation, ref index uniform in [0, min(R-1, frame-1)]. Everything derives from numpy's PCG64 seeded with
0x5EED0000 + config id + stream id, so the oracle and the CUDA path always see identical arrays.
"""
import numpy as np

from . import types as T

PLANE_Z = 5.0
TEX_N = 1024          # texture tile is TEX_N x TEX_N texels, periodic
TEX_SCALE = 64.0      # texels per metre (1 texel ~ 1 px at Z = 5 m with fx = 320)
_TEX_CACHE = {}


def texture(seed=0x5EED0000):
    """Piecewise-constant random rectangles: edges cross most 16x16 blocks, which is what EXPRESS looks for."""
    if seed not in _TEX_CACHE:
        rng = np.random.Generator(np.random.PCG64(seed))
        tex = np.full((TEX_N, TEX_N), 128, np.uint8)
        n = 5000
        xs, ys = rng.integers(0, TEX_N, n), rng.integers(0, TEX_N, n)
        ws, hs = rng.integers(10, 56, n), rng.integers(10, 56, n)
        vs = rng.choice(np.array([20, 60, 100, 140, 180, 220], np.uint8), n)
        for x, y, w, h, v in zip(xs, ys, ws, hs, vs):
            yy = np.arange(y, y + h) % TEX_N
            xx = np.arange(x, x + w) % TEX_N
            tex[np.ix_(yy, xx)] = v
        _TEX_CACHE[seed] = tex
    return _TEX_CACHE[seed]


class Spec:
    def __init__(self, width=640, height=480, n_frames=20, refs=4, seed=0x5EED0001, fx=320.0, fy=320.0, cx=None,
                 cy=None, stereo=False, dense4x4=False, baseline=0.25, phase=0.0, start_p=False):
        self.W, self.H, self.n_frames, self.refs, self.seed = width, height, n_frames, refs, seed
        self.fx, self.fy = fx, fy
        self.cx = width / 2 if cx is None else cx
        self.cy = height / 2 if cy is None else cy
        self.stereo, self.dense4x4, self.baseline, self.phase = stereo, dense4x4, baseline, phase
        self.start_p = start_p      # mid-stream clip: frame 0 is a P frame whose records reference frames before the clip

    def camera(self):
        return T.camera(self.fx, self.fy, self.cx, self.cy)


def pose_at(spec, f):
    """(R_cw, t_cw) of frame f (any integer), double. Stereo: frame 2k = left view at time k, 2k+1 = right view."""
    k, right = (f // 2, f % 2 == 1) if spec.stereo else (f, False)
    a = 0.12 * k + spec.phase
    x = 0.25 * np.sin(a)                 # <= 0.03 m / frame
    y = 0.05 * np.sin(0.7 * a + 1.0)
    yaw = np.deg2rad(1.2) * np.sin(0.15 * k + 0.5 * spec.phase)   # <= 0.2 deg / frame
    c, s = np.cos(yaw), np.sin(yaw)
    Rwc = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    Ow = np.array([x, y, 0.0])
    if right:
        Ow = Ow + Rwc @ np.array([spec.baseline, 0, 0])
    Rcw = Rwc.T
    return Rcw, -Rcw @ Ow


def trajectory(spec):
    return [pose_at(spec, f) for f in range(spec.n_frames)]


def _backproject(spec, pose, u, v):
    """Pixel -> point on the plane Z = PLANE_Z (world)."""
    Rcw, tcw = pose
    d = np.stack([(u - spec.cx) / spec.fx, (v - spec.cy) / spec.fy, np.ones_like(u, dtype=np.float64)], -1)
    dw = d @ Rcw            # R_wc d = R_cw^T d
    Ow = -Rcw.T @ tcw
    lam = (PLANE_Z - Ow[2]) / dw[..., 2]
    return Ow + lam[..., None] * dw


def _project(spec, pose, Xw):
    Rcw, tcw = pose
    Xc = Xw @ Rcw.T + tcw
    return spec.fx * Xc[..., 0] / Xc[..., 2] + spec.cx, spec.fy * Xc[..., 1] / Xc[..., 2] + spec.cy


def make_records(spec):
    """-> (recs MV_RECORD[], rec_off int64[n_frames+1], frame_flags uint8[n_frames])."""
    rng = np.random.Generator(np.random.PCG64(spec.seed))
    traj = trajectory(spec)
    all_recs, off, flags = [], [0], []
    for f in range(spec.n_frames):
        is_right = spec.stereo and f % 2 == 1
        fl = (T.FRAME_P if (f > 0 or spec.start_p) else 0)
        if (f == 0 and not spec.start_p) or is_right:
            flags.append(fl)
            off.append(off[-1])
            continue
        fl |= T.FRAME_MV
        if spec.dense4x4:
            gx, gy = np.meshgrid(np.arange(spec.W // 4), np.arange(spec.H // 4))
            cxs, cys = (4 * gx + 2).ravel(), (4 * gy + 2).ravel()
            ws = np.full(cxs.shape, 4, np.uint8)
            hs = ws.copy()
        else:
            mbw, mbh = spec.W // 16, spec.H // 16
            part = rng.choice(5, size=(mbh, mbw), p=[0.5, 0.125, 0.125, 0.2, 0.05])
            # records are emitted in macroblock raster order, sub-blocks in ffmpeg's order
            sub = {0: [(8, 8, 16, 16)], 1: [(8, 4, 16, 8), (8, 12, 16, 8)], 2: [(4, 8, 8, 16), (12, 8, 8, 16)],
                   3: [(4, 4, 8, 8), (12, 4, 8, 8), (4, 12, 8, 8), (12, 12, 8, 8)], 4: []}
            counts = np.array([1, 2, 2, 4, 0])[part.ravel()]
            n = counts.sum()
            cxs, cys = np.zeros(n, np.int64), np.zeros(n, np.int64)
            ws, hs = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
            starts = np.concatenate([[0], np.cumsum(counts)])[:-1]
            pr = part.ravel()
            for p, lst in sub.items():
                idx = np.nonzero(pr == p)[0]
                for i, (ox, oy, w, h) in enumerate(lst):
                    k = starts[idx] + i
                    cxs[k] = 16 * (idx % mbw) + ox
                    cys[k] = 16 * (idx // mbw) + oy
                    ws[k], hs[k] = w, h
        n = len(cxs)
        # stereo left frames: ref 0 = the right view just before, ref 1 = the previous left view
        max_ref = spec.refs - 1 if spec.start_p else min(spec.refs - 1, f - 1)
        refs = rng.integers(0, max_ref + 1, n) if max_ref > 0 else np.zeros(n, np.int64)
        Xw = _backproject(spec, traj[f], cxs.astype(np.float64), cys.astype(np.float64))
        sx, sy = np.zeros(n), np.zeros(n)
        for r in range(max_ref + 1):
            m = refs == r
            if m.any():
                sx[m], sy[m] = _project(spec, pose_at(spec, f - 1 - r), Xw[m])
        mot_x = np.clip(np.rint(4 * (sx - cxs)), -256, 256).astype(np.int64)
        mot_y = np.clip(np.rint(4 * (sy - cys)), -256, 256).astype(np.int64)
        recs = np.zeros(n, T.MV_RECORD)
        recs["source"], recs["w"], recs["h"] = -1, ws, hs
        recs["dst_x"], recs["dst_y"] = cxs, cys
        # src = dst + motion / motion_scale, C integer division (truncation toward zero), patch:44-45
        recs["src_x"] = cxs + np.trunc(mot_x / 4).astype(np.int64)
        recs["src_y"] = cys + np.trunc(mot_y / 4).astype(np.int64)
        recs["motion_x"], recs["motion_y"], recs["motion_scale"] = mot_x, mot_y, 4
        recs["ref"] = refs
        all_recs.append(recs)
        off.append(off[-1] + n)
        flags.append(fl)
    recs = np.concatenate(all_recs) if all_recs else np.zeros(0, T.MV_RECORD)
    return recs, np.array(off, np.int64), np.array(flags, np.uint8)


def make_grey(spec, frames=None):
    """uint8 [n, H, W]: the texture seen through each frame's pose (nearest-neighbour)."""
    tex = texture()
    traj = trajectory(spec)
    frames = range(spec.n_frames) if frames is None else frames
    v, u = np.mgrid[0:spec.H, 0:spec.W].astype(np.float64)
    out = np.zeros((len(frames), spec.H, spec.W), np.uint8)
    for i, f in enumerate(frames):
        Xw = _backproject(spec, traj[f], u, v)
        tx = np.floor(Xw[..., 0] * TEX_SCALE).astype(np.int64) % TEX_N
        ty = np.floor(Xw[..., 1] * TEX_SCALE).astype(np.int64) % TEX_N
        out[i] = tex[ty, tx]
    return out


def seed_tracks_lattice(spec):
    """MV-only configs seed 16x16 tracks on the 16-px lattice as input state."""
    ys = np.arange(8, spec.H - 8, 16)
    xs = np.arange(8, spec.W - 8, 16)
    gx, gy = np.meshgrid(xs, ys)
    gx, gy = gx.ravel(), gy.ravel()
    ok = (gx - 8 + 16 < spec.W) & (gy - 8 + 16 < spec.H)
    gx, gy = gx[ok], gy[ok]
    tr = np.zeros(len(gx), T.TRACK)
    tr["pt_x"], tr["pt_y"] = gx, gy
    tr["mb"]["x"], tr["mb"]["y"], tr["mb"]["w"], tr["mb"]["h"] = gx - 8, gy - 8, 16, 16
    tr["track_id"] = np.arange(1, len(gx) + 1)
    tr["q_indx"] = -1
    return tr


def map_from_tracks(spec, tracks, pose):
    """One map point per track: the track centre back-projected onto the plane through `pose` (R_cw, t_cw)."""
    Xw = _backproject(spec, pose, tracks["pt_x"].astype(np.float64), tracks["pt_y"].astype(np.float64))
    Ow = -pose[0].T @ pose[1]
    d = Xw - Ow
    dist = np.linalg.norm(d, axis=1)
    mp = np.zeros(len(tracks), T.MAP_POINT)
    mp["pos"] = Xw
    mp["normal"] = d / dist[:, None]
    mp["min_dist"], mp["max_dist"] = 0.5 * dist, 2.0 * dist
    mp["track_id"] = tracks["track_id"]
    return mp


def pose_struct(pose):
    return T.pose(pose[0], pose[1])


def project_np(cam, Xc):
    """numpy camera model (Pinhole.cpp:37-43; KannalaBrandt8 per SURVEY.md App. A.6), double."""
    x, y, z = Xc[:, 0], Xc[:, 1], Xc[:, 2]
    fx, fy, cx, cy = (float(cam[k]) for k in ("fx", "fy", "cx", "cy"))
    if int(cam["model"]) == T.CAM_FISHEYE:
        k = [float(v) for v in cam["k"]]
        r = np.sqrt(x * x + y * y)
        th = np.arctan2(r, z)
        t2 = th * th
        thd = th * (1 + t2 * (k[0] + t2 * (k[1] + t2 * (k[2] + t2 * k[3]))))
        s = np.where(r > 1e-12, thd / np.maximum(r, 1e-300), 1.0)
        return np.stack([fx * s * x + cx, fy * s * y + cy], 1)
    return np.stack([fx * x / z + cx, fy * y / z + cy], 1)


def pnp_problem(n, cam, seed, sigma=0.5, outlier_frac=0.10, width=640, height=480, perturb=(0.01, 0.03)):
    """Standalone PoseOptimization stress input: points in the frustum, depth U[2,20] m,
    observations pi(T_gt X) + N(0, sigma) px with gross outliers U[-50,50] px; initial pose = perturbed T_gt
    (stands for "previous frame's pose", Tracking.cc:807).
    Returns (pts float32[n,3], obs float32[n,2], pose_gt, pose_init)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    ang = rng.normal(0, 0.05, 3)
    Rgt = _rodrigues(ang)
    tgt = rng.normal(0, 0.2, 3)
    z = rng.uniform(2, 20, n)
    if int(cam["model"]) == T.CAM_FISHEYE:
        th = rng.uniform(0.0, 1.0, n)
        psi = rng.uniform(-np.pi, np.pi, n)
        Xc = np.stack([np.sin(th) * np.cos(psi), np.sin(th) * np.sin(psi), np.cos(th)], 1) * (z / np.cos(th))[:, None]
    else:
        u = rng.uniform(0.05 * width, 0.95 * width, n)
        v = rng.uniform(0.05 * height, 0.95 * height, n)
        Xc = np.stack([(u - float(cam["cx"])) / float(cam["fx"]) * z, (v - float(cam["cy"])) / float(cam["fy"]) * z, z], 1)
    Xw = ((Xc - tgt) @ Rgt).astype(np.float32)          # R^T (Xc - t), stored float like MapPoint::mWorldPos
    obs = project_np(cam, Xw.astype(np.float64) @ Rgt.T + tgt) + rng.normal(0, sigma, (n, 2))
    bad = rng.random(n) < outlier_frac
    obs[bad] += rng.uniform(-50, 50, (int(bad.sum()), 2))
    Ri = _rodrigues(ang + rng.normal(0, perturb[0], 3))
    ti = tgt + rng.normal(0, perturb[1], 3)
    return Xw, obs.astype(np.float32), T.pose(Rgt, tgt), T.pose(Ri, ti)


def _rodrigues(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K
