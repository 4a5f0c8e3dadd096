"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, nowhere else.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(_HERE), "mov-slam_b200", "python"))
from movfe import types as T  # noqa: E402

_LIB = None


def build():
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        vp, i32, i64, f32, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double
        L.orc_raster_clip.restype = vp
        L.orc_raster_clip.argtypes = [i32, i32, i32, vp, vp, vp, i32]
        L.orc_clip_free.argtypes = [vp]
        for name in ("orc_clip_n_hops", "orc_clip_n_kps"):
            getattr(L, name).restype = i32
            getattr(L, name).argtypes = [vp, i32]
        L.orc_clip_coverage.restype = f64
        L.orc_clip_coverage.argtypes = [vp, i32]
        L.orc_clip_bad_ref.restype = i64
        L.orc_clip_bad_ref.argtypes = [vp]
        for name in ("orc_clip_grid", "orc_clip_hops", "orc_clip_kps"):
            getattr(L, name).restype = vp
            getattr(L, name).argtypes = [vp, i32]
        L.orc_express_center.restype = i32
        L.orc_express_center.argtypes = [vp, i32, i32, i32, i32, i32]
        L.orc_express_descriptor.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp]
        L.orc_express_test.restype = i32
        L.orc_express_test.argtypes = [vp, i32, i32, i32, i32, i32, i32]
        L.orc_express_distance.restype = i32
        L.orc_express_distance.argtypes = [vp, vp]
        L.orc_extract_frame.restype = i32
        L.orc_extract_frame.argtypes = [i32, i32, C.c_uint32, vp, vp, vp, vp, i32, f64, vp, i32, vp, vp, vp, vp, vp, vp]
        L.orc_extract_frame_lost.restype = i32
        L.orc_extract_frame_lost.argtypes = [i32, i32, C.c_uint32, vp, vp, vp, vp, i32, f64, vp, i32, vp, vp, vp, i32, vp, vp, vp, vp]
        L.orc_frustum.argtypes = [vp, vp, i32, i32, f32, vp, i32, vp]
        L.orc_update_local_points.argtypes = [vp, i32, vp, i32, i32, vp, i32, vp]
        L.orc_search_by_video_feature.restype = i32
        L.orc_search_by_video_feature.argtypes = [vp, i32, vp, vp, i32, i32, f32, vp]
        L.orc_search_by_keyframe.restype = i32
        L.orc_search_by_keyframe.argtypes = [vp, i32, vp, i32, vp]
        L.orc_search_for_initialization.restype = i32
        L.orc_search_for_initialization.argtypes = [vp, i32, vp, i32, vp, vp]
        L.orc_assign_features_to_grid.argtypes = [vp, i32, i32, i32, vp, vp]
        L.orc_get_features_in_area.restype = i32
        L.orc_get_features_in_area.argtypes = [vp, i32, i32, i32, vp, vp, f32, f32, f32, vp]
        L.orc_search_by_projection.restype = i32
        L.orc_search_by_projection.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, i32, vp, vp, vp, vp]
        L.orc_project.argtypes = [vp, vp, vp]
        L.orc_project_jac.argtypes = [vp, vp, vp]
        L.orc_pose_jacobian.argtypes = [vp, vp, vp]
        L.orc_huber_weight.restype = f64
        L.orc_huber_weight.argtypes = [f64, f64]
        L.orc_se3_exp.argtypes = [vp, vp, vp]
        L.orc_pose_optimize.restype = i32
        L.orc_pose_optimize.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp]
        L.orc_frontend_run.restype = i32
        L.orc_frontend_run.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp, i32, vp, vp]
        L.orc_frontend_run_sched.restype = i32
        L.orc_frontend_run_sched.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp, i32, vp, i32, vp, vp, vp, vp, i32, vp, vp]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Clip:
    """Result of the raster over one stream's clip: per-frame grid / hops / kps / coverage."""

    def __init__(self, width, height, recs, rec_off, frame_flags, max_ref=10):
        self.W, self.H = width, height
        self.n_frames = len(frame_flags)
        recs = np.ascontiguousarray(recs, T.MV_RECORD)
        rec_off = np.ascontiguousarray(rec_off, np.int64)
        frame_flags = np.ascontiguousarray(frame_flags, np.uint8)
        assert len(rec_off) == self.n_frames + 1
        self._h = lib().orc_raster_clip(width, height, self.n_frames, _p(recs), _p(rec_off), _p(frame_flags), max_ref)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_clip_free(self._h)
            self._h = None

    def _arr(self, ptr, n, dtype):
        if n == 0 or not ptr:
            return np.zeros(0, dtype)
        buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype, n).copy()

    def n_hops(self, f):
        return lib().orc_clip_n_hops(self._h, f)

    def n_kps(self, f):
        return lib().orc_clip_n_kps(self._h, f)

    def coverage(self, f):
        return lib().orc_clip_coverage(self._h, f)

    def bad_ref(self):
        return lib().orc_clip_bad_ref(self._h)

    def grid(self, f):
        return self._arr(lib().orc_clip_grid(self._h, f), self.W * self.H * 4, np.int32).reshape(self.H, self.W, 4)

    def hops(self, f):
        return self._arr(lib().orc_clip_hops(self._h, f), self.n_hops(f), T.HOP)

    def kps(self, f):
        return self._arr(lib().orc_clip_kps(self._h, f), self.n_kps(f), T.RECT)


def express_descriptor(img, x0, y0, cols, rows, thr):
    img = np.ascontiguousarray(img, np.uint8)
    d = np.zeros(8, np.uint32)
    lib().orc_express_descriptor(_p(img), img.shape[1], x0, y0, cols, rows, thr, _p(d))
    return d


def express_center(img, x0, y0, cols, rows):
    img = np.ascontiguousarray(img, np.uint8)
    return lib().orc_express_center(_p(img), img.shape[1], x0, y0, cols, rows)


def express_test(img, x0, y0, cols, rows, thr):
    img = np.ascontiguousarray(img, np.uint8)
    return bool(lib().orc_express_test(_p(img), img.shape[1], x0, y0, cols, rows, thr))


def express_distance(a, b):
    a = np.ascontiguousarray(a, np.uint32)
    b = np.ascontiguousarray(b, np.uint32)
    return lib().orc_express_distance(_p(a), _p(b))


EXTRACT_PARAMS = np.dtype([("threshold", "<i4"), ("_p0", "<i4"), ("coverage_threshold", "<f8"),
                           ("max_tracks", "<i4"), ("_p1", "<i4")])
assert EXTRACT_PARAMS.itemsize == 24


def extract_frame(width, height, frame_flags, grey, grid, hops, kps, coverage_area, prev, current_id,
                  threshold=25, coverage_threshold=0.20, max_tracks=4096, lk_status=None, lk_pts=None, reloc=None):
    """Returns (tracks, sorted_prev, new_current_id, n_births). reloc: RELOC_SEED array (lost relocalisation)."""
    grey = None if grey is None else np.ascontiguousarray(grey, np.uint8)
    grid = np.ascontiguousarray(grid, np.int32)
    hops = np.ascontiguousarray(hops, T.HOP)
    kps = np.ascontiguousarray(kps, T.RECT)
    prev = np.array(prev, T.TRACK, copy=True)
    out = np.zeros(max_tracks, T.TRACK)
    ep = np.zeros((), EXTRACT_PARAMS)
    ep["threshold"], ep["coverage_threshold"], ep["max_tracks"] = threshold, coverage_threshold, max_tracks
    cid = np.array([current_id], np.int32)
    nb = np.zeros(1, np.int32)
    if lk_status is not None:
        lk_status = np.ascontiguousarray(lk_status, np.uint8)
        lk_pts = np.ascontiguousarray(lk_pts, np.float32)
    reloc = None if reloc is None else np.ascontiguousarray(reloc, T.RELOC_SEED)
    n = lib().orc_extract_frame_lost(width, height, int(frame_flags), _p(grey), _p(grid), _p(hops), _p(kps), len(kps),
                                     float(coverage_area), _p(prev), len(prev), _p(lk_status), _p(lk_pts), _p(reloc),
                                     0 if reloc is None else len(reloc), _p(ep), _p(cid), _p(out), _p(nb))
    return out[:max(n, 0)].copy(), prev, int(cid[0]), int(nb[0])


def frustum(pose, cam, width, height, cos_limit, pts):
    pts = np.ascontiguousarray(pts, T.MAP_POINT)
    out = np.zeros(len(pts), T.PROJECTION)
    pose = np.ascontiguousarray(pose, T.POSE)
    cam = np.ascontiguousarray(cam, T.CAMERA)
    lib().orc_frustum(_p(pose), _p(cam), width, height, cos_limit, _p(pts), len(pts), _p(out))
    return out


def search_by_video_feature(tracks, pts, proj, match, far_points=False, th_far=0.0):
    tracks = np.ascontiguousarray(tracks, T.TRACK)
    pts = np.ascontiguousarray(pts, T.MAP_POINT)
    proj = np.ascontiguousarray(proj, T.PROJECTION)
    match = np.array(match, np.int32, copy=True)
    n = lib().orc_search_by_video_feature(_p(tracks), len(tracks), _p(pts), _p(proj), len(pts), int(far_points),
                                          th_far, _p(match))
    return n, match


def search_by_keyframe(tracks, kf_pts):
    tracks = np.ascontiguousarray(tracks, T.TRACK)
    kf_pts = np.ascontiguousarray(kf_pts, T.MAP_POINT)
    match = np.zeros(len(tracks), np.int32)
    n = lib().orc_search_by_keyframe(_p(tracks), len(tracks), _p(kf_pts), len(kf_pts), _p(match))
    return n, match


def update_local_points(store, idx, n_first, capacity):
    store = np.ascontiguousarray(store, T.MAP_POINT)
    idx = np.ascontiguousarray(idx, np.int32)
    out = np.zeros(max(capacity, 1), T.MAP_POINT)
    nf = C.c_int32()
    n = lib().orc_update_local_points(_p(store), len(store), _p(idx), len(idx), int(n_first), _p(out), int(capacity), C.byref(nf))
    return out[:n], nf.value


def search_for_initialization(f1, f2, prev_matched):
    f1 = np.ascontiguousarray(f1, T.TRACK)
    f2 = np.ascontiguousarray(f2, T.TRACK)
    pm = np.array(prev_matched, np.float32, copy=True).reshape(len(f1), 2)
    m = np.zeros(len(f1), np.int32)
    n = lib().orc_search_for_initialization(_p(f1), len(f1), _p(f2), len(f2), _p(pm), _p(m))
    return n, m, pm


def assign_features_to_grid(tracks, width, height):
    tracks = np.ascontiguousarray(tracks, T.TRACK)
    start = np.zeros(64 * 48 + 1, np.int32)
    items = np.zeros(max(len(tracks), 1), np.int32)
    lib().orc_assign_features_to_grid(_p(tracks), len(tracks), width, height, _p(start), _p(items))
    return start, items[:start[-1]]


def get_features_in_area(tracks, width, height, start, items, x, y, r):
    tracks = np.ascontiguousarray(tracks, T.TRACK)
    items = np.ascontiguousarray(items, np.int32)
    out = np.zeros(max(len(tracks), 1), np.int32)
    n = lib().orc_get_features_in_area(_p(tracks), len(tracks), width, height, _p(start), _p(items), x, y, r, _p(out))
    return out[:n].copy()


def search_by_projection(feat, width, height, pts, proj, pt_desc, prm, taken=None):
    """Grid-bucketed search by projection for one frame -> (feat_match, pt_match, pt_dist, n_matches)."""
    feat = np.ascontiguousarray(feat, T.TRACK)
    pts = np.ascontiguousarray(pts, T.MAP_POINT)
    proj = np.ascontiguousarray(proj, T.PROJECTION)
    pt_desc = np.ascontiguousarray(pt_desc, np.uint32).reshape(-1, 8)
    prm = np.ascontiguousarray(prm, T.PROJECTION_SEARCH)
    if taken is not None:
        taken = np.ascontiguousarray(taken, np.uint8)
    fm = np.zeros(max(len(feat), 1), np.int32)
    pm = np.zeros(max(len(pts), 1), np.int32)
    pd = np.zeros(max(len(pts), 1), np.int32)
    n = lib().orc_search_by_projection(_p(feat), None if taken is None else _p(taken), len(feat), width, height, _p(pts), _p(proj),
                                       _p(pt_desc), len(pts), _p(prm), _p(fm), _p(pm), _p(pd))
    return fm[:len(feat)], pm[:len(pts)], pd[:len(pts)], n


def project(cam, Xc):
    cam = np.ascontiguousarray(cam, T.CAMERA)
    Xc = np.ascontiguousarray(Xc, np.float64)
    uv = np.zeros(2)
    lib().orc_project(_p(cam), _p(Xc), _p(uv))
    return uv


def project_jac(cam, Xc):
    cam = np.ascontiguousarray(cam, T.CAMERA)
    Xc = np.ascontiguousarray(Xc, np.float64)
    J = np.zeros(6)
    lib().orc_project_jac(_p(cam), _p(Xc), _p(J))
    return J.reshape(2, 3)


def pose_jacobian(cam, Xc):
    cam = np.ascontiguousarray(cam, T.CAMERA)
    Xc = np.ascontiguousarray(Xc, np.float64)
    J = np.zeros(12)
    lib().orc_pose_jacobian(_p(cam), _p(Xc), _p(J))
    return J.reshape(2, 6)


def huber_weight(chi2, delta):
    return lib().orc_huber_weight(chi2, delta)


def se3_exp(dx):
    dx = np.ascontiguousarray(dx, np.float64)
    R, t = np.zeros(9), np.zeros(3)
    lib().orc_se3_exp(_p(dx), _p(R), _p(t))
    return R.reshape(3, 3), t


def pose_optimize(cam, params, pts, obs, pose):
    """Returns (n_inliers, pose, outlier[n], stats[4])."""
    cam = np.ascontiguousarray(cam, T.CAMERA)
    params = np.ascontiguousarray(params, T.POSE_PARAMS)
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
    obs = np.ascontiguousarray(obs, np.float32).reshape(-1, 2)
    pose = np.array(pose, T.POSE, copy=True)
    outl = np.zeros(max(len(pts), 1), np.uint8)
    stats = np.zeros(4, np.int32)
    n = lib().orc_pose_optimize(_p(cam), _p(params), _p(pts), _p(obs), len(pts), _p(pose), _p(outl), _p(stats))
    return n, pose, outl[:len(pts)], stats


FRONTEND_CFG = np.dtype([("width", "<i4"), ("height", "<i4"), ("n_frames", "<i4"), ("max_ref", "<i4"),
                         ("max_tracks", "<i4"), ("threshold", "<i4"), ("coverage_threshold", "<f8"),
                         ("cam", T.CAMERA), ("_p0", "<i4"), ("pose_params", T.POSE_PARAMS), ("n_kf_points", "<i4"),
                         ("viewing_cos_limit", "<f4")])
assert FRONTEND_CFG.itemsize == 120, FRONTEND_CFG.itemsize


class _FrontendOut(C.Structure):
    _fields_ = [("poses", C.c_void_p), ("n_tracks", C.c_void_p), ("n_inliers", C.c_void_p),
                ("last_tracks", C.c_void_p), ("track_hash", C.c_void_p)]


def frontend_run(width, height, recs, rec_off, frame_flags, grey, seed_tracks, map_pts, pose0, cam, pose_params,
                 max_ref=3, max_tracks=4096, threshold=25, coverage_threshold=0.20, n_kf_points=0,
                 viewing_cos_limit=0.5, map_schedule=None, timed_from=0):
    """map_schedule: [(frame, MAP_POINT array, n_kf_points), ...] (ascending frames): the local map installed before that
    frame. The result carries tail_times = steady-clock seconds at the start of frame `timed_from` and at the end."""
    nf = len(frame_flags)
    cfg = np.zeros((), FRONTEND_CFG)
    cfg["width"], cfg["height"], cfg["n_frames"], cfg["max_ref"], cfg["max_tracks"] = width, height, nf, max_ref, max_tracks
    cfg["threshold"], cfg["coverage_threshold"] = threshold, coverage_threshold
    cfg["cam"], cfg["pose_params"] = cam, pose_params
    cfg["n_kf_points"], cfg["viewing_cos_limit"] = n_kf_points, viewing_cos_limit
    recs = np.ascontiguousarray(recs, T.MV_RECORD)
    rec_off = np.ascontiguousarray(rec_off, np.int64)
    frame_flags = np.ascontiguousarray(frame_flags, np.uint8)
    grey = None if grey is None else np.ascontiguousarray(grey, np.uint8)
    seeds = None if seed_tracks is None else np.ascontiguousarray(seed_tracks, T.TRACK)
    mp = np.ascontiguousarray(map_pts if map_pts is not None else np.zeros(0, T.MAP_POINT), T.MAP_POINT)
    pose0 = np.ascontiguousarray(pose0, T.POSE)
    poses = np.zeros(nf, T.POSE)
    n_tracks = np.zeros(nf, np.int32)
    n_inl = np.zeros(nf, np.int32)
    last = np.zeros(max_tracks, T.TRACK)
    hashes = np.zeros(nf, np.uint64)
    out = _FrontendOut(poses.ctypes.data, n_tracks.ctypes.data, n_inl.ctypes.data, last.ctypes.data, hashes.ctypes.data)
    sched = map_schedule or []
    s_frame = np.array([f for f, _, _ in sched], np.int32)
    s_nkf = np.array([k for _, _, k in sched], np.int32)
    s_off = np.cumsum([0] + [len(m) for _, m, _ in sched]).astype(np.int64)
    s_pts = np.ascontiguousarray(np.concatenate([np.ascontiguousarray(m, T.MAP_POINT) for _, m, _ in sched])) if sched else np.zeros(0, T.MAP_POINT)
    tail = np.zeros(2, np.float64)
    n_last = lib().orc_frontend_run_sched(_p(cfg), _p(recs), _p(rec_off), _p(frame_flags), _p(grey), _p(seeds),
                                          0 if seeds is None else len(seeds), _p(mp), len(mp), _p(pose0), len(sched), _p(s_frame),
                                          _p(s_off), _p(s_pts), _p(s_nkf), int(timed_from), _p(tail), C.byref(out))
    return dict(poses=poses, n_tracks=n_tracks, n_inliers=n_inl, last_tracks=last[:n_last].copy(), track_hash=hashes,
                tail_times=tail)


def table_checksum(tracks):
    """numpy mirror of frontend.cc::table_checksum (what frontend_run's track_hash holds per frame)."""
    w = np.ascontiguousarray(tracks, T.TRACK).view(np.uint64).ravel()
    i = np.arange(len(w), dtype=np.uint64)
    with np.errstate(over="ignore"):
        return int(np.sum((w ^ (i * np.uint64(0x9E3779B97F4A7C15))) * (np.uint64(2) * i + np.uint64(1)), dtype=np.uint64))
