"""ctypes binding of oracle/_ref/libmovref{,_canon}.so — TEST INFRASTRUCTURE ONLY.

Those libraries are the REFERENCE'S OWN front-end sources (src/VideoDecoder.cc, src/MOVExtractor.cc, include/EXPRESS.h,
include/MOVMatcher.h with include/Frame.h), compiled unmodified by `make -C oracle ref` against the stand-in headers of
oracle/ref_standin/. They exist to pin the oracle (tests/test_ref_parity.py) and to generate the golden fixtures under
tests/golden/; nothing under mov-slam_b200/ may load them.

variant "plain": the reference as written (std::sort of prev->mvVF, ties implementation-defined).
variant "canon": the same sources with that one sort call bound to a stable sort (oracle/ref_standin/sort_canon.h).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(_HERE), "mov-slam_b200", "python"))
from movfe import types as T  # noqa: E402

_LIBS = {}
REFERENCE_ROOT = os.environ.get("MOVFE_REFERENCE_ROOT", "/root/reference")


def so_path(variant="plain"):
    return os.path.join(_HERE, "_ref", "libmovref.so" if variant == "plain" else "libmovref_canon.so")


def available(variant="plain"):
    """True when the prebuilt library is there (it travels to the GPU box) or can be built (reference tree present)."""
    return os.path.exists(so_path(variant)) or os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def build():
    subprocess.run(["make", "-C", _HERE, "-s", "ref", "REF=" + REFERENCE_ROOT], check=True)


class AUX(C.Structure):
    _fields_ = [("n_keypoints", C.c_int32), ("n_descriptors", C.c_int32), ("consistent", C.c_int32), ("lk_calls", C.c_int32)]


def lib(variant="plain"):
    if variant not in _LIBS:
        so = so_path(variant)
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        vp, i32, f32, f64, u32 = C.c_void_p, C.c_int32, C.c_float, C.c_double, C.c_uint32
        L.ref_last_error.restype = C.c_char_p
        L.ref_decode_clip.restype = vp
        L.ref_decode_clip.argtypes = [i32, i32, i32, vp, vp, vp, vp, i32]
        L.ref_clip_free.argtypes = [vp]
        for name in ("ref_clip_n_hops", "ref_clip_n_kps", "ref_clip_frame_no", "ref_clip_is_p"):
            getattr(L, name).restype = i32
            getattr(L, name).argtypes = [vp, i32]
        L.ref_clip_coverage.restype = f64
        L.ref_clip_coverage.argtypes = [vp, i32]
        for name in ("ref_clip_grid", "ref_clip_hops", "ref_clip_kps", "ref_clip_grey"):
            getattr(L, name).restype = vp
            getattr(L, name).argtypes = [vp, i32]
        L.ref_express_center.restype = i32
        L.ref_express_center.argtypes = [vp, i32, i32, i32, i32, i32]
        L.ref_express_descriptor.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp]
        L.ref_express_test.restype = i32
        L.ref_express_test.argtypes = [vp, i32, i32, i32, i32, i32, i32]
        L.ref_express_distance.restype = i32
        L.ref_express_distance.argtypes = [vp, vp]
        L.ref_lk_push.argtypes = [vp, vp, i32]
        L.ref_lk_calls.restype = i32
        L.ref_lk_last_points.restype = i32
        L.ref_lk_last_points.argtypes = [vp, i32]
        L.ref_extract_frame.restype = i32
        L.ref_extract_frame.argtypes = [i32, i32, u32, vp, vp, vp, i32, vp, i32, f64, vp, i32, i32, i32, i32, vp, vp, vp, i32, f64, f64,
                                        vp, vp, i32, vp]
        L.ref_frontend_run.restype = i32
        L.ref_frontend_run.argtypes = [i32, i32, i32, vp, vp, vp, vp, i32, i32, f64, vp, vp, vp, vp]
        L.ref_search_by_video_feature.restype = i32
        L.ref_search_by_video_feature.argtypes = [vp, i32, vp, vp, i32, i32, f32, vp]
        L.ref_search_by_keyframe.restype = i32
        L.ref_search_by_keyframe.argtypes = [vp, i32, vp, i32, vp]
        L.ref_search_for_initialization.restype = i32
        L.ref_search_for_initialization.argtypes = [vp, i32, vp, i32, vp, vp]
        _LIBS[variant] = L
    return _LIBS[variant]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Clip:
    """MOV_SLAM::VideoDecoder (qlen frames of look-ahead) run over one stream's synthetic clip through a fake libav."""

    def __init__(self, width, height, recs, rec_off, frame_flags, grey=None, qlen=12):
        self.W, self.H = width, height
        self.n_frames = len(frame_flags)
        recs = np.ascontiguousarray(recs, T.MV_RECORD)
        rec_off = np.ascontiguousarray(rec_off, np.int64)
        frame_flags = np.ascontiguousarray(frame_flags, np.uint8)
        grey = None if grey is None else np.ascontiguousarray(grey, np.uint8)
        assert len(rec_off) == self.n_frames + 1
        self._keep = (recs, rec_off, frame_flags, grey)
        self._h = lib().ref_decode_clip(width, height, self.n_frames, _p(recs), _p(rec_off), _p(frame_flags), _p(grey), qlen)
        if not self._h:
            raise ValueError(lib().ref_last_error().decode())

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_clip_free(self._h)
            self._h = None

    def _arr(self, ptr, n, dtype):
        if n == 0 or not ptr:
            return np.zeros(0, dtype)
        buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype, n).copy()

    def n_hops(self, f):
        return lib().ref_clip_n_hops(self._h, f)

    def n_kps(self, f):
        return lib().ref_clip_n_kps(self._h, f)

    def coverage(self, f):
        return lib().ref_clip_coverage(self._h, f)

    def frame_no(self, f):
        return lib().ref_clip_frame_no(self._h, f)

    def is_p(self, f):
        return bool(lib().ref_clip_is_p(self._h, f))

    def grid(self, f):
        return self._arr(lib().ref_clip_grid(self._h, f), self.W * self.H * 4, np.int32).reshape(self.H, self.W, 4)

    def grey(self, f):
        return self._arr(lib().ref_clip_grey(self._h, f), self.W * self.H, np.uint8).reshape(self.H, self.W)

    def hops(self, f):
        return self._arr(lib().ref_clip_hops(self._h, f), self.n_hops(f), T.HOP)

    def kps(self, f):
        return self._arr(lib().ref_clip_kps(self._h, f), self.n_kps(f), T.RECT)


def express_descriptor(img, x0, y0, cols, rows, thr):
    img = np.ascontiguousarray(img, np.uint8)
    d = np.zeros(8, np.uint32)
    lib().ref_express_descriptor(_p(img), img.shape[1], x0, y0, cols, rows, thr, _p(d))
    return d


def express_center(img, x0, y0, cols, rows):
    img = np.ascontiguousarray(img, np.uint8)
    return lib().ref_express_center(_p(img), img.shape[1], x0, y0, cols, rows)


def express_test(img, x0, y0, cols, rows, thr):
    img = np.ascontiguousarray(img, np.uint8)
    return bool(lib().ref_express_test(_p(img), img.shape[1], x0, y0, cols, rows, thr))


def express_distance(a, b):
    a = np.ascontiguousarray(a, np.uint32)
    b = np.ascontiguousarray(b, np.uint32)
    return lib().ref_express_distance(_p(a), _p(b))


def extract_frame(width, height, frame_flags, grey, grid, hops, kps, coverage_area, prev, current_id, threshold=25,
                  coverage_threshold=0.20, relocalization_distance=0.25, lk_calls=(), lost=False, kf_points=None,
                  has_prev=True, variant="canon", capacity=1 << 16):
    """One MOVExtractor::operator() call of the reference. lk_calls: [(status, pts_xy), ...] handed to the successive
    cv::calcOpticalFlowPyrLK calls the reference makes. kf_points (lost mode): (in_view, proj_xy, track_id).
    Returns dict(tracks, sorted_prev, current_id, n_keypoints, n_descriptors, consistent, lk_calls, lk_last_points)."""
    L = lib(variant)
    grey = None if grey is None else np.ascontiguousarray(grey, np.uint8)
    grid = np.ascontiguousarray(grid, np.int32)
    hops = np.ascontiguousarray(hops, T.HOP)
    kps = np.ascontiguousarray(kps, T.RECT)
    prev = np.array(prev, T.TRACK, copy=True)
    out = np.zeros(capacity, T.TRACK)
    cid = np.array([current_id], np.int32)
    aux = AUX()
    L.ref_lk_reset()
    for st, pts in lk_calls:
        st = np.ascontiguousarray(st, np.uint8)
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        L.ref_lk_push(_p(st), _p(pts), len(st))
    if kf_points is not None:
        kv = np.ascontiguousarray(kf_points[0], np.uint8)
        kxy = np.ascontiguousarray(kf_points[1], np.float32).reshape(-1, 2)
        kid = np.ascontiguousarray(kf_points[2], np.int32)
        n_kf = len(kv)
    else:
        kv = kxy = kid = None
        n_kf = 0
    n = L.ref_extract_frame(width, height, int(frame_flags), _p(grey), _p(grid), _p(hops), len(hops), _p(kps), len(kps),
                            float(coverage_area), _p(prev), len(prev), int(has_prev), int(lost), n_kf, _p(kv), _p(kxy), _p(kid),
                            int(threshold), float(coverage_threshold), float(relocalization_distance), _p(cid), _p(out), capacity,
                            C.byref(aux))
    last = np.zeros((max(len(prev), n_kf, 1), 2), np.float32)
    n_last = L.ref_lk_last_points(_p(last), len(last))
    return dict(tracks=out[:max(min(n, capacity), 0)].copy(), sorted_prev=prev, current_id=int(cid[0]), n_keypoints=aux.n_keypoints,
                n_descriptors=aux.n_descriptors, consistent=bool(aux.consistent), lk_calls=aux.lk_calls,
                lk_last_points=last[:min(n_last, len(last))].copy())


def search_by_video_feature(tracks, pts, proj, match, far_points=False, th_far=0.0):
    tracks = np.ascontiguousarray(tracks, T.TRACK)
    pts = np.ascontiguousarray(pts, T.MAP_POINT)
    proj = np.ascontiguousarray(proj, T.PROJECTION)
    match = np.array(match, np.int32, copy=True)
    n = lib().ref_search_by_video_feature(_p(tracks), len(tracks), _p(pts), _p(proj), len(pts), int(far_points), th_far, _p(match))
    return n, match


def search_by_keyframe(tracks, kf_pts):
    tracks = np.ascontiguousarray(tracks, T.TRACK)
    kf_pts = np.ascontiguousarray(kf_pts, T.MAP_POINT)
    match = np.zeros(len(tracks), np.int32)
    n = lib().ref_search_by_keyframe(_p(tracks), len(tracks), _p(kf_pts), len(kf_pts), _p(match))
    return n, match


def search_for_initialization(f1, f2, prev_matched):
    f1 = np.ascontiguousarray(f1, T.TRACK)
    f2 = np.ascontiguousarray(f2, T.TRACK)
    pm = np.array(prev_matched, np.float32, copy=True).reshape(len(f1), 2)
    m = np.zeros(len(f1), np.int32)
    n = lib().ref_search_for_initialization(_p(f1), len(f1), _p(f2), len(f2), _p(pm), _p(m))
    return n, m, pm


def frontend_run(width, height, recs, rec_off, frame_flags, grey, qlen=12, threshold=25, coverage_threshold=0.20, variant="canon"):
    """The reference's own decoder + extractor loop over a clip (LK loses every point). -> dict(n_tracks, track_hash,
    seconds_decoder, seconds_extractor)."""
    nf = len(frame_flags)
    recs = np.ascontiguousarray(recs, T.MV_RECORD)
    rec_off = np.ascontiguousarray(rec_off, np.int64)
    frame_flags = np.ascontiguousarray(frame_flags, np.uint8)
    grey = None if grey is None else np.ascontiguousarray(grey, np.uint8)
    nt = np.zeros(nf, np.int32)
    hs = np.zeros(nf, np.uint64)
    td, te = C.c_double(), C.c_double()
    n = lib(variant).ref_frontend_run(width, height, nf, _p(recs), _p(rec_off), _p(frame_flags), _p(grey), qlen, threshold,
                                      float(coverage_threshold), C.byref(td), C.byref(te), _p(nt), _p(hs))
    assert n == nf, (n, nf)
    return dict(n_tracks=nt, track_hash=hs, seconds_decoder=td.value, seconds_extractor=te.value)
