"""ctypes binding of libmovfe.so (the C-ABI in include/movfe.h).

Thin convenience layer for the tests and bench.py: numpy in, numpy out, every call goes through the C-ABI.
It never falls back to a CPU path: if the library or a CUDA device is missing, it raises.
"""
import ctypes as C
import os

import numpy as np

from . import types as T

_PKG = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # mov-slam_b200/
SO_PATH = os.environ.get("MOVFE_LIB") or os.path.join(_PKG, "lib", "libmovfe.so")   # MOVFE_LIB: development builds
_LIB = None


class MovfeError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("n_streams", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("max_records_per_frame", C.c_int32), ("max_ref", C.c_int32), ("window_frames", C.c_int32),
                ("max_tracks", C.c_int32), ("max_map_points", C.c_int32), ("express_threshold", C.c_int32),
                ("coverage_threshold", C.c_double), ("has_grey", C.c_int32), ("flags", C.c_int32)]


def load():
    """Loads libmovfe.so; raises if it has not been built (python __graft_entry__.py build)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(SO_PATH):
        raise MovfeError("libmovfe.so not built: run `make -C mov-slam_b200` (there is no CPU fallback)")
    L = C.CDLL(SO_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.movfe_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.movfe_destroy.argtypes = [vp]
    L.movfe_last_error.restype = C.c_char_p
    L.movfe_last_error.argtypes = [vp]
    L.movfe_version.restype = C.c_char_p
    L.movfe_synchronize.argtypes = [vp]
    L.movfe_fence.argtypes = [vp]
    L.movfe_cuda_stream.restype = vp
    L.movfe_cuda_stream.argtypes = [vp]
    L.movfe_push_frames.argtypes = [vp, i32, vp, vp, vp, vp]
    L.movfe_push_frames_device.argtypes = [vp, i32, vp, vp, i64, vp, vp]
    L.movfe_push_frames_packed.argtypes = [vp, i32, vp, vp, vp, vp, i32]
    L.movfe_pack_records.argtypes = [vp, i64, vp]
    L.movfe_pack_records.restype = None
    L.movfe_frames_pushed.restype = i64
    L.movfe_frames_pushed.argtypes = [vp]
    L.movfe_raster.argtypes = [vp, i64, i32]
    L.movfe_raster_counts.argtypes = [vp, i32, i64, vp, vp, vp]
    L.movfe_download_grid.argtypes = [vp, i32, i64, vp]
    L.movfe_download_hops.argtypes = [vp, i32, i64, vp, i32]
    L.movfe_download_kps.argtypes = [vp, i32, i64, vp, i32]
    L.movfe_rejected_records.restype = i64
    L.movfe_rejected_records.argtypes = [vp]
    L.movfe_set_tracks.argtypes = [vp, i32, vp, i32, i32]
    L.movfe_extract.argtypes = [vp, i64, i32]
    L.movfe_extract_frame.argtypes = [vp, C.c_uint32, vp, i32, vp, vp, i32, vp, i32, C.c_double, vp, i32, vp, vp, i32, vp, i32, vp, vp, i32]
    L.movfe_set_lk_results.argtypes = [vp, i32, vp, vp, i32, vp, i32]
    L.movfe_dropped_lk_tracks.restype = i64
    L.movfe_dropped_lk_tracks.argtypes = [vp]
    L.movfe_track_count.argtypes = [vp, i32, i64, vp, vp]
    L.movfe_download_tracks.argtypes = [vp, i32, i64, vp, i32]
    L.movfe_set_camera.argtypes = [vp, vp, vp, C.c_float]
    L.movfe_set_map_points.argtypes = [vp, i32, vp, i32, i32]
    L.movfe_set_pose.argtypes = [vp, i32, vp]
    L.movfe_lk_carry.argtypes = [vp, i64]
    L.movfe_lk.argtypes = [vp, i32, vp, vp, i32, vp, vp, i32, i32, i32, C.c_double, C.c_double, vp, vp, vp]
    L.movfe_reserve_map_store.argtypes = [vp, i32]
    L.movfe_set_map_store.argtypes = [vp, i32, i32, vp, i32]
    L.movfe_update_local_points.argtypes = [vp, vp, vp, vp]
    L.movfe_download_map_points.argtypes = [vp, i32, vp, i32, vp]
    L.movfe_set_map_points_batch.argtypes = [vp, vp, vp, vp, i32, i32]
    L.movfe_track_poses.argtypes = [vp, i64, i32]
    L.movfe_download_poses.argtypes = [vp, i64, i32, vp, vp]
    L.movfe_download_matches.argtypes = [vp, i32, i64, vp, vp, i32]
    L.movfe_frustum.argtypes = [vp, i32, vp, vp, vp, vp]
    L.movfe_join.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp]
    L.movfe_assign_features_to_grid.argtypes = [vp, i32, vp, vp, vp, vp]
    L.movfe_track_feature_grid.argtypes = [vp, i32, i64, vp, vp, i32]
    L.movfe_features_in_area.argtypes = [vp, i32, vp, vp, vp, vp, i32, vp, i32, vp, vp]
    L.movfe_search_by_projection.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.movfe_pose_optimize.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.movfe_workload_stats.argtypes = [vp, vp, i32]
    L.movfe_profile_enable.argtypes = [vp, i32]
    L.movfe_profile_read.argtypes = [vp, vp, vp, i32]
    _LIB = L
    return L


EXPORTS = ["movfe_create", "movfe_destroy", "movfe_last_error", "movfe_synchronize", "movfe_fence", "movfe_cuda_stream",
           "movfe_version", "movfe_push_frames", "movfe_push_frames_device", "movfe_push_frames_packed", "movfe_pack_records", "movfe_frames_pushed", "movfe_raster",
           "movfe_raster_counts", "movfe_download_grid", "movfe_download_hops", "movfe_download_kps",
           "movfe_rejected_records", "movfe_set_tracks", "movfe_set_lk_results", "movfe_dropped_lk_tracks", "movfe_extract", "movfe_extract_frame", "movfe_track_count",
           "movfe_download_tracks", "movfe_set_camera", "movfe_set_map_points", "movfe_set_map_points_batch", "movfe_reserve_map_store", "movfe_set_map_store", "movfe_update_local_points", "movfe_download_map_points", "movfe_set_pose",
           "movfe_track_poses", "movfe_download_poses", "movfe_download_matches", "movfe_frustum", "movfe_join",
           "movfe_assign_features_to_grid", "movfe_features_in_area", "movfe_track_feature_grid", "movfe_search_by_projection",
           "movfe_pose_optimize", "movfe_lk", "movfe_lk_carry", "movfe_profile_enable", "movfe_profile_read", "movfe_workload_stats"]


def pack_records(recs, out=None):
    """movfe_pack_records: 40-byte side-data records -> the 16-byte form movfe_push_frames_packed takes (host code)."""
    recs = np.ascontiguousarray(recs, T.MV_RECORD)
    if out is None:
        out = np.empty(len(recs), T.PACKED_RECORD)
    assert out.dtype == T.PACKED_RECORD and len(out) >= len(recs) and out.flags.c_contiguous
    load().movfe_pack_records(recs.ctypes.data_as(C.c_void_p), len(recs), out.ctypes.data_as(C.c_void_p))
    return out


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))        # raw device pointer


CFG_SERIAL_RASTER = 1   # MOVFE_CFG_SERIAL_RASTER
CFG_NO_GRID = 2         # MOVFE_CFG_NO_GRID


class Context:
    def __init__(self, n_streams, width, height, max_records_per_frame=4800, max_ref=3, window_frames=16,
                 max_tracks=4096, max_map_points=4096, express_threshold=25, coverage_threshold=0.20, has_grey=True,
                 device=0, serial_raster=False, output_grid=True):
        self.L = load()
        self.cfg = Config(device, n_streams, width, height, max_records_per_frame, max_ref, window_frames, max_tracks,
                          max_map_points, express_threshold, coverage_threshold, int(bool(has_grey)),
                          (CFG_SERIAL_RASTER if serial_raster else 0) | (0 if output_grid else CFG_NO_GRID))
        h = C.c_void_p()
        rc = self.L.movfe_create(C.byref(self.cfg), C.byref(h))
        if rc != 0:
            raise MovfeError("movfe_create failed (%d): %s" % (rc, self.L.movfe_last_error(None).decode()))
        self.h = h
        self.S, self.W, self.H = n_streams, width, height

    def close(self):
        if getattr(self, "h", None):
            self.L.movfe_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _ck(self, rc):
        if rc < 0:
            raise MovfeError("movfe error %d: %s" % (rc, self.L.movfe_last_error(self.h).decode()))
        return rc

    def synchronize(self):
        self._ck(self.L.movfe_synchronize(self.h))

    def fence(self):
        """Primary stream waits (on the device) for the pose stream."""
        self._ck(self.L.movfe_fence(self.h))

    @property
    def stream_ptr(self):
        return self.L.movfe_cuda_stream(self.h)

    STAGES = ("ingest", "hops", "grid", "extract", "pose")

    def workload_stats(self, reset=True):
        """Per-frame workload figures from the device counters (movfe_workload_stats)."""
        c = np.zeros(8, np.uint64)
        self._ck(self.L.movfe_workload_stats(self.h, _p(c), int(reset)))
        c = c.astype(np.float64)
        return {"tracks_looked_up": c[0], "candidates_per_track": c[1] / max(c[0], 1.0), "pose_solves": c[2],
                "pose_correspondences": c[3] / max(c[2], 1.0), "pose_passes_per_solve": c[4] / max(c[2], 1.0),
                "hops_per_frame": c[5] / max(c[6], 1.0), "frames_rastered": c[6]}

    def profile_enable(self, on=True):
        self._ck(self.L.movfe_profile_enable(self.h, int(on)))

    def profile_read(self, reset=True):
        """-> ({stage: ms}, {stage: kernel launches}) since the last reset (synchronises the stream)."""
        ms = np.zeros(len(self.STAGES), np.float64)
        ln = np.zeros(len(self.STAGES), np.int64)
        self._ck(self.L.movfe_profile_read(self.h, _p(ms), _p(ln), int(reset)))
        return dict(zip(self.STAGES, ms.tolist())), dict(zip(self.STAGES, ln.tolist()))

    # -- ingest ----------------------------------------------------------------------------------------------
    def push_frames(self, n_frames, recs, rec_off, frame_flags, grey=None):
        recs = np.ascontiguousarray(recs, T.MV_RECORD)
        rec_off = np.ascontiguousarray(rec_off, np.int64)
        frame_flags = np.ascontiguousarray(frame_flags, np.uint8)
        assert len(rec_off) == self.S * n_frames + 1 and len(frame_flags) == self.S * n_frames
        if grey is not None:
            grey = np.ascontiguousarray(grey, np.uint8)
            assert grey.size == self.S * n_frames * self.W * self.H
        self._ck(self.L.movfe_push_frames(self.h, n_frames, _p(recs), _p(rec_off), _p(frame_flags), _p(grey)))

    def push_frames_packed(self, n_frames, recs, rec_off, frame_flags, grey=None, grey_stride=0):
        """16-byte records (movfe_pack_records / pack_records below) instead of the 40-byte side-data form; grey_stride = bytes
        between the rows of a luma plane (0 = width)."""
        recs = np.ascontiguousarray(recs, T.PACKED_RECORD)
        rec_off = np.ascontiguousarray(rec_off, np.int64)
        frame_flags = np.ascontiguousarray(frame_flags, np.uint8)
        assert len(rec_off) == self.S * n_frames + 1 and len(frame_flags) == self.S * n_frames
        if grey is not None:
            grey = np.ascontiguousarray(grey, np.uint8)
            assert grey.size == self.S * n_frames * (grey_stride or self.W) * self.H
        self._ck(self.L.movfe_push_frames_packed(self.h, n_frames, _p(recs), _p(rec_off), _p(frame_flags), _p(grey), grey_stride))

    def push_frames_device(self, n_frames, d_recs, d_rec_off, n_records, d_flags, d_grey=None):
        self._ck(self.L.movfe_push_frames_device(self.h, n_frames, _p(d_recs), _p(d_rec_off), n_records, _p(d_flags),
                                                 _p(d_grey)))

    def frames_pushed(self):
        return self.L.movfe_frames_pushed(self.h)

    # -- raster ----------------------------------------------------------------------------------------------
    def raster(self, first_frame, n_out):
        self._ck(self.L.movfe_raster(self.h, first_frame, n_out))

    def raster_counts(self, stream, frame):
        nh, nk, cov = C.c_int32(), C.c_int32(), C.c_double()
        self._ck(self.L.movfe_raster_counts(self.h, stream, frame, C.byref(nh), C.byref(nk), C.byref(cov)))
        return nh.value, nk.value, cov.value

    def grid(self, stream, frame):
        out = np.empty((self.H, self.W, 4), np.int32)
        self._ck(self.L.movfe_download_grid(self.h, stream, frame, _p(out)))
        return out

    def hops(self, stream, frame):
        nh, _, _ = self.raster_counts(stream, frame)
        out = np.zeros(max(nh, 1), T.HOP)
        n = self._ck(self.L.movfe_download_hops(self.h, stream, frame, _p(out), len(out)))
        return out[:n]

    def kps(self, stream, frame):
        _, nk, _ = self.raster_counts(stream, frame)
        out = np.zeros(max(nk, 1), T.RECT)
        n = self._ck(self.L.movfe_download_kps(self.h, stream, frame, _p(out), len(out)))
        return out[:n]

    def rejected_records(self):
        return self.L.movfe_rejected_records(self.h)

    # -- propagation -------------------------------------------------------------------------------------------
    def set_tracks(self, stream, tracks, current_id):
        tracks = np.ascontiguousarray(tracks, T.TRACK)
        self._ck(self.L.movfe_set_tracks(self.h, stream, _p(tracks), len(tracks), current_id))

    def extract(self, first_frame, n_frames):
        self._ck(self.L.movfe_extract(self.h, first_frame, n_frames))

    def set_lk_results(self, stream, status=None, pts_xy=None, reloc=None):
        """Host LK results for the next frame propagated on `stream` (see movfe.h). status None: n = -1 (none)."""
        n = -1
        if status is not None:
            status = np.ascontiguousarray(status, np.uint8)
            pts_xy = np.ascontiguousarray(pts_xy, np.float32).reshape(-1, 2)
            n = len(status)
        reloc = None if reloc is None else np.ascontiguousarray(reloc, T.RELOC_SEED)
        self._ck(self.L.movfe_set_lk_results(self.h, stream, _p(status), _p(pts_xy), n, _p(reloc), 0 if reloc is None else len(reloc)))

    def dropped_lk_tracks(self):
        return self.L.movfe_dropped_lk_tracks(self.h)

    def extract_frame(self, frame_flags, grey, grid, hops, kps, coverage_area, prev, current_id, lk_status=None, lk_pts=None,
                      reloc=None):
        """Single-shot MOVExtractor::operator() on host raster results -> (tracks, current_id). grey may be a strided view
        (rows of grey.strides[0] bytes, as a cv::Mat ROI / AVFrame plane)."""
        grid = np.ascontiguousarray(grid, np.int32)
        hops = np.ascontiguousarray(hops, T.HOP)
        kps = np.ascontiguousarray(kps, T.RECT)
        prev = np.ascontiguousarray(prev, T.TRACK)
        stride = 0
        if grey is not None:
            grey = np.asarray(grey, np.uint8)
            if grey.ndim != 2 or grey.strides[1] != 1:
                grey = np.ascontiguousarray(grey)
            stride = grey.strides[0] if grey.ndim == 2 else 0
        n_lk = -1
        if lk_status is not None:
            lk_status = np.ascontiguousarray(lk_status, np.uint8)
            lk_pts = np.ascontiguousarray(lk_pts, np.float32).reshape(-1, 2)
            n_lk = len(lk_status)
        reloc = None if reloc is None else np.ascontiguousarray(reloc, T.RELOC_SEED)
        cid = C.c_int32(current_id)
        out = np.zeros(self.cfg.max_tracks, T.TRACK)
        n = self._ck(self.L.movfe_extract_frame(self.h, int(frame_flags), _p(grey), stride, _p(grid), _p(hops), len(hops), _p(kps), len(kps),
                                                float(coverage_area), _p(prev), len(prev), _p(lk_status), _p(lk_pts), n_lk, _p(reloc),
                                                0 if reloc is None else len(reloc), C.byref(cid), _p(out), len(out)))
        return out[:n], cid.value

    def track_count(self, stream, frame):
        n, cid = C.c_int32(), C.c_int32()
        self._ck(self.L.movfe_track_count(self.h, stream, frame, C.byref(n), C.byref(cid)))
        return n.value, cid.value

    def tracks(self, stream, frame):
        n, _ = self.track_count(stream, frame)
        out = np.zeros(max(n, 1), T.TRACK)
        n = self._ck(self.L.movfe_download_tracks(self.h, stream, frame, _p(out), len(out)))
        return out[:n]

    # -- match / pose --------------------------------------------------------------------------------------------
    def set_camera(self, cam, pose_params, viewing_cos_limit=0.5):
        cam = np.ascontiguousarray(cam, T.CAMERA)
        pp = np.ascontiguousarray(pose_params, T.POSE_PARAMS)
        self._ck(self.L.movfe_set_camera(self.h, _p(cam), _p(pp), viewing_cos_limit))

    def set_map_points(self, stream, pts, n_keyframe_points):
        pts = np.ascontiguousarray(pts, T.MAP_POINT)
        self._ck(self.L.movfe_set_map_points(self.h, stream, _p(pts), len(pts), n_keyframe_points))

    def set_map_points_batch(self, pts, off, n_kf, max_points_per_stream, on_device=False):
        """Local maps of all streams at once; host numpy arrays, or raw device pointers with on_device=True."""
        if not on_device:
            pts = np.ascontiguousarray(pts, T.MAP_POINT)
            off = np.ascontiguousarray(off, np.int64)
            n_kf = np.ascontiguousarray(n_kf, np.int32)
            assert len(off) == self.S + 1 and len(n_kf) == self.S
            self._keep_map = (pts, off, n_kf)       # host arrays must stay unchanged until the next synchronising call
        self._ck(self.L.movfe_set_map_points_batch(self.h, _p(pts), _p(off), _p(n_kf), int(max_points_per_stream), int(on_device)))

    # -- device-side Tracking::UpdateLocalPoints --------------------------------------------------------------------
    def reserve_map_store(self, max_points_per_stream):
        self._ck(self.L.movfe_reserve_map_store(self.h, int(max_points_per_stream)))

    def set_map_store(self, stream, first_index, pts):
        pts = np.ascontiguousarray(pts, T.MAP_POINT)
        self._ck(self.L.movfe_set_map_store(self.h, stream, int(first_index), _p(pts), len(pts)))

    def update_local_points(self, idx, off, n_kf_entries):
        idx = np.ascontiguousarray(idx, np.int32)
        off = np.ascontiguousarray(off, np.int64)
        n_kf_entries = np.ascontiguousarray(n_kf_entries, np.int32)
        assert len(off) == self.S + 1 and len(n_kf_entries) == self.S and len(idx) == off[-1]
        self._ck(self.L.movfe_update_local_points(self.h, _p(idx), _p(off), _p(n_kf_entries)))

    def map_points(self, stream, capacity=1 << 16):
        out = np.zeros(capacity, T.MAP_POINT)
        nk = C.c_int32()
        n = self.L.movfe_download_map_points(self.h, stream, _p(out), capacity, C.byref(nk))
        self._ck(n if n < 0 else 0)
        return out[:n], nk.value

    def lk_carry(self, frame):
        """device-resident LK results for `frame` of every stream (call between extract(.., frame - 1) and extract(frame, ..))"""
        self._ck(self.L.movfe_lk_carry(self.h, frame))

    def lk(self, prev, nxt, pts, off, win=31, max_level=3, max_count=20, eps=0.01, min_eig=1e-4):
        """cv::calcOpticalFlowPyrLK for len(off) - 1 image pairs of the context's size. -> (next points, status, err)."""
        prev = np.ascontiguousarray(prev, np.uint8).reshape(-1, self.H, self.W)
        nxt = np.ascontiguousarray(nxt, np.uint8).reshape(-1, self.H, self.W)
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        off = np.ascontiguousarray(off, np.int32)
        assert len(prev) == len(nxt) == len(off) - 1 and off[-1] == len(pts)
        out = np.zeros((len(pts), 2), np.float32)
        st = np.zeros(len(pts), np.uint8)
        err = np.zeros(len(pts), np.float32)
        self._ck(self.L.movfe_lk(self.h, len(prev), _p(prev), _p(nxt), 0, _p(pts), _p(off), win, max_level, max_count, eps, min_eig, _p(out), _p(st), _p(err)))
        return out, st, err

    def set_pose(self, stream, pose):
        pose = np.ascontiguousarray(pose, T.POSE)
        self._ck(self.L.movfe_set_pose(self.h, stream, _p(pose)))

    def track_poses(self, first_frame, n_frames):
        self._ck(self.L.movfe_track_poses(self.h, first_frame, n_frames))

    def poses(self, first_frame, n_frames):
        poses = np.zeros((self.S, n_frames), T.POSE)
        ninl = np.zeros((self.S, n_frames), np.int32)
        self._ck(self.L.movfe_download_poses(self.h, first_frame, n_frames, _p(poses), _p(ninl)))
        return poses, ninl

    def matches(self, stream, frame):
        cap = self.cfg.max_tracks
        m = np.zeros(cap, np.int32)
        o = np.zeros(cap, np.uint8)
        n = self._ck(self.L.movfe_download_matches(self.h, stream, frame, _p(m), _p(o), cap))
        return m[:n], o[:n]

    # -- single-shot operators -------------------------------------------------------------------------------------
    def frustum(self, poses, pts, off):
        poses = np.ascontiguousarray(poses, T.POSE)
        pts = np.ascontiguousarray(pts, T.MAP_POINT)
        off = np.ascontiguousarray(off, np.int32)
        out = np.zeros(len(pts), T.PROJECTION)
        self._ck(self.L.movfe_frustum(self.h, len(off) - 1, _p(poses), _p(pts), _p(off), _p(out)))
        return out

    def join(self, track_ids, track_off, probe_ids, probe_valid, probe_off, match_init):
        track_ids = np.ascontiguousarray(track_ids, np.int32)
        track_off = np.ascontiguousarray(track_off, np.int32)
        probe_ids = np.ascontiguousarray(probe_ids, np.int32)
        probe_valid = np.ascontiguousarray(probe_valid, np.uint8)
        probe_off = np.ascontiguousarray(probe_off, np.int32)
        match = np.array(match_init, np.int32, copy=True)
        n = np.zeros(len(track_off) - 1, np.int32)
        self._ck(self.L.movfe_join(self.h, len(track_off) - 1, _p(track_ids), _p(track_off), _p(probe_ids),
                                   _p(probe_valid), _p(probe_off), _p(match), _p(n)))
        return match, n

    def assign_features_to_grid(self, pts_xy, off):
        """Frame::AssignFeaturesToGrid for packed keypoint sets -> (cell_start [n_sets, 64*48+1], cell_items [n])."""
        pts_xy = np.ascontiguousarray(pts_xy, np.float32).reshape(-1, 2)
        off = np.ascontiguousarray(off, np.int32)
        start = np.zeros((len(off) - 1, 64 * 48 + 1), np.int32)
        items = np.full(max(len(pts_xy), 1), -1, np.int32)
        self._ck(self.L.movfe_assign_features_to_grid(self.h, len(off) - 1, _p(pts_xy), _p(off), _p(start), _p(items)))
        return start, items[:len(pts_xy)]

    def track_feature_grid(self, stream, frame):
        """Bucket grid of a resident track table -> (cell_start [64*48+1], cell_items [n_tracks])."""
        start = np.zeros(64 * 48 + 1, np.int32)
        items = np.full(self.cfg.max_tracks, -1, np.int32)
        n = self._ck(self.L.movfe_track_feature_grid(self.h, stream, frame, _p(start), _p(items), len(items)))
        return start, items[:n]

    def features_in_area(self, pts_xy, off, start, items, queries, capacity):
        """Frame::GetFeaturesInArea for a batch of (set, x, y, r) queries -> (indices [n_queries, capacity], counts)."""
        pts_xy = np.ascontiguousarray(pts_xy, np.float32).reshape(-1, 2)
        off = np.ascontiguousarray(off, np.int32)
        queries = np.ascontiguousarray(queries, T.AREA_QUERY)
        out = np.full((len(queries), max(capacity, 1)), -1, np.int32)
        counts = np.zeros(max(len(queries), 1), np.int32)
        self._ck(self.L.movfe_features_in_area(self.h, len(off) - 1, _p(pts_xy), _p(off), _p(np.ascontiguousarray(start, np.int32)),
                                               _p(np.ascontiguousarray(items, np.int32)), len(queries), _p(queries), capacity,
                                               _p(out), _p(counts)))
        return out[:, :capacity], counts[:len(queries)]

    def search_by_projection(self, feat, feat_off, pts, proj, pt_desc, pt_off, prm, taken=None):
        """Grid-bucketed search by projection for a batch of frames -> (feat_match, pt_match, pt_dist, n_matches)."""
        feat = np.ascontiguousarray(feat, T.TRACK)
        feat_off = np.ascontiguousarray(feat_off, np.int32)
        pts = np.ascontiguousarray(pts, T.MAP_POINT)
        proj = np.ascontiguousarray(proj, T.PROJECTION)
        pt_desc = np.ascontiguousarray(pt_desc, np.uint32).reshape(-1, 8)
        pt_off = np.ascontiguousarray(pt_off, np.int32)
        prm = np.ascontiguousarray(prm, T.PROJECTION_SEARCH)
        assert len(proj) == len(pts) == len(pt_desc)
        if taken is not None:
            taken = np.ascontiguousarray(taken, np.uint8)
            assert len(taken) == len(feat)
        fm = np.zeros(max(len(feat), 1), np.int32)
        pm = np.zeros(max(len(pts), 1), np.int32)
        pd = np.zeros(max(len(pts), 1), np.int32)
        nm = np.zeros(len(feat_off) - 1, np.int32)
        self._ck(self.L.movfe_search_by_projection(self.h, len(feat_off) - 1, _p(feat), None if taken is None else _p(taken), _p(feat_off),
                                                   _p(pts), _p(proj), _p(pt_desc), _p(pt_off), _p(prm), _p(fm), _p(pm), _p(pd), _p(nm)))
        return fm[:len(feat)], pm[:len(pts)], pd[:len(pts)], nm

    def pose_optimize(self, cam, pp, pts, obs, off, poses):
        cam = np.ascontiguousarray(cam, T.CAMERA)
        pp = np.ascontiguousarray(pp, T.POSE_PARAMS)
        pts = np.ascontiguousarray(pts, np.float32)
        obs = np.ascontiguousarray(obs, np.float32)
        off = np.ascontiguousarray(off, np.int32)
        poses = np.array(poses, T.POSE, copy=True)
        n = len(off) - 1
        outl = np.zeros(max(int(off[-1]), 1), np.uint8)
        ninl = np.zeros(n, np.int32)
        stats = np.zeros((n, 4), np.int32)
        self._ck(self.L.movfe_pose_optimize(self.h, n, _p(cam), _p(pp), _p(pts), _p(obs), _p(off), _p(poses), _p(outl),
                                            _p(ninl), _p(stats)))
        return poses, outl[:int(off[-1])], ninl, stats
