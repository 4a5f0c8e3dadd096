"""numpy mirrors of include/movfe_types.h (layouts are asserted against the C side in tests)."""
import numpy as np

MV_RECORD = np.dtype({
    "names": ["source", "w", "h", "src_x", "src_y", "dst_x", "dst_y", "flags", "motion_x", "motion_y",
              "motion_scale", "ref"],
    "formats": ["<i4", "u1", "u1", "<i2", "<i2", "<i2", "<i2", "<u8", "<i4", "<i4", "<u2", "<i4"],
    "offsets": [0, 4, 5, 6, 8, 10, 12, 16, 24, 28, 32, 36],
    "itemsize": 40,
})
PACKED_RECORD = np.dtype([("src_x", "<i2"), ("src_y", "<i2"), ("dst_x", "<i2"), ("dst_y", "<i2"), ("w", "u1"), ("h", "u1"),
                          ("source_sign", "i1"), ("reserved", "u1"), ("ref", "<i4")])
HOP = np.dtype([("mv_x", "<f4"), ("mv_y", "<f4"), ("d_indx", "<i4"), ("_pad", "<i4")])
RECT = np.dtype([("x", "<i2"), ("y", "<i2"), ("w", "<i2"), ("h", "<i2")])
TRACK = np.dtype([("pt_x", "<f4"), ("pt_y", "<f4"), ("mb", RECT), ("track_id", "<i4"), ("age", "<i4"),
                  ("q_indx", "<i4"), ("flags", "<u4"), ("desc", "<u4", (8,))])
MAP_POINT = np.dtype([("pos", "<f4", (3,)), ("normal", "<f4", (3,)), ("min_dist", "<f4"), ("max_dist", "<f4"),
                      ("track_id", "<i4"), ("flags", "<u4")])
PROJECTION = np.dtype([("u", "<f4"), ("v", "<f4"), ("depth", "<f4"), ("view_cos", "<f4"), ("in_view", "<i4")])
CAMERA = np.dtype([("model", "<i4"), ("fx", "<f4"), ("fy", "<f4"), ("cx", "<f4"), ("cy", "<f4"),
                   ("k", "<f4", (4,))])
POSE = np.dtype([("R", "<f8", (9,)), ("t", "<f8", (3,))])
POSE_PARAMS = np.dtype([("is_lost", "<i4"), ("iteration_count", "<i4"), ("reprojection_error", "<f8"),
                        ("reprojection_error_lost", "<f8"), ("confidence", "<f8"), ("algorithm", "<i4"),
                        ("_pad", "<i4")])

RELOC_SEED = np.dtype([("track_id", "<i4"), ("q_indx", "<i4"), ("x", "<f4"), ("y", "<f4")])
assert RELOC_SEED.itemsize == 16
assert PACKED_RECORD.itemsize == 16 and MV_RECORD.itemsize == 40 and HOP.itemsize == 16 and RECT.itemsize == 8 and TRACK.itemsize == 64
assert MAP_POINT.itemsize == 40 and PROJECTION.itemsize == 20 and CAMERA.itemsize == 36
assert POSE.itemsize == 96 and POSE_PARAMS.itemsize == 40

FRAME_P = 0x1
FRAME_MV = 0x2
AREA_QUERY = np.dtype([("problem", "<i4"), ("x", "<f4"), ("y", "<f4"), ("r", "<f4")])
PROJECTION_SEARCH = np.dtype([("th", "<f4"), ("far_points", "<i4"), ("th_far", "<f4"), ("th_high", "<i4"), ("nn_ratio", "<f4")])
assert PROJECTION_SEARCH.itemsize == 20
TRACK_COVERAGE = 0x1
MP_BAD, MP_SKIP, MP_NULL = 0x1, 0x2, 0x4
CAM_PINHOLE, CAM_FISHEYE = 0, 1


def camera(fx, fy, cx, cy, k=(0, 0, 0, 0), model=CAM_PINHOLE):
    c = np.zeros((), CAMERA)
    c["model"], c["fx"], c["fy"], c["cx"], c["cy"], c["k"] = model, fx, fy, cx, cy, k
    return c


def pose(R=None, t=None):
    p = np.zeros((), POSE)
    p["R"] = np.eye(3).ravel() if R is None else np.asarray(R, np.float64).ravel()
    p["t"] = 0 if t is None else np.asarray(t, np.float64)
    return p


def pose_params(is_lost=False, iteration_count=50, reprojection_error=5.0, reprojection_error_lost=8.0,
                confidence=0.95, algorithm=38):
    p = np.zeros((), POSE_PARAMS)
    p["is_lost"], p["iteration_count"] = int(is_lost), iteration_count
    p["reprojection_error"], p["reprojection_error_lost"] = reprojection_error, reprojection_error_lost
    p["confidence"], p["algorithm"] = confidence, algorithm
    return p
