/*
 * movfe_types.h — plain-old-data records shared by the C-ABI (include/movfe.h), the CUDA library
 * and the CPU oracle (oracle/). No CUDA, torch or C++ types appear here.
 *
 * Every type cites the reference structure it stands in for (paths relative to the MoV-SLAM tree).
 */
#ifndef MOVFE_TYPES_H
#define MOVFE_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One libavcodec motion-vector side-data record, patched with the reference index.
 * Layout == AVMotionVector of FFmpeg 4.4.3 + ffmpeg-ref-patch.patch:122-129 (sizeof == 40).
 * The hot path reads source, w, h, src_*, dst_*, ref only (src/VideoDecoder.cc:211-228);
 * flags/motion_x/motion_y/motion_scale are carried but never change a result. */
typedef struct movfe_mv_record {
    int32_t  source;        /* <0: past frame (P), >0: future frame (B), 0 treated like P by the reference */
    uint8_t  w, h;          /* block size: 16/8 from ffmpeg (patch:21-22); 4 only in synthetic stress input */
    int16_t  src_x, src_y;  /* block centre in the source frame */
    int16_t  dst_x, dst_y;  /* block centre in the current frame */
    uint16_t _pad0;
    uint64_t flags;
    int32_t  motion_x, motion_y;
    uint16_t motion_scale;
    uint16_t _pad1;
    int32_t  ref;           /* reference index: the source frame is ref+1 frames back */
} movfe_mv_record;

/* The seven fields of a record that the path reads (VideoDecoder.cc:211-228), one 128-bit word: what the device keeps of a
 * record, and what a decoder shim may hand over instead of the 40-byte form (movfe_pack_records, movfe_push_frames_packed):
 * 60 % fewer record bytes over PCIe. */
typedef struct movfe_packed_record {
    int16_t src_x, src_y, dst_x, dst_y;
    uint8_t w, h;
    int8_t  source_sign;   /* sign of AVMotionVector::source: -1, 0, +1 */
    uint8_t reserved;      /* 0 */
    int32_t ref;
} movfe_packed_record;


/* Per-frame flags (VideoDecoder.cc:193,200): */
#define MOVFE_FRAME_P        0x1u  /* pict_type != I  -> FrameType::P_FRAME */
#define MOVFE_FRAME_MV       0x2u  /* NextImage(mv=true) and side data present: records are consumed */

/* One per-frame hop of a record: MotionVector{pt, dIndx} (include/Frame.h:55-77).
 * 16 bytes so a hop is one aligned 128-bit load; the 4th word is padding (always 0). */
typedef struct movfe_hop {
    float   mv_x, mv_y;     /* per-frame displacement (dst-src)/(ref+1), VideoDecoder.cc:220-224 */
    int32_t d_indx;         /* index into the frame's kps list, -1 if none (VideoDecoder.cc:243-253) */
    int32_t _pad;
} movfe_hop;

/* cv::Rect with 16-bit fields: candidate-keypoint block (VideoImage::kps, Frame.h:115) and track block. */
typedef struct movfe_rect {
    int16_t x, y, w, h;
} movfe_rect;

/* VideoFeature (include/Frame.h:79-107) as a 64-byte record.
 * VideoFeature::dIndx (index of the matching cv::KeyPoint) always equals the record's own index in the
 * frame's table (MOVExtractor.cc:318-331), and the KeyPoint is (pt, size = mb.w), so neither is stored. */
typedef struct movfe_track {
    float      pt_x, pt_y;  /* VideoFeature::pt == KeyPoint::pt */
    movfe_rect mb;          /* VideoFeature::mb */
    int32_t    track_id;    /* VideoFeature::trackId, join key to MapPoint::mTrackId (MapPoint.h:175) */
    int32_t    age;
    int32_t    q_indx;      /* index in the (sorted) previous table, -1 for births */
    uint32_t   flags;       /* bit0: VideoFeature::coverage */
    uint32_t   desc[8];     /* bitset<256>, bit i at desc[i>>5] bit (i&31) */
} movfe_track;
#define MOVFE_TRACK_COVERAGE 0x1u

/* One lost-relocalisation seed (src/MOVExtractor.cc:199-241): a reference-keyframe map point that the host carried into
 * the current image with cv::calcOpticalFlowPyrLK and that passed the status / image-bounds / distance tests (:207-215).
 * The device adds the 16x16 block around it, its bounds test and its descriptor (:218-238). 16 bytes. */
typedef struct movfe_reloc_seed {
    int32_t track_id;       /* MapPoint::mTrackId (:191) */
    int32_t q_indx;         /* index in the point list handed to LK (:230) */
    float   x, y;           /* ptR (:205) */
} movfe_reloc_seed;

/* A local map point as the front-end reads it (MapPoint::GetWorldPos/GetNormal/mfMin/MaxDistance/
 * mTrackId/isBad; Frame.cc:456-519, MOVMatcher.h:35-68). 40 bytes. */
typedef struct movfe_map_point {
    float    pos[3];        /* world position (float, MapPoint.h:179) */
    float    normal[3];     /* mean viewing direction */
    float    min_dist, max_dist; /* mfMinDistance, mfMaxDistance; the 0.8/1.2 factors are applied on read */
    int32_t  track_id;      /* MapPoint::mTrackId */
    uint32_t flags;         /* MOVFE_MP_* */
} movfe_map_point;
#define MOVFE_MP_BAD        0x1u  /* MapPoint::isBad() */
#define MOVFE_MP_SKIP       0x2u  /* mnLastFrameSeen == current frame: already matched, not re-projected (Tracking.cc:1136) */
#define MOVFE_MP_NULL       0x4u  /* null entry of a keyframe's map-point list (MOVMatcher.h:82) */

/* Result of Frame::isInFrustum for one map point (Frame.cc:505-516). */
typedef struct movfe_projection {
    float    u, v;          /* mTrackProjX/Y (-1,-1 when rejected before the bounds test passes) */
    float    depth;         /* mTrackDepth = |Pc| */
    float    view_cos;      /* mTrackViewCos */
    int32_t  in_view;       /* mbTrackInView */
} movfe_projection;

/* One Frame::GetFeaturesInArea(x, y, r) call (src/Frame.cc:602; minLevel = 0, maxLevel = -1) against keypoint set `problem`. */
typedef struct movfe_area_query {
    int32_t problem;
    float   x, y, r;
} movfe_area_query;

/* Parameters of the grid-bucketed search by projection (movfe_search_by_projection). The names are those of the ORB-SLAM3-lineage
 * matcher MoV-SLAM descends from (ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th, bFarPoints, thFarPoints),
 * TH_HIGH, mfNNratio); distances are EXPRESS Hamming distances, 0..256 (include/EXPRESS.h:112-115). 20 bytes. */
typedef struct movfe_projection_search_params {
    float   th;          /* search radius = th * (view_cos > 0.998 ? 2.5 : 4.0) pixels (RadiusByViewingCos; one pyramid level) */
    int32_t far_points;  /* bFarPoints: skip points whose depth exceeds th_far */
    float   th_far;      /* thFarPoints */
    int32_t th_high;     /* a best distance above this is no match */
    float   nn_ratio;    /* with a second candidate: no match when best > nn_ratio * second best */
} movfe_projection_search_params;

/* Camera: GeometricCamera::mvParameters = [fx,fy,cx,cy,(k1..k4)] (GeometricCamera.h:61-101). */
#define MOVFE_CAM_PINHOLE 0
#define MOVFE_CAM_FISHEYE 1   /* KannalaBrandt8: not in the reference tree; ORB-SLAM3 lineage formulae */
typedef struct movfe_camera {
    int32_t model;
    float   fx, fy, cx, cy;
    float   k[4];
} movfe_camera;

/* Rigid pose T_cw, row-major rotation + translation, double (Sophus::SE3f is widened on entry). */
typedef struct movfe_pose {
    double R[9];
    double t[3];
} movfe_pose;

/* Parameters of Optimizer::PoseOptimization (include/Optimizer.h:55). */
typedef struct movfe_pose_params {
    int32_t is_lost;
    int32_t iteration_count;          /* total Gauss-Newton budget, split over 4 rounds */
    double  reprojection_error;
    double  reprojection_error_lost;
    double  confidence;               /* accepted, unused by the GN/Huber solver */
    int32_t algorithm;                /* accepted, unused */
    int32_t _pad;
} movfe_pose_params;

#ifdef __cplusplus
}
#endif
#endif /* MOVFE_TYPES_H */
