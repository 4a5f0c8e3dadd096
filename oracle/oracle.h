/*
 * oracle.h — C API of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY. The oracle is a CPU restatement of the reference's algorithm for the
 * tracking front-end; it is the checker for the CUDA path and the timed CPU baseline of bench.py.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * liboracle.so. Nothing under mov-slam_b200/ links, imports or calls it.
 *
 * Parity pinning: the reference ships no tests, golden vectors or fixtures. The raster, EXPRESS,
 * extractor and matcher functions below are pinned against THE REFERENCE'S OWN SOURCES, compiled unmodified
 * into oracle/_ref (Makefile target `ref`; stand-in headers in ref_standin/) and compared on random inputs by
 * tests/test_ref_parity.py, plus the hand-derived known-answer vectors of SURVEY.md Appendix B
 * (tests/test_oracle_kat.py). For arithmetic that lives in un-vendored third-party code (cv::solvePnPRansac,
 * Eigen evaluation order inside Frame::isInFrustum, KannalaBrandt8) the status is "parity unpinned" — see
 * DESIGN.md §Oracle.
 *
 * Float semantics: non-contracted IEEE-754 binary32 (-ffp-contract=off), see SURVEY.md §7 "hard parts".
 */
#ifndef MOVFE_ORACLE_H
#define MOVFE_ORACLE_H

#include "../include/movfe_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- raster: VideoDecoder::NextImage MV loop (src/VideoDecoder.cc:198-351) over a clip ---------------- */
typedef struct orc_clip orc_clip;

/* Runs the decoder's MV loop over n_frames consecutive frames of ONE stream.
 * recs/rec_off: CSR of the frames' side-data records (rec_off has n_frames+1 entries).
 * frame_flags: MOVFE_FRAME_* per frame. max_ref: records with ref > max_ref are dropped and counted.
 * Hops / kps aimed at a frame before the clip start are dropped (the reference under-runs its deque there,
 * VideoDecoder.cc:247,322 — undefined behaviour; the window semantics are defined in DESIGN.md). */
orc_clip *orc_raster_clip(int width, int height, int n_frames, const movfe_mv_record *recs,
                          const int64_t *rec_off, const uint8_t *frame_flags, int max_ref);
void    orc_clip_free(orc_clip *c);
int     orc_clip_n_hops(const orc_clip *c, int frame);
int     orc_clip_n_kps(const orc_clip *c, int frame);
double  orc_clip_coverage(const orc_clip *c, int frame);
int64_t orc_clip_bad_ref(const orc_clip *c);
const int32_t    *orc_clip_grid(const orc_clip *c, int frame); /* H*W*4 int32, row-major, slot-minor */
const movfe_hop  *orc_clip_hops(const orc_clip *c, int frame);
const movfe_rect *orc_clip_kps(const orc_clip *c, int frame);

/* ---- EXPRESS (include/EXPRESS.h) ------------------------------------------------------------------------ */
/* ROI (x0,y0,cols,rows) inside an image with the given row stride. */
int  orc_express_center(const uint8_t *img, int stride, int x0, int y0, int cols, int rows);
void orc_express_descriptor(const uint8_t *img, int stride, int x0, int y0, int cols, int rows,
                            int threshold, uint32_t desc[8]);
int  orc_express_test(const uint8_t *img, int stride, int x0, int y0, int cols, int rows, int threshold);
int  orc_express_distance(const uint32_t a[8], const uint32_t b[8]);

/* ---- propagation: MOVExtractor::operator() (src/MOVExtractor.cc:63-455) for one frame of one stream ----- */
typedef struct orc_extract_params {
    int32_t threshold;            /* MOVExtractor::mThreshold */
    double  coverage_threshold;   /* MOVExtractor::mCoverageThreshold */
    int32_t max_tracks;           /* output capacity; emission stops silently at this count */
} orc_extract_params;

/* prev: previous frame's table, n_prev entries; it is stably sorted in place exactly as the reference
 * sorts prev->mvVF (MOVExtractor.cc:249-252, canonicalised to a stable sort).
 * grid/hops/kps/n_kps/coverage_area: this frame's raster outputs. grey: H*W uint8 (stride = width).
 * lk_status/lk_pts: optional results of the host LK step for the carried features: on a P frame one entry per
 *   coverage feature of prev in sorted order (:337-377), on an I frame one entry per feature of prev in table order
 *   (:81-120); NULL means every carried feature is dropped (cv::calcOpticalFlowPyrLK stays on the host, SURVEY.md §8a
 *   row a8). reloc/n_reloc (orc_extract_frame_lost only): seeds of the lost-relocalisation branch (:161-243) that passed
 *   the host-side tests (:207-215); they are emitted first.
 * current_id: MOVExtractor::mCurrentId, read and updated. Returns the number of tracks written to out. */
int orc_extract_frame(int width, int height, uint32_t frame_flags, const uint8_t *grey,
                      const int32_t *grid, const movfe_hop *hops, const movfe_rect *kps, int n_kps,
                      double coverage_area, movfe_track *prev, int n_prev,
                      const uint8_t *lk_status, const float *lk_pts,
                      const orc_extract_params *params, int32_t *current_id, movfe_track *out,
                      int32_t *n_births /* mov_cnt, may be NULL */);
int orc_extract_frame_lost(int width, int height, uint32_t frame_flags, const uint8_t *grey,
                           const int32_t *grid, const movfe_hop *hops, const movfe_rect *kps, int n_kps,
                           double coverage_area, movfe_track *prev, int n_prev,
                           const uint8_t *lk_status, const float *lk_pts,
                           const movfe_reloc_seed *reloc, int n_reloc,
                           const orc_extract_params *params, int32_t *current_id, movfe_track *out,
                           int32_t *n_births);

/* ---- frustum + joins (src/Frame.cc:456-519, include/MOVMatcher.h:35-137) --------------------------------- */
/* isInFrustum (mono branch) for n points; bounds are [0,width]x[0,height] (Frame.cc:739-745). */
void orc_frustum(const movfe_pose *Tcw, const movfe_camera *cam, int width, int height,
                 float viewing_cos_limit, const movfe_map_point *pts, int n, movfe_projection *out);

/* SearchByVideoFeature(Frame&, vector<MapPoint*>&, bFarPoints, thFarPoints): match[i] = index of the LAST
 * map point (in list order) whose track id maps to track i through the first-wins vfmap; -1 if none.
 * match must be pre-initialised by the caller (entries not hit keep their value). Returns nmatches. */
int orc_search_by_video_feature(const movfe_track *tracks, int n_tracks, const movfe_map_point *pts,
                                const movfe_projection *proj, int n_pts, int far_points, float th_far,
                                int32_t *match);
/* SearchByVideoFeature(KeyFrame*, Frame&, out): match is reset to -1 first. */
int orc_search_by_keyframe(const movfe_track *tracks, int n_tracks, const movfe_map_point *kf_pts, int n_pts,
                           int32_t *match);
/* SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize). prev_matched has n1*2 floats. */
/* Tracking::UpdateLocalPoints (src/Tracking.cc:1171-1198) on index lists into a per-stream point store. */
int orc_update_local_points(const movfe_map_point *store, int n_store, const int32_t *idx, int n_idx, int n_first,
                            movfe_map_point *out, int capacity, int32_t *n_from_first);
int orc_search_for_initialization(const movfe_track *f1, int n1, const movfe_track *f2, int n2,
                                  float *prev_matched, int32_t *matches12);

/* ---- bucket grid (src/Frame.cc:356-388, 602-680) ----------------------------------------------------------- */
/* AssignFeaturesToGrid: cell_start[64*48+1] CSR (cell = ix*48+iy), cell_items[n] keypoint indices. */
void orc_assign_features_to_grid(const movfe_track *tracks, int n, int width, int height,
                                 int32_t *cell_start, int32_t *cell_items);
/* GetFeaturesInArea(x,y,r): returns the count, indices written to out (capacity n). */
int orc_get_features_in_area(const movfe_track *tracks, int n, int width, int height,
                             const int32_t *cell_start, const int32_t *cell_items, float x, float y, float r,
                             int32_t *out);

/* Grid-bucketed search by projection for one frame (include/movfe.h: movfe_search_by_projection; ORB-SLAM3 lineage, not in the
 * reference). taken may be NULL. Returns the number of matches. */
int orc_search_by_projection(const movfe_track *feat, const uint8_t *taken, int n_feat, int width, int height,
                             const movfe_map_point *pts, const movfe_projection *proj, const uint32_t *pt_desc, int n_pts,
                             const movfe_projection_search_params *prm, int32_t *feat_match, int32_t *pt_match,
                             int32_t *pt_dist);

/* ---- pose-only Gauss-Newton / Huber (OptimizableTypes.cpp:54-69, Pinhole.cpp:77-88) --- */
/* Camera model maths in double. */
void orc_project(const movfe_camera *cam, const double Xc[3], double uv[2]);
void orc_project_jac(const movfe_camera *cam, const double Xc[3], double J[6]); /* 2x3 row-major */
/* 2x6 Jacobian of e = obs - pi(T*Xw) wrt the left se3 update [omega, upsilon] */
void orc_pose_jacobian(const movfe_camera *cam, const double Xc[3], double J[12]);
double orc_huber_weight(double chi2, double delta);
void orc_se3_exp(const double dx[6], double R[9], double t[3]);

/* Optimizer::PoseOptimization on gathered correspondences (Optimizer.cc:404-413 gather is done by caller).
 * pts: n world points (xyz float), obs: n image points (uv float). pose: in = initial T_cw, out = result
 * (untouched when 0 is returned for n<4). outlier: n bytes (1 = outlier). stats (optional, 4 ints):
 * {gauss-newton iterations executed, rounds executed, passes over the correspondences, solver failures}.
 * Returns the inlier count. */
int orc_pose_optimize(const movfe_camera *cam, const movfe_pose_params *params, const float *pts,
                      const float *obs, int n, movfe_pose *pose, uint8_t *outlier, int32_t *stats);

/* ---- whole front-end over one stream (CPU baseline driver) ------------------------------------------------- */
typedef struct orc_frontend_cfg {
    int32_t width, height, n_frames, max_ref, max_tracks;
    int32_t threshold;
    double  coverage_threshold;
    movfe_camera cam;
    movfe_pose_params pose_params;
    int32_t n_kf_points;          /* the first n_kf_points map points play the reference keyframe's list */
    float   viewing_cos_limit;
} orc_frontend_cfg;

typedef struct orc_frontend_out {
    movfe_pose *poses;            /* n_frames */
    int32_t    *n_tracks;         /* n_frames */
    int32_t    *n_inliers;        /* n_frames (second PoseOptimization) */
    movfe_track *last_tracks;     /* max_tracks: table of the last frame */
    uint64_t   *track_hash;       /* n_frames: checksum of the frame's track table (frontend.cc: table_checksum; may be NULL) */
} orc_frontend_out;

/* Runs raster -> extract -> [join(kf) -> pose -> frustum -> join(local) -> pose] per frame, as Tracking.cc
 * drives them (Tracking.cc:796-811, 890-905, 1109-1158). grey: n_frames*H*W or NULL (flat 128 image).
 * seed_tracks/n_seed: optional table of the frame before the clip (MV-only configs seed tracks as input state). */
int orc_frontend_run(const orc_frontend_cfg *cfg, const movfe_mv_record *recs, const int64_t *rec_off,
                     const uint8_t *frame_flags, const uint8_t *grey, const movfe_track *seed_tracks, int n_seed,
                     const movfe_map_point *map_pts, int n_map, const movfe_pose *pose0,
                     orc_frontend_out *out);

/* The same with (a) a schedule of local maps: before frame sched_frame[i] (ascending) the map becomes
 * sched_pts[sched_off[i] .. sched_off[i+1]) with sched_nkf[i] keyframe points - what the mapping side (out of scope)
 * hands the tracker after a keyframe insertion; (b) wall-clock stamps: tail_times[0] when frame `timed_from` starts,
 * tail_times[1] at the end (steady_clock seconds; may be NULL) - bench.py times only the frames the GPU arm times. */
int orc_frontend_run_sched(const orc_frontend_cfg *cfg, const movfe_mv_record *recs, const int64_t *rec_off,
                           const uint8_t *frame_flags, const uint8_t *grey, const movfe_track *seed_tracks, int n_seed,
                           const movfe_map_point *map_pts, int n_map, const movfe_pose *pose0,
                           int n_sched, const int32_t *sched_frame, const int64_t *sched_off,
                           const movfe_map_point *sched_pts, const int32_t *sched_nkf, int timed_from,
                           double *tail_times, orc_frontend_out *out);

#ifdef __cplusplus
}
#endif
#endif
