/*
 * movfe.h — C-ABI of the B200 tracking front-end (libmovfe.so).
 *
 * The reference (MoV-SLAM) has no plugin/FFI layer: its boundary is the C++ signatures of VideoDecoder,
 * MOVExtractor, MOVMatcher and Optimizer::PoseOptimization. This header is what those classes
 * bind to in the drop-in shims under mov-slam_b200/shim/ (see INTEGRATION.md). Each entry point cites the
 * reference interface it replaces. Plain pointers and sizes only; no CUDA, torch or C++ types.
 *
 * Conventions
 *  - Every call returns MOVFE_OK (0) or a negative MOVFE_E_* code; movfe_last_error() gives the text.
 *  - A context owns one GPU, its CUDA streams and all device buffers. It is not thread-safe; use one context
 *    per host thread / GPU (streams shard across GPUs with no collective).
 *  - Work is batched over the context's n_streams independent video streams. All batched arrays are
 *    stream-major: element (stream s, frame f) of a call with n frames lives at index s*n + f.
 *  - Calls enqueue work on the context's streams and return; movfe_synchronize() or any download waits.
 *  - There is no CPU fallback: without a CUDA device movfe_create fails with MOVFE_E_CUDA.
 */
#ifndef MOVFE_H
#define MOVFE_H

#include "movfe_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define MOVFE_OK            0
#define MOVFE_E_INVALID    -1   /* bad argument */
#define MOVFE_E_CUDA       -2   /* CUDA runtime error (no device, launch failure, out of memory) */
#define MOVFE_E_CAPACITY   -3   /* input exceeds a capacity fixed at movfe_create */
#define MOVFE_E_STATE      -4   /* call order violated (e.g. raster of frames that were never pushed) */

typedef struct movfe_ctx movfe_ctx;

typedef struct movfe_config {
    int32_t device;                 /* CUDA device ordinal */
    int32_t n_streams;              /* independent video streams batched on this GPU */
    int32_t width, height;          /* frame size, identical for all streams of a context */
    int32_t max_records_per_frame;  /* capacity of one frame's side data (4 per macroblock for H.264) */
    int32_t max_ref;                /* largest accepted reference index K; look-ahead is K+1 frames.
                                       The reference's 12-deep decoder queue bounds K at 10 (VideoDecoder.cc:163) */
    int32_t window_frames;          /* F: frames per stream rasterised / tracked per call */
    int32_t max_tracks;             /* capacity of a frame's track table (VideoFeature list) */
    int32_t max_map_points;         /* capacity of a stream's local map-point table */
    int32_t express_threshold;      /* MOVExtractor::mThreshold (Settings.cc:376-382) */
    double  coverage_threshold;     /* MOVExtractor::mCoverageThreshold */
    int32_t has_grey;               /* 1: grey planes are pushed (descriptor gating on); 0: MV-only mode ==
                                       the reference's behaviour on a flat image */
    int32_t flags;                  /* MOVFE_CFG_* bits, 0 by default */
} movfe_config;

/* By default the ingest + raster kernels of window k+1 run on their own low-priority CUDA stream beside the propagation
 * of window k (raster results are double-buffered). With this flag every movfe_raster first waits for all propagation
 * and pose work enqueued so far, so that the raster kernels run alone - bench.py uses it to time grid_kernel for the roofline. */
#define MOVFE_CFG_SERIAL_RASTER 1
/* Fused (grid-free) mode: VideoImage::mvi, the per-pixel slot grid, is not materialised. Propagation reads the grid at one
 * pixel per track (src/MOVExtractor.cc:264-272) and at the 16-px lattice of a back-fill pass (:431), so the raster stage keeps
 * only the ordered per-32x32-tile hop queues it builds anyway and every lookup resolves its four slots from its tile's queue.
 * Results are bit-identical to the grid path (tests/test_pipeline_gpu.py); 16*W*H bytes per frame are never written and the
 * grid's memory (2 x 16 W H F S bytes) is not allocated. movfe_download_grid and movfe_extract_frame are unavailable. */
#define MOVFE_CFG_NO_GRID 2

/* -- lifetime -------------------------------------------------------------------------------------------- */
int         movfe_create(const movfe_config *cfg, movfe_ctx **out);
void        movfe_destroy(movfe_ctx *ctx);
const char *movfe_last_error(const movfe_ctx *ctx);   /* ctx may be NULL: error of the last failed create */
int         movfe_synchronize(movfe_ctx *ctx);        /* waits for all of the context's streams */
void       *movfe_cuda_stream(movfe_ctx *ctx);        /* the primary cudaStream_t (propagation, single-shot operators) */
/* The context also owns a copy stream (host->device staging of movfe_push_frames), a raster stream (ingest + raster of
 * window k+1 run beside the propagation of window k) and a pose stream (movfe_track_poses runs beside raster/propagation
 * of the next window). movfe_fence makes the primary stream wait, on the device, for everything enqueued so far on the
 * raster and pose streams - call it before recording a timing event or enqueueing dependent work on
 * movfe_cuda_stream(). Host buffers handed to movfe_push_frames must stay unchanged until the next call that
 * synchronises (any download, movfe_synchronize) or until the second following push. Device buffers handed to
 * movfe_push_frames_device must be complete when the call is made and stay unchanged until the next movfe_fence /
 * movfe_synchronize has completed. */
int         movfe_fence(movfe_ctx *ctx);
const char *movfe_version(void);

/* -- ingest: replaces av_frame_get_side_data -> the loop head of VideoDecoder::NextImage
 *    (src/VideoDecoder.cc:198-211). Appends n_frames frames to every stream; frames get consecutive absolute
 *    indices starting at movfe_frames_pushed(). recs: all records, packed, stream-major then frame order;
 *    rec_off: n_streams*n_frames+1 offsets into recs; frame_flags: MOVFE_FRAME_*; grey: n_streams*n_frames
 *    planes of width*height bytes or NULL. The *_device variant takes device pointers (inputs already in HBM). */
int     movfe_push_frames(movfe_ctx *ctx, int n_frames, const movfe_mv_record *recs, const int64_t *rec_off,
                          const uint8_t *frame_flags, const uint8_t *grey);
int     movfe_push_frames_device(movfe_ctx *ctx, int n_frames, const movfe_mv_record *d_recs,
                                 const int64_t *d_rec_off, int64_t n_records, const uint8_t *d_frame_flags,
                                 const uint8_t *d_grey);
/* The same hand-over with 16-byte records (movfe_packed_record): the shim packs while it copies the side data out of the
 * AVFrame (it has to copy it anyway - the side data belongs to the frame), and the push moves 16 instead of 40 bytes per
 * record. movfe_pack_records is plain host code (no GPU involved); results are identical to pushing the 40-byte records.
 * grey_stride: bytes between the rows of a luma plane (AVFrame::linesize[0] / cv::Mat::step; 0 = width); a plane is grey_stride *
 * height bytes and the planes follow each other without gaps. The rows go straight into the device's pitched ring. */
void    movfe_pack_records(const movfe_mv_record *recs, int64_t n_records, movfe_packed_record *out);
int     movfe_push_frames_packed(movfe_ctx *ctx, int n_frames, const movfe_packed_record *recs, const int64_t *rec_off,
                                 const uint8_t *frame_flags, const uint8_t *grey, int grey_stride);
int64_t movfe_frames_pushed(const movfe_ctx *ctx);

/* -- raster: replaces the MV loop of VideoDecoder::NextImage (src/VideoDecoder.cc:211-350).
 *    Produces, for frames [first_frame, first_frame+n_out) of every stream, the hop list (VideoImage::mvs),
 *    the candidate-keypoint list (kps), the per-pixel slot grid (mvi) and coverageArea, bit-identical to the
 *    reference given the same records. Records of up to max_ref+1 later frames that were already pushed are
 *    used as look-ahead (they back-fill hops/kps into earlier frames, VideoDecoder.cc:245-248,315-323);
 *    hops aimed at frames before first_frame are not re-emitted (they belong to the previous window). */
int movfe_raster(movfe_ctx *ctx, int64_t first_frame, int n_out);

/* Results of the last movfe_raster call (download = device->host copy + synchronise). */
int movfe_raster_counts(movfe_ctx *ctx, int stream, int64_t frame, int32_t *n_hops, int32_t *n_kps,
                        double *coverage_area);
int movfe_download_grid(movfe_ctx *ctx, int stream, int64_t frame, int32_t *out /* height*width*4 */);
int movfe_download_hops(movfe_ctx *ctx, int stream, int64_t frame, movfe_hop *out, int capacity);
int movfe_download_kps(movfe_ctx *ctx, int stream, int64_t frame, movfe_rect *out, int capacity);
int64_t movfe_rejected_records(movfe_ctx *ctx);   /* records dropped for ref > max_ref or capacity, since create */

/* -- propagation: replaces MOVExtractor::operator() (include/MOVExtractor.h:36-37, src/MOVExtractor.cc:63-455)
 *    for frames [first_frame, first_frame+n_frames) of every stream, in frame order, on the raster results of
 *    the last movfe_raster call. Track tables persist per stream across calls (prev frame = Frame::mpPrevFrame).
 *
 *    LK hand-over. cv::calcOpticalFlowPyrLK (:91,:196,:347) is OpenCV arithmetic and stays on the host; its RESULTS enter
 *    through movfe_set_lk_results and are merged on the device exactly where the reference merges them. They apply to the
 *    NEXT frame propagated on `stream` (the first frame of the next movfe_extract / the frame of movfe_extract_frame) and
 *    are cleared after it, so a caller that needs LK propagates one frame per call (the carried set of frame f+1 depends on
 *    the table of frame f, which the host needs to run LK). Which features the reference hands to LK:
 *      I frame with a non-empty previous table (:81-120): every previous track, in TABLE order: n = n_prev;
 *      P frame (:337-377): the previous table's coverage tracks (MOVFE_TRACK_COVERAGE) in SORTED order (stable, age
 *        descending then descriptor popcount descending, :249-252): n = their count;
 *      status[i] / pts_xy[2i..2i+1] = LK status and position of the i-th such feature; the bounds test (:98,:354) is applied
 *        on the device. n = -1: nothing installed (carried tracks are dropped and counted, see below).
 *      reloc / n_reloc: lost relocalisation (prev->mLost, :161-243): keyframe points the host carried with LK that passed
 *        status, image bounds and the distance test (:207-215); the device adds block, bounds test and descriptor
 *        (:218-238) and emits them ahead of the propagated tracks. Needs grey planes.
 *    Without results every carried track of a frame is dropped - what the reference does when LK loses them all - and
 *    movfe_dropped_lk_tracks() counts them since create, so a drop-in can tell. */
int movfe_set_tracks(movfe_ctx *ctx, int stream, const movfe_track *tracks, int n, int32_t current_id);
int movfe_set_lk_results(movfe_ctx *ctx, int stream, const uint8_t *status, const float *pts_xy, int n,
                         const movfe_reloc_seed *reloc, int n_reloc);
int64_t movfe_dropped_lk_tracks(movfe_ctx *ctx);
int movfe_extract(movfe_ctx *ctx, int64_t first_frame, int n_frames);
int movfe_track_count(movfe_ctx *ctx, int stream, int64_t frame, int32_t *n_tracks, int32_t *current_id);
int movfe_download_tracks(movfe_ctx *ctx, int stream, int64_t frame, movfe_track *out, int capacity);

/* -- matching: replaces Frame::isInFrustum (src/Frame.cc:456-519) + MOVMatcher::SearchByVideoFeature
 *    (include/MOVMatcher.h:35-68, 70-103) and -- pose: replaces Optimizer::PoseOptimization
 *    (include/Optimizer.h:55, src/Optimizer.cc:397-459), batched over streams and run per frame in the order
 *    Tracking.cc drives them (TrackReferenceKeyFrame :796-811, TrackLocalMap :890-905). */
int movfe_set_camera(movfe_ctx *ctx, const movfe_camera *cam, const movfe_pose_params *pp, float viewing_cos_limit);
int movfe_set_map_points(movfe_ctx *ctx, int stream, const movfe_map_point *pts, int n, int n_keyframe_points);
/* The local maps of ALL streams in one call - what the mapping side hands the tracker after a keyframe insertion
 * (Tracking::UpdateLocalPoints, src/Tracking.cc:1171-1198: the list building stays with the mapping code; its result arrives
 * here). pts: packed points, stream-major; off: n_streams+1 offsets; n_keyframe_points: per stream, how many of its leading
 * points are the reference keyframe's list (TrackReferenceKeyFrame's join); max_points_per_stream: upper bound of any
 * stream's count (<= max_map_points). on_device = 1: all three arrays are device memory, complete when the call is made (a
 * device-side mapper / a map kept resident); 0: host memory, unchanged until the next call that synchronises. Ordered with
 * movfe_track_poses: frames tracked after the call see the new maps. */
int movfe_set_map_points_batch(movfe_ctx *ctx, const movfe_map_point *pts, const int64_t *off, const int32_t *n_keyframe_points,
                               int max_points_per_stream, int on_device);
/* -- local map building on the device: replaces Tracking::UpdateLocalPoints (src/Tracking.cc:1171-1198) for batched callers.
 *    The map points of a stream live in a device-resident STORE (movfe_set_map_store: `first_index`.. are the caller's own point
 *    indices; call it again to patch points the mapping side added, moved or culled - flags carry MOVFE_MP_BAD). A local map is
 *    then built from INDEX lists: for every stream the map-point matches of its local keyframes, concatenated in the order the
 *    caller walks them (the reference: mvpLocalKeyFrames reversed, each keyframe's GetMapPointMatches() in order; a NULL entry
 *    is -1). The kernel skips NULL and bad points, keeps the FIRST occurrence of every point in order (mnTrackReferenceForFrame)
 *    and installs the result as the stream's local map, as movfe_set_map_points would; n_keyframe_entries[s] leading list
 *    entries are the reference keyframe's list (their surviving points become the local map's keyframe prefix). 4 bytes per
 *    list entry cross PCIe instead of 40 per point. idx: all lists, stream-major; off: n_streams + 1 offsets. */
int movfe_reserve_map_store(movfe_ctx *ctx, int max_points_per_stream);
int movfe_set_map_store(movfe_ctx *ctx, int stream, int first_index, const movfe_map_point *pts, int n);
int movfe_update_local_points(movfe_ctx *ctx, const int32_t *idx, const int64_t *off, const int32_t *n_keyframe_entries);
int movfe_download_map_points(movfe_ctx *ctx, int stream, movfe_map_point *out, int capacity, int32_t *n_keyframe_points);
int movfe_set_pose(movfe_ctx *ctx, int stream, const movfe_pose *pose);
int movfe_track_poses(movfe_ctx *ctx, int64_t first_frame, int n_frames);
int movfe_download_poses(movfe_ctx *ctx, int64_t first_frame, int n_frames, movfe_pose *poses /* S*n */,
                         int32_t *n_inliers /* S*n, may be NULL */);
int movfe_download_matches(movfe_ctx *ctx, int stream, int64_t frame, int32_t *match, uint8_t *outlier, int capacity);

/* -- instrumentation (bench.py): CUDA events around every stage on the context's stream, kernel-launch counts --- */
#define MOVFE_STAGE_INGEST   0   /* ingest_kernel (+ meta, grey copies) */
#define MOVFE_STAGE_HOPS     1   /* count / bases / emit / bbox */
#define MOVFE_STAGE_GRID     2   /* grid_kernel: the per-pixel slot grid (dominant HBM writer) */
#define MOVFE_STAGE_EXTRACT  3   /* propagation kernels */
#define MOVFE_STAGE_POSE     4   /* join / frustum / pose kernels */
#define MOVFE_N_STAGES       5
/* Workload counters since the last reset (synchronises): out[0] tracks looked up in the slot grid, [1] candidate hops of those
 * tracks, [2] pose solves, [3] correspondences of those solves, [4] passes over the correspondences (Gauss-Newton iterations +
 * re-classifications), [5] hops of the rastered frames, [6] frames rastered, [7] reserved. bench.py turns them into the
 * per-frame figures SURVEY.md 8d's byte formulas need. */
int movfe_workload_stats(movfe_ctx *ctx, uint64_t *out /* 8 */, int reset);
int movfe_profile_enable(movfe_ctx *ctx, int on);
/* Waits for the stream, then adds up the event-timed milliseconds and kernel launches per stage since the last
 * reset. ms / launches have MOVFE_N_STAGES entries (either may be NULL). */
int movfe_profile_read(movfe_ctx *ctx, double *ms, int64_t *launches, int reset);

/* -- single-shot operators (batch of independent problems; used by the drop-in shims and the parity tests) --- */
/* MOVExtractor::operator() (include/MOVExtractor.h:36-37) for ONE frame whose raster results the caller holds on the
 * host: grid = VideoImage::mvi (height*width*4 int32), hops = mvs, kps, coverage_area, frame_flags = MOVFE_FRAME_*,
 * grey = imGray with rows of grey_stride bytes (cv::Mat::step / AVFrame::linesize[0]; 0 = width; NULL for a context
 * created with has_grey = 0); prev/n_prev = prev->mvVF (any order: the reference's stable sort is applied);
 * lk_status / lk_pts / n_lk / reloc / n_reloc = host LK results for this frame as in movfe_set_lk_results (n_lk = -1 and
 * n_reloc = 0: none); *current_id = MOVExtractor::mCurrentId, read and updated. Writes the new frame's table to out and
 * returns its size (which may exceed `capacity`: MOVFE_E_CAPACITY). The context must have n_streams == 1 and must not be
 * mixed with the batched push/raster/extract calls. */
int movfe_extract_frame(movfe_ctx *ctx, uint32_t frame_flags, const uint8_t *grey, int grey_stride, const int32_t *grid,
                        const movfe_hop *hops, int n_hops, const movfe_rect *kps, int n_kps, double coverage_area,
                        const movfe_track *prev, int n_prev, const uint8_t *lk_status, const float *lk_pts, int n_lk,
                        const movfe_reloc_seed *reloc, int n_reloc, int32_t *current_id, movfe_track *out, int capacity);
/* Frame::isInFrustum for n_problems point sets. pts/out are packed; off has n_problems+1 entries. */
int movfe_frustum(movfe_ctx *ctx, int n_problems, const movfe_pose *poses, const movfe_map_point *pts,
                  const int32_t *off, movfe_projection *out);
/* Track-id join (MOVMatcher.h:35-137): for problem p, match[t] = index of the LAST probe (in order) whose key
 * equals track t's id through the first-wins id->index map; probes with valid[i]==0 are skipped.
 * match entries not hit keep their input value. n_matches[p] counts hits. */
int movfe_join(movfe_ctx *ctx, int n_problems, const int32_t *track_ids, const int32_t *track_off,
               const int32_t *probe_ids, const uint8_t *probe_valid, const int32_t *probe_off, int32_t *match,
               int32_t *n_matches);
/* Frame::AssignFeaturesToGrid (src/Frame.cc:356-388; Frame::PosInGrid :670-680, which rounds) for n_problems keypoint
 * sets of a width x height frame (mono, undistorted: mnMinX = mnMinY = 0). pts_xy: packed (x, y) pairs = mvKeysUn[i].pt,
 * off: n_problems+1 offsets. cell_start: n_problems * (64*48+1) CSR offsets local to the set, cell = ix*48 + iy
 * (mGrid[ix][iy]); cell_items: same length as the points, the set's keypoint indices cell by cell in insertion order
 * (entries past cell_start[64*48] are -1: keypoints outside the grid). At most 16384 keypoints per set. */
int movfe_assign_features_to_grid(movfe_ctx *ctx, int n_problems, const float *pts_xy, const int32_t *off,
                                  int32_t *cell_start, int32_t *cell_items);
/* The same grid for the RESIDENT track table of (stream, frame) after movfe_extract - what the Frame constructors build
 * (src/Frame.cc:118,216) - without the keypoints leaving the device. Returns the number of keypoints (<= capacity);
 * cell_start has 64*48+1 entries, cell_items `capacity`. */
int movfe_track_feature_grid(movfe_ctx *ctx, int stream, int64_t frame, int32_t *cell_start, int32_t *cell_items,
                             int capacity);
/* Frame::GetFeaturesInArea (src/Frame.cc:602-668) for n_queries queries against the grids built above: out[q*capacity..]
 * receives the indices in the reference's (ix, iy, insertion) order, counts[q] the full count (it may exceed capacity,
 * in which case only the first `capacity` indices were written). */
int movfe_features_in_area(movfe_ctx *ctx, int n_problems, const float *pts_xy, const int32_t *off,
                           const int32_t *cell_start, const int32_t *cell_items, int n_queries,
                           const movfe_area_query *queries, int capacity, int32_t *out, int32_t *counts);
/* Grid-bucketed search by projection (north_star subsystem 3). The reference builds the
 * bucket grid for every frame (src/Frame.cc:356-388) and keeps Frame::GetFeaturesInArea (:602-668), but its matcher joins by track
 * id and never queries the grid; this operator is the query those two exist for, restated from the ORB-SLAM3 lineage for MoV-SLAM's
 * types: for every map point that movfe_frustum left in view (and that is not bad / already matched / beyond th_far), the
 * keypoints of GetFeaturesInArea(u, v, r) are visited in the reference's (ix, iy, insertion) order, features with taken[i] != 0
 * (already holding an observed map point) are passed over, and the candidate with the smallest EXPRESS distance
 * (descriptor1 ^ descriptor2).count() is kept with the second smallest beside it; th_high and the nn_ratio test decide whether
 * it is a match. Map points are searched independently (that is what makes the search a batch); when two of them choose the same
 * keypoint the smaller (distance, point index) holds it and the other stays unmatched - where the sequential original lets
 * the later point take another keypoint. n_problems frames: feat = the frame's keypoints with their descriptors as track records
 * (pt_x, pt_y, desc are read), feat_off / pt_off = n_problems + 1 offsets, pts / proj = the arrays given to and filled by
 * movfe_frustum, pt_desc = 8 words per map point. Outputs: feat_match[i] = index (local to the frame's point list) of the
 * map point matched to keypoint i or -1, pt_match[k] = keypoint index of map point k or -1, pt_dist[k] = its best distance when it
 * passed th_high / nn_ratio (whether or not it then held the keypoint) or -1, n_matches[p]. At most 16384 keypoints per frame. */
int movfe_search_by_projection(movfe_ctx *ctx, int n_problems, const movfe_track *feat, const uint8_t *feat_taken /* may be NULL */,
                               const int32_t *feat_off, const movfe_map_point *pts, const movfe_projection *proj,
                               const uint32_t *pt_desc, const int32_t *pt_off, const movfe_projection_search_params *prm,
                               int32_t *feat_match, int32_t *pt_match, int32_t *pt_dist, int32_t *n_matches);
/* -- pyramidal Lucas-Kanade: replaces cv::calcOpticalFlowPyrLK as the reference calls it for its carry-over branches
 *    (src/MOVExtractor.cc:91-92 I-frame carry-over, :196-197 lost relocalisation, :347-348 coverage tracks: win_size 31, max_level 3,
 *    max_count 20, epsilon 0.01, OPTFLOW_LK_GET_MIN_EIGENVALS, min_eig_threshold 1e-4; src/Frame.cc:305: win_size 21).
 *    n_problems image pairs of the context's width x height (host memory, rows `stride` bytes apart, 0 = width; pair i at
 *    prev + i*stride*height); the points of pair i are pts_xy[2*off[i] .. 2*off[i+1]). Per point: next position, status (1 =
 *    tracked) and err = the minimum eigenvalue of the window's gradient matrix at level 0. OpenCV's arithmetic (fixed-point
 *    bilinear weights, int16 Scharr derivatives, reflected image border, zero derivative border); positions agree with
 *    OpenCV to ~1e-3 px. The results are what movfe_set_lk_results / movfe_extract_frame take. */
int movfe_lk(movfe_ctx *ctx, int n_problems, const uint8_t *prev, const uint8_t *next, int stride, const float *pts_xy, const int32_t *off,
             int win_size, int max_level, int max_count, double epsilon, double min_eig_threshold, float *out_xy, uint8_t *status, float *err);
/* The same tracker for the batched path, device-resident: Lucas-Kanade results for frame `frame` of EVERY stream, from the grey planes
 * of frame - 1 and frame in the ring and the track table of frame - 1 (an intra picture: every track in table order,
 * src/MOVExtractor.cc:81-120; a P picture: the coverage tracks in sorted order, :337-377), installed as movfe_set_lk_results
 * would install them. Call it between movfe_extract(.., frame - 1) and movfe_extract(frame, ..). Nothing crosses PCIe. */
int movfe_lk_carry(movfe_ctx *ctx, int64_t frame);
/* Optimizer::PoseOptimization for n_problems correspondence sets (pts xyz float, obs uv float, packed). */
int movfe_pose_optimize(movfe_ctx *ctx, int n_problems, const movfe_camera *cam, const movfe_pose_params *pp,
                        const float *pts, const float *obs, const int32_t *off, movfe_pose *poses /* in/out */,
                        uint8_t *outlier, int32_t *n_inliers, int32_t *stats /* 4 per problem, may be NULL */);

#ifdef __cplusplus
}
#endif
#endif /* MOVFE_H */
