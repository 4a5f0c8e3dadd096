// ref_driver.cc — TEST INFRASTRUCTURE. C API over the REFERENCE'S OWN front-end code, compiled unmodified from
// /root/reference by oracle/Makefile (target _ref) against the stand-in headers of this directory:
//   src/VideoDecoder.cc      VideoDecoder::Init / NextImage (the MV loop, :198-351) on a fake libav back-end
//   src/MOVExtractor.cc      MOVExtractor::operator()        (cv::calcOpticalFlowPyrLK = injected results)
//   include/EXPRESS.h        compute_center / compute_descriptor / compute_express / compute_distance
//   include/MOVMatcher.h     SearchByVideoFeature x2, SearchForInitialization
//   include/Frame.h, include/MOVExtractor.h, include/VideoDecoder.h, include/VideoBase.h   (declarations)
// This file only marshals flat arrays (include/movfe_types.h) into the reference's containers and back; it holds no
// algorithm. tests/test_ref_parity.py compares the oracle (oracle/*.cc) against it; nothing under mov-slam_b200/ may
// load it.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/movfe_types.h"
#include "VideoDecoder.h"
#include "MOVExtractor.h"
#include "MOVMatcher.h"
#include "EXPRESS.h"

// ---- definitions the stand-in headers declare ------------------------------------------------------------------------
namespace cvstub {
std::vector<LkCall> &lk_queue() {
    static std::vector<LkCall> q;
    return q;
}
int &lk_calls() {
    static int n = 0;
    return n;
}
std::vector<cv::Point2f> &lk_last_prev_pts() {
    static std::vector<cv::Point2f> p;
    return p;
}
}  // namespace cvstub

// Frame's constructors live in src/Frame.cc, which is not compiled here (it drags in g2o / boost / Pinhole); the driver
// only needs an empty Frame to fill by hand.
MOV_SLAM::Frame::Frame() : mpcpi(nullptr), mbHasPose(false), mbHasVelocity(false), mpMOVExtractor(nullptr), mTimeStamp(0), mLost(false), N(0),
                           imageCols(0), imageRows(0), mpImuPreintegrated(nullptr), mpLastKeyFrame(nullptr), mpPrevFrame(nullptr),
                           mpImuPreintegratedFrame(nullptr), mpReferenceKF(nullptr), mbIsSet(false), mbImuPreintegrated(false),
                           mpMutexImu(nullptr), mpCamera(nullptr), mpCamera2(nullptr) {}

namespace {
std::string g_err;

movfe_track to_track(const MOV_SLAM::VideoFeature &vf) {
    movfe_track t;
    t.pt_x = vf.pt.x;
    t.pt_y = vf.pt.y;
    t.mb = {(int16_t)vf.mb.x, (int16_t)vf.mb.y, (int16_t)vf.mb.width, (int16_t)vf.mb.height};
    t.track_id = vf.trackId;
    t.age = vf.age;
    t.q_indx = vf.qIndx;
    t.flags = vf.coverage ? MOVFE_TRACK_COVERAGE : 0u;
    for (int k = 0; k < 8; k++) t.desc[k] = 0;
    for (int b = 0; b < 256; b++)
        if (vf.desc[b]) t.desc[b >> 5] |= 1u << (b & 31);
    return t;
}

MOV_SLAM::VideoFeature to_vf(const movfe_track &t, int d_indx) {
    MOV_SLAM::VideoFeature vf;
    vf.trackId = t.track_id;
    vf.qIndx = t.q_indx;
    vf.dIndx = d_indx;
    vf.pt = cv::Point2f(t.pt_x, t.pt_y);
    vf.mb = cv::Rect(t.mb.x, t.mb.y, t.mb.w, t.mb.h);
    vf.age = t.age;
    vf.coverage = (t.flags & MOVFE_TRACK_COVERAGE) != 0;
    for (int b = 0; b < 256; b++) vf.desc[b] = (t.desc[b >> 5] >> (b & 31)) & 1u;
    return vf;
}

cv::Mat view_u8(const uint8_t *img, int stride, int x0, int y0, int cols, int rows) {
    return cv::Mat(rows, cols, CV_8UC1, const_cast<uint8_t *>(img) + (size_t)y0 * stride + x0, (size_t)stride);
}
}  // namespace

struct ref_clip {
    int W, H;
    std::vector<std::shared_ptr<MOV_SLAM::MotionVectorImage>> frames;
    std::vector<std::vector<movfe_hop>> hops;
    std::vector<std::vector<movfe_rect>> kps;
};

extern "C" {

const char *ref_last_error(void) { return g_err.c_str(); }

// -------------------------------------------------------------------------------------------------- VideoDecoder -----
// Runs MOV_SLAM::VideoDecoder(path, qlen) over a synthetic clip and keeps every VideoImage it returns.
// Frames without MOVFE_FRAME_MV carry no side data (same effect as NextImage(false): VideoDecoder.cc:200).
// Returns NULL when the clip would make the reference index its deque out of range (VideoDecoder.cc:247,322: undefined
// behaviour, not a result): a record with source <= 0 and ref > 0 needs ref < (frames queued when it is decoded).
ref_clip *ref_decode_clip(int width, int height, int n_frames, const movfe_mv_record *recs, const int64_t *rec_off,
                          const uint8_t *frame_flags, const uint8_t *grey, int qlen) {
    static_assert(sizeof(AVMotionVector) == sizeof(movfe_mv_record) && sizeof(AVMotionVector) == 40, "record layout");
    for (int f = 0; f < n_frames; f++) {
        if (!(frame_flags[f] & MOVFE_FRAME_MV)) continue;
        const int queued = f < qlen - 1 ? f : qlen - 1;
        for (int64_t i = rec_off[f]; i < rec_off[f + 1]; i++)
            if (recs[i].source <= 0 && recs[i].ref > 0 && recs[i].ref > queued - 1) {
                g_err = "record " + std::to_string(i) + " of frame " + std::to_string(f) + ": ref " + std::to_string(recs[i].ref) +
                        " under-runs the reference's decoder queue (undefined behaviour in the reference)";
                return nullptr;
            }
    }
    std::vector<uint8_t> is_p(n_frames);
    std::vector<const uint8_t *> luma(n_frames, nullptr), side(n_frames, nullptr);
    std::vector<int> side_bytes(n_frames, 0);
    for (int f = 0; f < n_frames; f++) {
        is_p[f] = (frame_flags[f] & MOVFE_FRAME_P) ? 1 : 0;
        if (grey) luma[f] = grey + (size_t)f * width * height;
        if (frame_flags[f] & MOVFE_FRAME_MV) {
            side[f] = reinterpret_cast<const uint8_t *>(recs + rec_off[f]);
            side_bytes[f] = (int)((rec_off[f + 1] - rec_off[f]) * (int64_t)sizeof(movfe_mv_record));
            if (side_bytes[f] == 0) side[f] = nullptr;  // libavcodec attaches no side data to a picture without vectors
        }
    }
    fake_av_clip fc = {width, height, n_frames, is_p.data(), luma.data(), side.data(), side_bytes.data()};
    fake_av_install(&fc);

    ref_clip *c = new ref_clip;
    c->W = width;
    c->H = height;
    {
        MOV_SLAM::VideoDecoder dec("fake://clip", qlen);
        if (!dec.Init()) {
            g_err = "VideoDecoder::Init failed";
            delete c;
            return nullptr;
        }
        while (true) {
            std::shared_ptr<MOV_SLAM::MotionVectorImage> smv = dec.NextImage(true);
            if (!smv) break;
            c->frames.push_back(smv);
        }
    }
    if ((int)c->frames.size() != n_frames) {
        g_err = "VideoDecoder returned " + std::to_string(c->frames.size()) + " frames, expected " + std::to_string(n_frames);
        delete c;
        return nullptr;
    }
    c->hops.resize(n_frames);
    c->kps.resize(n_frames);
    for (int f = 0; f < n_frames; f++) {
        for (const auto &mv : c->frames[f]->mvs) c->hops[f].push_back({mv.pt.x, mv.pt.y, mv.dIndx, 0});
        for (const auto &r : c->frames[f]->kps) c->kps[f].push_back({(int16_t)r.x, (int16_t)r.y, (int16_t)r.width, (int16_t)r.height});
    }
    return c;
}
void ref_clip_free(ref_clip *c) { delete c; }
int ref_clip_n_hops(const ref_clip *c, int f) { return (int)c->hops[f].size(); }
int ref_clip_n_kps(const ref_clip *c, int f) { return (int)c->kps[f].size(); }
double ref_clip_coverage(const ref_clip *c, int f) {
    // coverageArea is only assigned when side data was processed (VideoDecoder.cc:350); otherwise the member is uninitialised
    return c->frames[f]->coverageArea;
}
int ref_clip_frame_no(const ref_clip *c, int f) { return c->frames[f]->frame; }
int ref_clip_is_p(const ref_clip *c, int f) { return c->frames[f]->ft == MOV_SLAM::FrameType::P_FRAME; }
const int32_t *ref_clip_grid(const ref_clip *c, int f) { return reinterpret_cast<const int32_t *>(c->frames[f]->mvi.data); }
const uint8_t *ref_clip_grey(const ref_clip *c, int f) { return c->frames[f]->imGray.data; }
const movfe_hop *ref_clip_hops(const ref_clip *c, int f) { return c->hops[f].data(); }
const movfe_rect *ref_clip_kps(const ref_clip *c, int f) { return c->kps[f].data(); }

// ------------------------------------------------------------------------------------------------------- EXPRESS -----
int ref_express_center(const uint8_t *img, int stride, int x0, int y0, int cols, int rows) {
    cv::Mat m = view_u8(img, stride, x0, y0, cols, rows);
    return compute_center(m);
}
void ref_express_descriptor(const uint8_t *img, int stride, int x0, int y0, int cols, int rows, int threshold, uint32_t desc[8]) {
    cv::Mat m = view_u8(img, stride, x0, y0, cols, rows);
    std::bitset<256> d;
    compute_descriptor(m, threshold, d);
    for (int k = 0; k < 8; k++) desc[k] = 0;
    for (int b = 0; b < 256; b++)
        if (d[b]) desc[b >> 5] |= 1u << (b & 31);
}
int ref_express_test(const uint8_t *img, int stride, int x0, int y0, int cols, int rows, int threshold) {
    cv::Mat m = view_u8(img, stride, x0, y0, cols, rows);
    return compute_express(m, threshold) ? 1 : 0;
}
int ref_express_distance(const uint32_t a[8], const uint32_t b[8]) {
    std::bitset<256> x, y;
    for (int i = 0; i < 256; i++) {
        x[i] = (a[i >> 5] >> (i & 31)) & 1u;
        y[i] = (b[i >> 5] >> (i & 31)) & 1u;
    }
    return compute_distance(x, y);
}

// -------------------------------------------------------------------------------------------------- MOVExtractor -----
// Results of the next cv::calcOpticalFlowPyrLK calls made by the reference (one entry per call, in call order).
void ref_lk_reset(void) {
    cvstub::lk_queue().clear();
    cvstub::lk_calls() = 0;
}
void ref_lk_push(const uint8_t *status, const float *pts_xy, int n) {
    cvstub::LkCall c;
    for (int i = 0; i < n; i++) {
        c.status.push_back(status[i]);
        c.pts.push_back(cv::Point2f(pts_xy[2 * i], pts_xy[2 * i + 1]));
    }
    cvstub::lk_queue().push_back(c);
}
int ref_lk_calls(void) { return cvstub::lk_calls(); }
// points the reference handed to its last LK call (so a test can check which features it wanted carried); returns the count
int ref_lk_last_points(float *pts_xy, int capacity) {
    const auto &p = cvstub::lk_last_prev_pts();
    for (size_t i = 0; i < p.size() && (int)i < capacity; i++) {
        pts_xy[2 * i] = p[i].x;
        pts_xy[2 * i + 1] = p[i].y;
    }
    return (int)p.size();
}

typedef struct ref_extract_aux {
    int32_t n_keypoints;     /* return value of operator() */
    int32_t n_descriptors;   /* descriptors.size() after the call (smaller than the table on back-fill frames, :421) */
    int32_t consistent;      /* 1: vf.dIndx == own index, keypoint == (pt, mb.width) and vfmap == first index per id, for every entry */
    int32_t lk_calls;        /* cv::calcOpticalFlowPyrLK calls made */
} ref_extract_aux;

// One MOVExtractor::operator() call (src/MOVExtractor.cc:63-455). prev/n_prev: prev->mvVF (sorted IN PLACE by the
// reference, :249-252, and written back); lost-relocalisation inputs (prev->mLost, :161-243): n_kf points of the reference
// keyframe's map-point list with {mbTrackInView, mTrackProjX/Y, mTrackId} = kf_in_view / kf_proj_xy / kf_track_id.
int ref_extract_frame(int width, int height, uint32_t frame_flags, const uint8_t *grey, const int32_t *grid, const movfe_hop *hops,
                      int n_hops, const movfe_rect *kps, int n_kps, double coverage_area, movfe_track *prev, int n_prev, int has_prev,
                      int lost, int n_kf, const uint8_t *kf_in_view, const float *kf_proj_xy, const int32_t *kf_track_id,
                      int threshold, double coverage_threshold, double relocalization_distance, int32_t *current_id, movfe_track *out,
                      int capacity, ref_extract_aux *aux) {
    using namespace MOV_SLAM;
    std::shared_ptr<MotionVectorImage> smv(new MotionVectorImage(width, height));
    smv->ft = (frame_flags & MOVFE_FRAME_P) ? FrameType::P_FRAME : FrameType::I_FRAME;
    smv->frame = 0;
    smv->coverageArea = coverage_area;
    if (grey) smv->imGray = cv::Mat(height, width, CV_8UC1, const_cast<uint8_t *>(grey), (size_t)width);
    if (grid) std::memcpy(smv->mvi.data, grid, (size_t)width * height * 16);
    for (int i = 0; i < n_hops; i++) {
        MotionVector mv;
        mv.occupied = true;
        mv.pt = cv::Point2f(hops[i].mv_x, hops[i].mv_y);
        mv.dIndx = hops[i].d_indx;
        smv->mvs.push_back(mv);
    }
    for (int i = 0; i < n_kps; i++) smv->kps.push_back(cv::Rect(kps[i].x, kps[i].y, kps[i].w, kps[i].h));

    Frame prevF;
    prevF.imageCols = width;
    prevF.imageRows = height;
    prevF.mLost = lost != 0;
    for (int i = 0; i < n_prev; i++) prevF.mvVF.push_back(to_vf(prev[i], i));
    KeyFrame kf;
    std::vector<MapPoint> mps(n_kf > 0 ? n_kf : 0);
    for (int i = 0; i < n_kf; i++) {
        mps[i].mbTrackInView = kf_in_view[i] != 0;
        mps[i].mTrackProjX = kf_proj_xy[2 * i];
        mps[i].mTrackProjY = kf_proj_xy[2 * i + 1];
        mps[i].mTrackId = kf_track_id[i];
        kf.mvpMapPoints.push_back(&mps[i]);
    }
    prevF.mpReferenceKF = &kf;

    MOVExtractor ext(threshold, coverage_threshold, relocalization_distance);
    ext.mCurrentId = *current_id;
    std::vector<cv::KeyPoint> keypoints;
    std::vector<VideoFeature> vf;
    std::map<int, int> vfmap;
    std::vector<std::bitset<256>> descriptors;
    const int calls0 = cvstub::lk_calls();
    const int ret = ext(smv, keypoints, vf, vfmap, descriptors, has_prev ? &prevF : nullptr);
    *current_id = ext.mCurrentId;
    for (int i = 0; i < n_prev; i++) prev[i] = to_track(prevF.mvVF[i]);

    bool ok = ret == (int)keypoints.size() && keypoints.size() == vf.size();
    std::map<int, int> first;
    for (size_t i = 0; i < vf.size() && ok; i++) {
        ok = ok && vf[i].dIndx == (int)i && keypoints[i].pt.x == vf[i].pt.x && keypoints[i].pt.y == vf[i].pt.y &&
             keypoints[i].size == (float)vf[i].mb.width;
        first.insert({vf[i].trackId, (int)i});
    }
    ok = ok && first == vfmap;
    if (aux) {
        aux->n_keypoints = ret;
        aux->n_descriptors = (int)descriptors.size();
        aux->consistent = ok ? 1 : 0;
        aux->lk_calls = cvstub::lk_calls() - calls0;
    }
    const int n = (int)vf.size();
    for (int i = 0; i < n && i < capacity; i++) out[i] = to_track(vf[i]);
    return n;
}

// The reference's own per-frame loop over a clip, as its mains drive it (Examples/Monocular/mono_video_tartan.cc:71-100 up to the
// Frame constructor): VideoDecoder::NextImage -> MOVExtractor::operator() with the previous frame's table. LK returns "lost" for
// every point (nothing injected). Timed inside with steady_clock like the mains do (:73-86). checksums: per frame, the table
// checksum of oracle/frontend.cc (recomputed here over the converted tables). Returns the number of frames processed.
int ref_frontend_run(int width, int height, int n_frames, const movfe_mv_record *recs, const int64_t *rec_off, const uint8_t *frame_flags,
                     const uint8_t *grey, int qlen, int threshold, double coverage_threshold, double *seconds_decoder, double *seconds_extractor,
                     int32_t *n_tracks, uint64_t *checksums) {
    using namespace MOV_SLAM;
    using clk = std::chrono::steady_clock;
    std::vector<uint8_t> is_p(n_frames);
    std::vector<const uint8_t *> luma(n_frames, nullptr), side(n_frames, nullptr);
    std::vector<int> side_bytes(n_frames, 0);
    for (int f = 0; f < n_frames; f++) {
        is_p[f] = (frame_flags[f] & MOVFE_FRAME_P) ? 1 : 0;
        if (grey) luma[f] = grey + (size_t)f * width * height;
        if ((frame_flags[f] & MOVFE_FRAME_MV) && rec_off[f + 1] > rec_off[f]) {
            side[f] = reinterpret_cast<const uint8_t *>(recs + rec_off[f]);
            side_bytes[f] = (int)((rec_off[f + 1] - rec_off[f]) * (int64_t)sizeof(movfe_mv_record));
        }
    }
    fake_av_clip fc = {width, height, n_frames, is_p.data(), luma.data(), side.data(), side_bytes.data()};
    fake_av_install(&fc);
    VideoDecoder dec("fake://clip", qlen);
    if (!dec.Init()) return -1;
    MOVExtractor ext(threshold, coverage_threshold, 0.25);
    ext.mCurrentId = 0;
    cvstub::lk_queue().clear();
    std::shared_ptr<Frame> prev;
    double t_dec = 0, t_ext = 0;
    int done = 0;
    while (true) {
        const auto a = clk::now();
        std::shared_ptr<MotionVectorImage> smv = dec.NextImage(true);
        const auto b = clk::now();
        t_dec += std::chrono::duration<double>(b - a).count();
        if (!smv) break;
        std::shared_ptr<Frame> F(new Frame());
        F->imageCols = width;
        F->imageRows = height;
        F->imgLeft = smv->imGray;
        const auto c = clk::now();
        F->N = ext(smv, F->mvKeys, F->mvVF, F->mvVFMap, F->mDescriptors, prev.get());
        t_ext += std::chrono::duration<double>(clk::now() - c).count();
        if (n_tracks) n_tracks[done] = (int32_t)F->mvVF.size();
        if (checksums) {
            std::vector<movfe_track> tab;
            for (const auto &vf : F->mvVF) tab.push_back(to_track(vf));
            const uint64_t *w = reinterpret_cast<const uint64_t *>(tab.data());
            uint64_t h = 0;
            for (size_t i = 0; i < tab.size() * 8; i++) h += (w[i] ^ (i * 0x9E3779B97F4A7C15ull)) * (2 * i + 1);
            checksums[done] = h;
        }
        prev = F;
        done++;
    }
    if (seconds_decoder) *seconds_decoder = t_dec;
    if (seconds_extractor) *seconds_extractor = t_ext;
    return done;
}

// ---------------------------------------------------------------------------------------------------- MOVMatcher -----
static void fill_frame(MOV_SLAM::Frame &F, const movfe_track *tracks, int n) {
    F.N = n;
    for (int i = 0; i < n; i++) {
        F.mvVF.push_back(to_vf(tracks[i], i));
        F.mvVFMap.insert(std::pair<int, int>(tracks[i].track_id, i));  // as MOVExtractor.cc:117,151,237,330,374,411,448
        cv::KeyPoint kp(cv::Point2f(tracks[i].pt_x, tracks[i].pt_y), (float)tracks[i].mb.w);
        F.mvKeys.push_back(kp);
        F.mvKeysUn.push_back(kp);
    }
    F.mvpMapPoints.assign(n, nullptr);
}

// MOVMatcher::SearchByVideoFeature(Frame&, const vector<MapPoint*>&, bFarPoints, thFarPoints) (MOVMatcher.h:35-68).
// match (in/out) mirrors F.mvpMapPoints as indices into pts (-1 = NULL).
int ref_search_by_video_feature(const movfe_track *tracks, int n_tracks, const movfe_map_point *pts, const movfe_projection *proj,
                                int n_pts, int far_points, float th_far, int32_t *match) {
    using namespace MOV_SLAM;
    Frame F;
    fill_frame(F, tracks, n_tracks);
    std::vector<MapPoint> mps(n_pts);
    std::vector<MapPoint *> v;
    for (int i = 0; i < n_pts; i++) {
        mps[i].mTrackId = pts[i].track_id;
        mps[i].mbBad = (pts[i].flags & MOVFE_MP_BAD) != 0;
        mps[i].mbTrackInView = proj[i].in_view != 0;
        mps[i].mTrackDepth = proj[i].depth;
        v.push_back(&mps[i]);
    }
    for (int i = 0; i < n_tracks; i++) F.mvpMapPoints[i] = match[i] >= 0 ? &mps[match[i]] : nullptr;
    const int n = MOVMatcher::SearchByVideoFeature(F, v, far_points != 0, th_far);
    for (int i = 0; i < n_tracks; i++) match[i] = F.mvpMapPoints[i] ? (int32_t)(F.mvpMapPoints[i] - mps.data()) : -1;
    return n;
}

// MOVMatcher::SearchByVideoFeature(KeyFrame*, Frame&, vector<MapPoint*>&) (MOVMatcher.h:70-103)
int ref_search_by_keyframe(const movfe_track *tracks, int n_tracks, const movfe_map_point *kf_pts, int n_pts, int32_t *match) {
    using namespace MOV_SLAM;
    Frame F;
    fill_frame(F, tracks, n_tracks);
    std::vector<MapPoint> mps(n_pts);
    KeyFrame kf;
    for (int i = 0; i < n_pts; i++) {
        mps[i].mTrackId = kf_pts[i].track_id;
        mps[i].mbBad = (kf_pts[i].flags & MOVFE_MP_BAD) != 0;
        kf.mvpMapPoints.push_back((kf_pts[i].flags & MOVFE_MP_NULL) ? nullptr : &mps[i]);
    }
    std::vector<MapPoint *> out;
    const int n = MOVMatcher::SearchByVideoFeature(&kf, F, out);
    for (int i = 0; i < n_tracks; i++) match[i] = out[i] ? (int32_t)(out[i] - mps.data()) : -1;
    return n;
}

// MOVMatcher::SearchForInitialization (MOVMatcher.h:105-137)
int ref_search_for_initialization(const movfe_track *f1, int n1, const movfe_track *f2, int n2, float *prev_matched, int32_t *matches12) {
    using namespace MOV_SLAM;
    Frame F1, F2;
    fill_frame(F1, f1, n1);
    fill_frame(F2, f2, n2);
    std::vector<cv::Point2f> pm(n1);
    for (int i = 0; i < n1; i++) pm[i] = cv::Point2f(prev_matched[2 * i], prev_matched[2 * i + 1]);
    std::vector<int> m12;
    const int n = MOVMatcher::SearchForInitialization(F1, F2, pm, m12, 100);
    for (int i = 0; i < n1; i++) {
        matches12[i] = m12[i];
        prev_matched[2 * i] = pm[i].x;
        prev_matched[2 * i + 1] = pm[i].y;
    }
    return n;
}

}  // extern "C"
