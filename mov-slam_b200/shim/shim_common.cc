// shim_common.cc — contexts and record conversions shared by the drop-in shims.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <tuple>

#include "movfe_shim.h"

namespace movfe_shim {

void fail(movfe_ctx *ctx, const char *what) { fprintf(stderr, "movfe: %s: %s\n", what, movfe_last_error(ctx)); }

static movfe_ctx *create(const movfe_config &cfg, const char *what) {
    movfe_ctx *ctx = nullptr;
    if (movfe_create(&cfg, &ctx) != MOVFE_OK) {
        fail(nullptr, what);
        return nullptr;  // there is no CPU fallback: the caller reports failure the way the reference does
    }
    return ctx;
}

movfe_ctx *extractor_context(int width, int height, int threshold, double coverage_threshold, bool has_grey) {
    static std::map<std::tuple<int, int, int, double, bool>, movfe_ctx *> cache;
    const auto key = std::make_tuple(width, height, threshold, coverage_threshold, has_grey);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    movfe_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.n_streams = 1;
    cfg.width = width;
    cfg.height = height;
    cfg.max_records_per_frame = ((width + 15) / 16) * ((height + 15) / 16) * 4;  // H.264: 4 records per macroblock
    cfg.max_ref = 10;  // the reference's 12-deep queue (VideoDecoder.cc:163)
    cfg.window_frames = 1;
    cfg.max_tracks = 8192;
    cfg.max_map_points = 1;
    cfg.express_threshold = threshold;
    cfg.coverage_threshold = coverage_threshold;
    cfg.has_grey = has_grey;
    return cache[key] = create(cfg, "extractor context");
}

movfe_ctx *operator_context() {
    static movfe_ctx *ctx = nullptr;
    if (ctx) return ctx;
    movfe_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.n_streams = 1;
    cfg.width = cfg.height = 16;
    cfg.max_records_per_frame = 1;
    cfg.window_frames = 1;
    cfg.max_tracks = 1;
    cfg.max_map_points = 1;
    cfg.express_threshold = 20;
    return ctx = create(cfg, "operator context");
}

// operators that need the frame size (the bucket grid's cell size): geometry-only contexts, one per size
movfe_ctx *frame_operator_context(int width, int height) {
    static std::map<std::pair<int, int>, movfe_ctx *> cache;
    const auto key = std::make_pair(width, height);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    movfe_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.n_streams = 1;
    cfg.width = width;
    cfg.height = height;
    cfg.max_records_per_frame = 1;
    cfg.window_frames = 1;
    cfg.max_tracks = 1;
    cfg.max_map_points = 1;
    cfg.express_threshold = 20;
    return cache[key] = create(cfg, "frame operator context");
}

movfe_track pack(const MOV_SLAM::VideoFeature &vf) {
    movfe_track t;
    memset(&t, 0, sizeof t);
    t.pt_x = vf.pt.x;
    t.pt_y = vf.pt.y;
    t.mb = {(int16_t)vf.mb.x, (int16_t)vf.mb.y, (int16_t)vf.mb.width, (int16_t)vf.mb.height};
    t.track_id = vf.trackId;
    t.age = vf.age;
    t.q_indx = vf.qIndx;
    t.flags = vf.coverage ? MOVFE_TRACK_COVERAGE : 0u;
    for (int i = 0; i < 256; i++)
        if (vf.desc[i]) t.desc[i >> 5] |= 1u << (i & 31);
    return t;
}

MOV_SLAM::VideoFeature unpack(const movfe_track &t, int index) {
    MOV_SLAM::VideoFeature vf;
    vf.trackId = t.track_id;
    vf.qIndx = t.q_indx;
    vf.dIndx = index;  // index of the matching cv::KeyPoint == position in the table (MOVExtractor.cc:318-331)
    vf.pt = cv::Point2f(t.pt_x, t.pt_y);
    vf.mb = cv::Rect(t.mb.x, t.mb.y, t.mb.w, t.mb.h);
    vf.age = t.age;
    vf.coverage = (t.flags & MOVFE_TRACK_COVERAGE) != 0;
    for (int i = 0; i < 256; i++) vf.desc[i] = (t.desc[i >> 5] >> (i & 31)) & 1u;
    return vf;
}

movfe_camera pack(MOV_SLAM::GeometricCamera *cam) {
    movfe_camera c;
    memset(&c, 0, sizeof c);
    c.model = cam->GetType() == MOV_SLAM::GeometricCamera::CAM_FISHEYE ? MOVFE_CAM_FISHEYE : MOVFE_CAM_PINHOLE;
    c.fx = cam->getParameter(0);
    c.fy = cam->getParameter(1);
    c.cx = cam->getParameter(2);
    c.cy = cam->getParameter(3);
    if (c.model == MOVFE_CAM_FISHEYE && cam->size() >= 8)
        for (int i = 0; i < 4; i++) c.k[i] = cam->getParameter(4 + i);
    return c;
}

// ------------------------------------------------------------------------------------------------ RasterQueue ----
struct RasterQueue::Impl {
    movfe_ctx *ctx = nullptr;
    int W = 0, H = 0, K = 0;
    int64_t pushed = 0, popped = 0;
    std::deque<std::shared_ptr<MOV_SLAM::MotionVectorImage>> pending;
};

RasterQueue::RasterQueue(int width, int height, int max_ref, int max_records_per_frame) : d(new Impl) {
    movfe_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.n_streams = 1;
    cfg.width = width;
    cfg.height = height;
    cfg.max_records_per_frame = max_records_per_frame > 0 ? max_records_per_frame : ((width + 15) / 16) * ((height + 15) / 16) * 4;
    cfg.max_ref = max_ref;
    cfg.window_frames = 1;
    cfg.max_tracks = 1;
    cfg.max_map_points = 1;
    cfg.express_threshold = 20;
    d->ctx = create(cfg, "raster context");
    d->W = width;
    d->H = height;
    d->K = max_ref;
}

RasterQueue::~RasterQueue() {
    movfe_destroy(d->ctx);
    delete d;
}

bool RasterQueue::push(const std::shared_ptr<MOV_SLAM::MotionVectorImage> &img, const void *side_data, int n_records, bool mv) {
    if (!d->ctx) return false;
    const int64_t off[2] = {0, n_records};
    uint8_t flags = (img->ft == MOV_SLAM::FrameType::P_FRAME ? MOVFE_FRAME_P : 0u) | (mv && n_records > 0 ? MOVFE_FRAME_MV : 0u);
    if (getenv("MOVFE_SHIM_DEBUG") && n_records > 0) {
        const movfe_mv_record *r = (const movfe_mv_record *)side_data;
        fprintf(stderr, "shim push: n=%d flags=%u first rec: src=%d w=%d h=%d s=(%d,%d) d=(%d,%d) ref=%d\n", n_records, flags, r->source, r->w, r->h,
                r->src_x, r->src_y, r->dst_x, r->dst_y, r->ref);
    }
    if (movfe_push_frames(d->ctx, 1, (const movfe_mv_record *)side_data, off, &flags, nullptr) != MOVFE_OK) {
        fail(d->ctx, "push_frames");
        return false;
    }
    // side_data belongs to the AVFrame and is gone after av_frame_unref: wait for the host->device copy before returning
    // (one frame of one stream per call - the batched harness keeps its buffers pinned and never waits here)
    movfe_synchronize(d->ctx);
    img->frame = (int)d->pushed;
    d->pending.push_back(img);
    d->pushed++;
    return true;
}

int RasterQueue::pending() const { return (int)d->pending.size(); }

std::shared_ptr<MOV_SLAM::MotionVectorImage> RasterQueue::pop(bool flush) {
    if (!d->ctx || d->pending.empty()) return nullptr;
    // frame f is final once frames f+1 .. f+K+1 were pushed (records with ref r back-fill frames up to r+1 earlier)
    if (!flush && d->pushed - d->popped < d->K + 2) return nullptr;
    const int64_t f = d->popped;
    if (movfe_raster(d->ctx, f, 1) != MOVFE_OK) {
        fail(d->ctx, "raster");
        return nullptr;
    }
    std::shared_ptr<MOV_SLAM::MotionVectorImage> img = d->pending.front();
    d->pending.pop_front();
    d->popped++;
    int32_t nh = 0, nk = 0;
    const int rc_counts = movfe_raster_counts(d->ctx, 0, f, &nh, &nk, &img->coverageArea);
    if (getenv("MOVFE_SHIM_DEBUG"))
        fprintf(stderr, "shim raster: frame %lld rc=%d hops=%d kps=%d cov=%f pushed=%lld rejected=%lld\n", (long long)f, rc_counts, nh, nk,
                img->coverageArea, (long long)d->pushed, (long long)movfe_rejected_records(d->ctx));
    std::vector<movfe_hop> hops((size_t)std::max(nh, 1));
    std::vector<movfe_rect> kps((size_t)std::max(nk, 1));
    movfe_download_hops(d->ctx, 0, f, hops.data(), (int)hops.size());
    movfe_download_kps(d->ctx, 0, f, kps.data(), (int)kps.size());
    movfe_download_grid(d->ctx, 0, f, reinterpret_cast<int32_t *>(img->mvi.data));  // CV_32SC4, continuous (Frame.h:123)
    img->mvs.clear();
    img->kps.clear();
    for (int i = 0; i < nh; i++) {
        MOV_SLAM::MotionVector m;
        m.pt = cv::Point2f(hops[i].mv_x, hops[i].mv_y);
        m.dIndx = hops[i].d_indx;
        img->mvs.push_back(m);
    }
    for (int i = 0; i < nk; i++) img->kps.push_back(cv::Rect(kps[i].x, kps[i].y, kps[i].w, kps[i].h));
    return img;
}

}  // namespace movfe_shim
