// oracle/frontend.cc — TEST INFRASTRUCTURE (see oracle.h).
// Drives the oracle's stages over one stream in the order Tracking.cc drives the reference's:
//   VideoDecoder::NextImage (raster)                         mono_video_tartan.cc:74
//   Frame ctor -> MOVExtractor::operator()                   Tracking.cc:178-186, Frame.cc:121-123
//   TrackReferenceKeyFrame: join(KF) -> pose := last -> PoseOptimization      Tracking.cc:796-811
//   TrackLocalMap: SearchLocalPoints (frustum + join) -> PoseOptimization     Tracking.cc:890-905, 1109-1158
// Used as the timed CPU baseline and as the checker for the batched GPU pipeline.
#include "oracle.h"

#include <chrono>
#include <cstring>
#include <vector>

// Order-sensitive checksum of a track table viewed as 64-bit words: sum of (w[i] ^ i*K) * (2i+1) mod 2^64. Chosen over a
// byte-serial hash because bench.py recomputes it over the GPU's tables with three vectorised numpy operations.
static uint64_t table_checksum(const void *data, size_t n_bytes) {
    const uint64_t *w = (const uint64_t *)data;
    uint64_t h = 0;
    for (size_t i = 0; i < n_bytes / 8; i++) h += (w[i] ^ (i * 0x9E3779B97F4A7C15ull)) * (2 * i + 1);
    return h;
}

extern "C" int orc_frontend_run(const orc_frontend_cfg *cfg, const movfe_mv_record *recs, const int64_t *rec_off,
                                const uint8_t *frame_flags, const uint8_t *grey, const movfe_track *seed_tracks,
                                int n_seed, const movfe_map_point *map_pts, int n_map, const movfe_pose *pose0,
                                orc_frontend_out *out) {
    return orc_frontend_run_sched(cfg, recs, rec_off, frame_flags, grey, seed_tracks, n_seed, map_pts, n_map, pose0, 0, nullptr,
                                  nullptr, nullptr, nullptr, 0, nullptr, out);
}

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

extern "C" int orc_frontend_run_sched(const orc_frontend_cfg *cfg, const movfe_mv_record *recs, const int64_t *rec_off,
                                      const uint8_t *frame_flags, const uint8_t *grey, const movfe_track *seed_tracks,
                                      int n_seed, const movfe_map_point *map_pts, int n_map, const movfe_pose *pose0,
                                      int n_sched, const int32_t *sched_frame, const int64_t *sched_off,
                                      const movfe_map_point *sched_pts, const int32_t *sched_nkf, int timed_from,
                                      double *tail_times, orc_frontend_out *out) {
    const int W = cfg->width, H = cfg->height, NF = cfg->n_frames;
    int n_kf_points = cfg->n_kf_points;
    int next_sched = 0;
    orc_clip *clip = orc_raster_clip(W, H, NF, recs, rec_off, frame_flags, cfg->max_ref);

    std::vector<uint8_t> flat;
    if (!grey) flat.assign((size_t)W * H, 128);

    orc_extract_params ep;
    ep.threshold = cfg->threshold;
    ep.coverage_threshold = cfg->coverage_threshold;
    ep.max_tracks = cfg->max_tracks;

    std::vector<movfe_track> prev(cfg->max_tracks), cur(cfg->max_tracks);
    int n_prev = 0;
    if (seed_tracks && n_seed > 0) {
        n_prev = n_seed < cfg->max_tracks ? n_seed : cfg->max_tracks;
        std::memcpy(prev.data(), seed_tracks, sizeof(movfe_track) * n_prev);
    }
    int32_t current_id = 0;
    for (int i = 0; i < n_prev; i++)
        if (prev[i].track_id > current_id) current_id = prev[i].track_id;

    movfe_pose pose = *pose0;
    std::vector<int32_t> match(cfg->max_tracks);
    std::vector<movfe_map_point> pts(map_pts, map_pts + n_map);
    std::vector<movfe_projection> proj(n_map);
    std::vector<float> gx, go;
    std::vector<uint8_t> outl;

    for (int f = 0; f < NF; f++) {
        if (tail_times && f == timed_from) tail_times[0] = now_s();
        // the local map handed over by the mapping side before this frame (keyframe insertion -> UpdateLocalPoints,
        // Tracking.cc:947-1107,1171-1198: out of scope, so its result arrives as a schedule)
        while (next_sched < n_sched && sched_frame[next_sched] <= f) {
            if (sched_frame[next_sched] == f) {
                pts.assign(sched_pts + sched_off[next_sched], sched_pts + sched_off[next_sched + 1]);
                n_map = (int)pts.size();
                n_kf_points = sched_nkf[next_sched];
                proj.resize(n_map);
            }
            next_sched++;
        }
        const uint8_t *img = grey ? grey + (size_t)f * W * H : flat.data();
        // seed_tracks (if any) are the table of the frame before the clip (Frame::mpPrevFrame of frame 0)
        const int n = orc_extract_frame(W, H, frame_flags[f], img, orc_clip_grid(clip, f), orc_clip_hops(clip, f),
                                        orc_clip_kps(clip, f), orc_clip_n_kps(clip, f), orc_clip_coverage(clip, f),
                                        prev.data(), n_prev, nullptr, nullptr, &ep, &current_id, cur.data(), nullptr);
        int n_inl = 0;
        if (n_map > 0 && n > 0) {
            auto gather = [&]() {
                gx.clear();
                go.clear();
                for (int i = 0; i < n; i++)
                    if (match[i] >= 0) {  // Optimizer.cc:404-413
                        const movfe_map_point &mp = pts[match[i]];
                        gx.insert(gx.end(), mp.pos, mp.pos + 3);
                        go.push_back(cur[i].pt_x);
                        go.push_back(cur[i].pt_y);
                    }
                outl.assign(go.size() / 2 + 1, 0);
                return (int)(go.size() / 2);
            };
            // TrackReferenceKeyFrame
            orc_search_by_keyframe(cur.data(), n, pts.data(), n_kf_points < n_map ? n_kf_points : n_map, match.data());
            int P = gather();
            orc_pose_optimize(&cfg->cam, &cfg->pose_params, gx.data(), go.data(), P, &pose, outl.data(), nullptr);
            // TrackLocalMap / SearchLocalPoints
            for (int k = 0; k < n_map; k++) pts[k].flags &= ~MOVFE_MP_SKIP;
            for (int i = 0; i < n; i++)
                if (match[i] >= 0) pts[match[i]].flags |= MOVFE_MP_SKIP;  // Tracking.cc:1112-1128
            orc_frustum(&pose, &cfg->cam, W, H, cfg->viewing_cos_limit, pts.data(), n_map, proj.data());
            orc_search_by_video_feature(cur.data(), n, pts.data(), proj.data(), n_map, 0, 0.f, match.data());
            P = gather();
            n_inl = orc_pose_optimize(&cfg->cam, &cfg->pose_params, gx.data(), go.data(), P, &pose, outl.data(), nullptr);
        }
        if (out) {
            if (out->poses) out->poses[f] = pose;
            if (out->n_tracks) out->n_tracks[f] = n;
            if (out->n_inliers) out->n_inliers[f] = n_inl;
            if (out->track_hash) out->track_hash[f] = table_checksum(cur.data(), sizeof(movfe_track) * (size_t)n);
        }
        prev.swap(cur);
        n_prev = n;
    }
    if (tail_times) tail_times[1] = now_s();
    if (out && out->last_tracks) std::memcpy(out->last_tracks, prev.data(), sizeof(movfe_track) * (size_t)n_prev);
    orc_clip_free(clip);
    return n_prev;
}
