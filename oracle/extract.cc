// oracle/extract.cc — TEST INFRASTRUCTURE (see oracle.h).
// CPU restatement of MOVExtractor::operator(), src/MOVExtractor.cc:63-455, for one frame of one stream.
// cv::calcOpticalFlowPyrLK (:91,:196,:347) is third-party arithmetic that stays on the host: its results
// enter through lk_status/lk_pts; with NULL every carried feature is dropped. The lost-relocalisation
// branch (:161-243) prepends features at LK-carried keyframe points: the LK call and the status / bounds /
// distance tests on its output (:194-215) are the host's, the block + descriptor part (:218-238) is restated here.
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

inline int popcount256(const uint32_t d[8]) {
    int c = 0;
    for (int i = 0; i < 8; i++) c += __builtin_popcount(d[i]);
    return c;
}

struct Out {
    movfe_track *out;
    int cap;
    int n = 0;  // logical size (may exceed cap; entries beyond cap are dropped)
    void push(const movfe_track &t) {
        if (n < cap) out[n] = t;
        n++;
    }
};

inline bool in_bounds(int x, int y, int w, int h, int cols, int rows) {
    return x >= 0 && y >= 0 && (x + w) < cols && (y + h) < rows;
}

}  // namespace

extern "C" int orc_extract_frame(int width, int height, uint32_t frame_flags, const uint8_t *grey,
                                 const int32_t *grid, const movfe_hop *hops, const movfe_rect *kps, int n_kps,
                                 double coverage_area, movfe_track *prev, int n_prev, const uint8_t *lk_status,
                                 const float *lk_pts, const orc_extract_params *params, int32_t *current_id,
                                 movfe_track *out_tracks, int32_t *n_births) {
    return orc_extract_frame_lost(width, height, frame_flags, grey, grid, hops, kps, n_kps, coverage_area, prev, n_prev, lk_status,
                                  lk_pts, nullptr, 0, params, current_id, out_tracks, n_births);
}

extern "C" int orc_extract_frame_lost(int width, int height, uint32_t frame_flags, const uint8_t *grey,
                                      const int32_t *grid, const movfe_hop *hops, const movfe_rect *kps, int n_kps,
                                      double coverage_area, movfe_track *prev, int n_prev, const uint8_t *lk_status,
                                      const float *lk_pts, const movfe_reloc_seed *reloc, int n_reloc,
                                      const orc_extract_params *params, int32_t *current_id,
                                      movfe_track *out_tracks, int32_t *n_births) {
    if (!grey) return -1;  // :71-72 imGray.empty()
    const int cols = width, rows = height;
    const int thr = params->threshold;
    Out out{out_tracks, params->max_tracks};
    std::vector<bool> lbFound(n_kps, false);  // :68
    int mov_cnt = 0;
    int mCurrentId = *current_id;

    auto slot = [&](int y, int x, int j) -> int32_t { return grid[((size_t)y * cols + x) * 4 + j]; };
    auto carried = [&](const movfe_track &pvf, int i, float px, float py, bool coverage) {
        movfe_track vf;  // :102-113 / :362-375
        vf.pt_x = px;
        vf.pt_y = py;
        vf.mb = pvf.mb;
        vf.track_id = pvf.track_id;
        vf.age = pvf.age + 1;
        vf.q_indx = i;
        vf.flags = coverage ? MOVFE_TRACK_COVERAGE : 0;
        std::memcpy(vf.desc, pvf.desc, 32);
        out.push(vf);
    };

    if (!(frame_flags & MOVFE_FRAME_P)) {  // :79 I_FRAME
        if (prev && n_prev > 0) {          // :81-120 LK carry-over of every previous feature
            for (int i = 0; i < n_prev; i++) {
                if (!lk_status || lk_status[i] == 0) continue;
                const float x = lk_pts[2 * i], y = lk_pts[2 * i + 1];
                if (x < 0 || y < 0 || x >= cols || y >= rows) continue;  // :97
                carried(prev[i], i, x, y, false);
            }
        } else {  // :123-157 seeding on the 16-px lattice
            for (int y = 8; y < rows - 8; y += 16) {
                for (int x = 8; x < cols - 8; x += 16) {
                    const int mx = x - 8, my = y - 8;
                    if (in_bounds(mx, my, 16, 16, cols, rows)) {
                        if (orc_express_test(grey, cols, mx, my, 16, 16, thr)) {
                            movfe_track vf;
                            orc_express_descriptor(grey, cols, mx, my, 16, 16, thr, vf.desc);
                            mCurrentId++;
                            vf.pt_x = (float)x;
                            vf.pt_y = (float)y;
                            vf.mb = {(int16_t)mx, (int16_t)my, 16, 16};
                            vf.track_id = mCurrentId;
                            vf.age = 0;
                            vf.q_indx = -1;
                            vf.flags = 0;
                            out.push(vf);
                        }
                    }
                }
            }
        }
    } else {
        // Lost relocalisation (:161-243): the seeds already passed status / bounds / distance on the host (:207-215)
        for (int i = 0; i < n_reloc; i++) {
            const float x = reloc[i].x, y = reloc[i].y;
            const int mx = (int)(x - 8), my = (int)(y - 8);  // :218 cv::Rect(float...) truncates
            if (in_bounds(mx, my, 16, 16, cols, rows)) {      // :219
                movfe_track vf;
                orc_express_descriptor(grey, cols, mx, my, 16, 16, thr, vf.desc);  // :221-223
                vf.pt_x = x;
                vf.pt_y = y;
                vf.mb = {(int16_t)mx, (int16_t)my, 16, 16};
                vf.track_id = reloc[i].track_id;
                vf.age = 0;
                vf.q_indx = reloc[i].q_indx;
                vf.flags = 0;
                out.push(vf);
            }
        }
        // Project forward the previous frame keypoints (:246-335)
        std::vector<int> covFeat;  // indices (in sorted order) of coverage features
        if (n_prev > 0) {
            // :249-252, canonicalised to a stable sort (std::sort leaves ties implementation-defined)
            std::stable_sort(prev, prev + n_prev, [](const movfe_track &a, const movfe_track &b) {
                return a.age == b.age ? popcount256(a.desc) > popcount256(b.desc) : a.age > b.age;
            });

            for (int i = 0; i < n_prev; i++) {
                const movfe_track &pvf = prev[i];

                if (pvf.flags & MOVFE_TRACK_COVERAGE) {  // :258-262
                    covFeat.push_back(i);
                    continue;
                }

                const int x = (int)pvf.pt_x, y = (int)pvf.pt_y;  // :264
                if (slot(y, x, 0) == -1) continue;                // :265-268

                int indx = slot(y, x, 0);  // :270

                if (slot(y, x, 1) >= 0) {  // :272
                    int bestDesc = 256;
                    for (int j = 0; j < 4; j++) {
                        if (slot(y, x, j) == -1) break;  // :277-278
                        const movfe_hop &mv = hops[slot(y, x, j)];

                        // Shift to the destination (:283-284)
                        const float px = pvf.pt_x + mv.mv_x, py = pvf.pt_y + mv.mv_y;
                        const int mx = (int)(px - (pvf.mb.w / 2)), my = (int)(py - (pvf.mb.h / 2));
                        const int mw = pvf.mb.w, mh = pvf.mb.h;

                        if (in_bounds(mx, my, mw, mh, cols, rows)) {  // :286
                            uint32_t desc[8];
                            orc_express_descriptor(grey, cols, mx, my, mw, mh, thr, desc);
                            const int dist = orc_express_distance(pvf.desc, desc);
                            if (dist < bestDesc) {  // :292-296
                                bestDesc = dist;
                                indx = slot(y, x, j);
                            }
                        }
                    }
                }

                const movfe_hop &mv = hops[indx];  // :301

                const float px = pvf.pt_x + mv.mv_x, py = pvf.pt_y + mv.mv_y;  // :303-304
                const int mx = (int)(px - (pvf.mb.w / 2)), my = (int)(py - (pvf.mb.h / 2));
                const int mw = pvf.mb.w, mh = pvf.mb.h;

                // :306 (bounds are the previous frame's imageCols/imageRows == this stream's frame size)
                if ((mv.d_indx == -1 || !lbFound[mv.d_indx]) && in_bounds(mx, my, mw, mh, cols, rows)) {
                    if (mv.d_indx >= 0) lbFound[mv.d_indx] = true;  // :308-309

                    uint32_t desc[8];  // :311-314
                    orc_express_descriptor(grey, cols, mx, my, mw, mh, thr, desc);
                    const int dist = orc_express_distance(pvf.desc, desc);

                    if (dist <= 40) {  // :316-331
                        movfe_track vf;
                        vf.pt_x = px;
                        vf.pt_y = py;
                        vf.mb = {(int16_t)mx, (int16_t)my, (int16_t)mw, (int16_t)mh};
                        vf.track_id = pvf.track_id;
                        vf.age = pvf.age + 1;
                        vf.q_indx = i;
                        vf.flags = 0;
                        std::memcpy(vf.desc, desc, 32);
                        out.push(vf);
                    }
                }
            }
        }

        // Coverage features carried by LK (:337-377)
        for (size_t i = 0; i < covFeat.size(); i++) {
            if (!lk_status || lk_status[i] == 0) continue;
            const float x = lk_pts[2 * i], y = lk_pts[2 * i + 1];
            if (x < 0 || y < 0 || x >= cols || y >= rows) continue;  // :356
            carried(prev[covFeat[i]], (int)i, x, y, true);
        }

        // New features from unclaimed MV blocks (:379-416)
        for (int i = 0; i < n_kps; i++) {
            if (lbFound[i]) continue;

            const movfe_rect mb = kps[i];
            // cv::Point2f pt = (mb.br() + mb.tl()) * 0.5 — Point_<int> * double rounds through saturate_cast<int>
            const float ptx = (float)std::lrint((mb.x + mb.w + mb.x) * 0.5);
            const float pty = (float)std::lrint((mb.y + mb.h + mb.y) * 0.5);

            if (in_bounds(mb.x, mb.y, mb.w, mb.h, cols, rows)) {  // :388
                if (orc_express_test(grey, cols, mb.x, mb.y, mb.w, mb.h, thr)) {
                    movfe_track vf;
                    orc_express_descriptor(grey, cols, mb.x, mb.y, mb.w, mb.h, thr, vf.desc);
                    mCurrentId++;
                    vf.pt_x = ptx;
                    vf.pt_y = pty;
                    vf.mb = mb;
                    vf.track_id = mCurrentId;
                    vf.age = 0;
                    vf.q_indx = -1;
                    vf.flags = 0;
                    out.push(vf);
                    mov_cnt++;
                }
            }
        }

        // Coverage back-fill (:418-451): extract_moves (:39-61) on the 16-px lattice
        if (coverage_area < params->coverage_threshold || mov_cnt < 60) {
            for (int y = 8; y < rows - 8; y += 16) {
                for (int x = 8; x < cols - 8; x += 16) {
                    const int mx = x - 8, my = y - 8;
                    if (!in_bounds(mx, my, 16, 16, cols, rows)) continue;         // :46
                    if (!orc_express_test(grey, cols, mx, my, 16, 16, thr)) continue;  // :50
                    // :427-432: same rectangle, bounds against the previous frame's size, then the slot test
                    if (slot(y, x, 0) >= 0) continue;
                    movfe_track vf;
                    orc_express_descriptor(grey, cols, mx, my, 16, 16, thr, vf.desc);
                    mCurrentId++;
                    vf.pt_x = (float)x;
                    vf.pt_y = (float)y;
                    vf.mb = {(int16_t)mx, (int16_t)my, 16, 16};
                    vf.track_id = mCurrentId;
                    vf.age = 0;
                    vf.q_indx = -1;
                    vf.flags = MOVFE_TRACK_COVERAGE;
                    out.push(vf);
                }
            }
        }
    }

    *current_id = mCurrentId;
    if (n_births) *n_births = mov_cnt;
    return out.n < out.cap ? out.n : out.cap;
}
