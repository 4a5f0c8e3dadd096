// oracle/match.cc — TEST INFRASTRUCTURE (see oracle.h).
// CPU restatement of Frame::isInFrustum (mono branch, src/Frame.cc:456-519) with Pinhole::project
// (src/CameraModels/Pinhole.cpp:45-52), the trackId joins of include/MOVMatcher.h:35-137 and the bucket grid
// (src/Frame.cc:356-388, 602-680).
//
// Float evaluation order. Eigen is not in this container; the fixed-size 3-vector reductions are restated in
// the order Eigen 3.4's unrolled reduction produces, c0 + (c1 + c2) (redux_novec_unroller splits at Length/2),
// without FMA contraction. That order is recalled, not re-read: "parity unpinned" for the last bit of the
// projections. The camera centre mOw comes from Sophus' quaternion inverse in the
// reference (Frame.cc:429-431); here it is -R^T t evaluated in double and rounded to float.
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <map>
#include <vector>

namespace {

inline float dot3(const float a[3], const float b[3]) { return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]); }

// KannalaBrandt8 is absent from the reference tree; ORB-SLAM3 lineage formula
// (App. A.6), evaluated in double and rounded to float so that host and device libm agree.
inline void project_fisheye_f(const movfe_camera *cam, const float Pc[3], float uv[2]) {
    const double x = Pc[0], y = Pc[1], z = Pc[2];
    const double r = std::sqrt(x * x + y * y);
    const double theta = std::atan2(r, z);
    const double t2 = theta * theta;
    const double thetad = theta * (1.0 + t2 * ((double)cam->k[0] + t2 * ((double)cam->k[1] + t2 * ((double)cam->k[2] + t2 * (double)cam->k[3]))));
    const double s = r > 1e-12 ? thetad / r : 1.0;
    uv[0] = (float)((double)cam->fx * s * x + (double)cam->cx);
    uv[1] = (float)((double)cam->fy * s * y + (double)cam->cy);
}

}  // namespace

extern "C" void orc_frustum(const movfe_pose *Tcw, const movfe_camera *cam, int width, int height,
                            float viewingCosLimit, const movfe_map_point *pts, int n, movfe_projection *out) {
    float mRcw[9], mtcw[3], mOw[3];
    for (int i = 0; i < 9; i++) mRcw[i] = (float)Tcw->R[i];
    for (int i = 0; i < 3; i++) mtcw[i] = (float)Tcw->t[i];
    for (int i = 0; i < 3; i++)
        mOw[i] = (float)(-(Tcw->R[0 * 3 + i] * Tcw->t[0] + Tcw->R[1 * 3 + i] * Tcw->t[1] + Tcw->R[2 * 3 + i] * Tcw->t[2]));
    const float mnMinX = 0.0f, mnMaxX = (float)width, mnMinY = 0.0f, mnMaxY = (float)height;  // Frame.cc:739-745

    for (int k = 0; k < n; k++) {
        const movfe_map_point &mp = pts[k];
        movfe_projection &o = out[k];
        o.in_view = 0;  // :460-462
        o.u = -1;
        o.v = -1;
        o.depth = 0;
        o.view_cos = 0;
        // Tracking::SearchLocalPoints (:1136-1139) never calls isInFrustum for these
        if (mp.flags & (MOVFE_MP_BAD | MOVFE_MP_SKIP | MOVFE_MP_NULL)) continue;

        const float *P = mp.pos;  // :465
        float Pc[3];              // :468  Pc = mRcw * P + mtcw
        for (int i = 0; i < 3; i++) Pc[i] = dot3(&mRcw[3 * i], P) + mtcw[i];
        const float Pc_dist = std::sqrt(dot3(Pc, Pc));  // :469

        const float PcZ = Pc[2];  // :472-475
        if (PcZ < 0.0f) continue;

        float uv[2];  // :477
        if (cam->model == MOVFE_CAM_FISHEYE) {
            project_fisheye_f(cam, Pc, uv);
        } else {  // Pinhole.cpp:48-49: fx * x / z + cx
            uv[0] = cam->fx * Pc[0] / Pc[2] + cam->cx;
            uv[1] = cam->fy * Pc[1] / Pc[2] + cam->cy;
        }

        if (uv[0] < mnMinX || uv[0] > mnMaxX) continue;  // :479-482
        if (uv[1] < mnMinY || uv[1] > mnMaxY) continue;

        o.u = uv[0];  // :484-485
        o.v = uv[1];

        const float maxDistance = 1.2f * mp.max_dist;  // :488-489, MapPoint.cc:443-453
        const float minDistance = 0.8f * mp.min_dist;
        const float PO[3] = {P[0] - mOw[0], P[1] - mOw[1], P[2] - mOw[2]};  // :490
        const float dist = std::sqrt(dot3(PO, PO));

        if (dist < minDistance || dist > maxDistance) continue;  // :493

        const float viewCos = dot3(PO, mp.normal) / dist;  // :497-499

        if (viewCos < viewingCosLimit) continue;  // :501

        o.in_view = 1;  // :508-516
        o.depth = Pc_dist;
        o.view_cos = viewCos;
    }
}

// F.mvVFMap: std::map<int,int>::insert never overwrites, so the first index of a track id wins
// (MOVExtractor.cc:330).
static std::map<int, int> build_vfmap(const movfe_track *tracks, int n) {
    std::map<int, int> m;
    for (int i = 0; i < n; i++) m.insert({tracks[i].track_id, i});
    return m;
}

// MOVMatcher.h:35-68
extern "C" int orc_search_by_video_feature(const movfe_track *tracks, int n_tracks, const movfe_map_point *pts,
                                           const movfe_projection *proj, int n_pts, int bFarPoints,
                                           float thFarPoints, int32_t *match) {
    std::map<int, int> vfmap = build_vfmap(tracks, n_tracks);
    int nmatches = 0;
    for (int iMP = 0; iMP < n_pts; iMP++) {
        if (bFarPoints && proj[iMP].depth > thFarPoints) continue;  // :43-44
        if (pts[iMP].flags & MOVFE_MP_BAD) continue;                // :46-47
        if (proj[iMP].in_view) {                                    // :49
            auto it = vfmap.find(pts[iMP].track_id);
            if (it != vfmap.end()) {  // :51-55
                match[it->second] = iMP;
                nmatches++;
            }
        }
    }
    return nmatches;
}

// MOVMatcher.h:70-103
extern "C" int orc_search_by_keyframe(const movfe_track *tracks, int n_tracks, const movfe_map_point *kf_pts,
                                      int n_pts, int32_t *match) {
    std::map<int, int> vfmap = build_vfmap(tracks, n_tracks);
    for (int i = 0; i < n_tracks; i++) match[i] = -1;  // :73
    int nmatches = 0;
    for (int i = 0; i < n_pts; i++) {
        if (kf_pts[i].flags & MOVFE_MP_NULL) continue;  // :81
        if (kf_pts[i].flags & MOVFE_MP_BAD) continue;   // :83-84
        auto it = vfmap.find(kf_pts[i].track_id);
        if (it != vfmap.end()) {  // :86-90
            match[it->second] = i;
            nmatches++;
        }
    }
    return nmatches;
}

// MOVMatcher.h:105-137
extern "C" int orc_search_for_initialization(const movfe_track *f1, int n1, const movfe_track *f2, int n2,
                                             float *vbPrevMatched, int32_t *vnMatches12) {
    std::map<int, int> vfmap1 = build_vfmap(f1, n1);
    int nmatches = 0;
    for (int i = 0; i < n1; i++) vnMatches12[i] = -1;  // :108
    for (int i1 = 0; i1 < n2; i1++) {                  // :111-119 (iterates F2 despite the name)
        auto it = vfmap1.find(f2[i1].track_id);
        if (it != vfmap1.end()) {
            vnMatches12[it->second] = i1;
            nmatches++;
        }
    }
    for (int i1 = 0; i1 < n1; i1++)  // :132-134
        if (vnMatches12[i1] >= 0) {
            vbPrevMatched[2 * i1] = f2[vnMatches12[i1]].pt_x;
            vbPrevMatched[2 * i1 + 1] = f2[vnMatches12[i1]].pt_y;
        }
    return nmatches;
}

// ---- bucket grid ---------------------------------------------------------------------------------------------
#define FRAME_GRID_ROWS 48  // Frame.h:40-41
#define FRAME_GRID_COLS 64

// Frame.cc:670-680
static bool PosInGrid(float x, float y, float mnMinX, float mnMinY, float wInv, float hInv, int &posX, int &posY) {
    posX = (int)std::round((x - mnMinX) * wInv);
    posY = (int)std::round((y - mnMinY) * hInv);
    if (posX < 0 || posX >= FRAME_GRID_COLS || posY < 0 || posY >= FRAME_GRID_ROWS) return false;
    return true;
}

// Frame.cc:356-388 (mono: Nleft == -1, mvKeysUn == mvKeys when k1 == 0, Frame.cc:684-688)
extern "C" void orc_assign_features_to_grid(const movfe_track *tracks, int n, int width, int height,
                                            int32_t *cell_start, int32_t *cell_items) {
    const float mnMinX = 0.0f, mnMaxX = (float)width, mnMinY = 0.0f, mnMaxY = (float)height;
    const float wInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(mnMaxX - mnMinX);  // :147-148
    const float hInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(mnMaxY - mnMinY);
    std::vector<std::vector<int>> mGrid(FRAME_GRID_COLS * FRAME_GRID_ROWS);
    for (int i = 0; i < n; i++) {
        int gx, gy;
        if (PosInGrid(tracks[i].pt_x, tracks[i].pt_y, mnMinX, mnMinY, wInv, hInv, gx, gy))
            mGrid[gx * FRAME_GRID_ROWS + gy].push_back(i);
    }
    int k = 0;
    for (int c = 0; c < FRAME_GRID_COLS * FRAME_GRID_ROWS; c++) {
        cell_start[c] = k;
        for (int v : mGrid[c]) cell_items[k++] = v;
    }
    cell_start[FRAME_GRID_COLS * FRAME_GRID_ROWS] = k;
}

// Frame.cc:602-668 (minLevel = 0, maxLevel = -1: no level checks; all keypoints are octave 0)
extern "C" int orc_get_features_in_area(const movfe_track *tracks, int n, int width, int height,
                                        const int32_t *cell_start, const int32_t *cell_items, float x, float y,
                                        float r, int32_t *out) {
    (void)n;
    const float mnMinX = 0.0f, mnMaxX = (float)width, mnMinY = 0.0f, mnMaxY = (float)height;
    const float wInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(mnMaxX - mnMinX);
    const float hInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(mnMaxY - mnMinY);
    int cnt = 0;
    float factorX = r, factorY = r;

    const int nMinCellX = std::max(0, (int)std::floor((x - mnMinX - factorX) * wInv));
    if (nMinCellX >= FRAME_GRID_COLS) return 0;
    const int nMaxCellX = std::min((int)FRAME_GRID_COLS - 1, (int)std::ceil((x - mnMinX + factorX) * wInv));
    if (nMaxCellX < 0) return 0;
    const int nMinCellY = std::max(0, (int)std::floor((y - mnMinY - factorY) * hInv));
    if (nMinCellY >= FRAME_GRID_ROWS) return 0;
    const int nMaxCellY = std::min((int)FRAME_GRID_ROWS - 1, (int)std::ceil((y - mnMinY + factorY) * hInv));
    if (nMaxCellY < 0) return 0;

    for (int ix = nMinCellX; ix <= nMaxCellX; ix++) {
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            const int c = ix * FRAME_GRID_ROWS + iy;
            for (int j = cell_start[c]; j < cell_start[c + 1]; j++) {
                const movfe_track &kp = tracks[cell_items[j]];
                const float distx = kp.pt_x - x;
                const float disty = kp.pt_y - y;
                if (std::fabs(distx) < factorX && std::fabs(disty) < factorY) out[cnt++] = cell_items[j];
            }
        }
    }
    return cnt;
}

// Grid-bucketed search by projection. NOT in the reference: its matcher joins by track id (include/MOVMatcher.h:35-68) and
// Frame::GetFeaturesInArea (src/Frame.cc:602-668) has no caller. Restated from the ORB-SLAM3 lineage
// (ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints); published algorithm, recalled) on
// MoV-SLAM's types: EXPRESS distance (include/EXPRESS.h:112-115), every keypoint on octave 0. "Parity unpinned": CUDA == this.
// One deliberate difference, stated in include/movfe.h: map points are searched independently and a keypoint chosen by several
// goes to the smallest (distance, point index); the original walks the points in order and lets a later one skip a keypoint an
// earlier one took.
extern "C" int orc_search_by_projection(const movfe_track *feat, const uint8_t *taken, int n_feat, int width, int height,
                                        const movfe_map_point *pts, const movfe_projection *proj, const uint32_t *pt_desc,
                                        int n_pts, const movfe_projection_search_params *prm, int32_t *feat_match,
                                        int32_t *pt_match, int32_t *pt_dist) {
    std::vector<int32_t> cell_start(FRAME_GRID_COLS * FRAME_GRID_ROWS + 1), cell_items((size_t)std::max(n_feat, 1)), cand((size_t)std::max(n_feat, 1));
    orc_assign_features_to_grid(feat, n_feat, width, height, cell_start.data(), cell_items.data());
    std::vector<int> prop((size_t)std::max(n_pts, 1), -1);
    for (int i = 0; i < n_feat; i++) feat_match[i] = -1;
    for (int k = 0; k < n_pts; k++) {
        pt_match[k] = -1;
        pt_dist[k] = -1;
        const movfe_map_point &mp = pts[k];
        if (!proj[k].in_view) continue;                                          // if(!pMP->mbTrackInView) continue;
        if (prm->far_points && proj[k].depth > prm->th_far) continue;            // if(bFarPoints && pMP->mTrackDepth>thFarPoints) continue;
        if (mp.flags & (MOVFE_MP_BAD | MOVFE_MP_SKIP | MOVFE_MP_NULL)) continue;  // if(pMP->isBad()) continue;
        float r = proj[k].view_cos > 0.998f ? 2.5f : 4.0f;                       // RadiusByViewingCos(pMP->mTrackViewCos)
        r = r * prm->th;                                                         // if(bFactor) r*=th;  (scale factor of level 0 = 1)
        const int nc = orc_get_features_in_area(feat, n_feat, width, height, cell_start.data(), cell_items.data(), proj[k].u,
                                                proj[k].v, r, cand.data());
        if (nc == 0) continue;
        const uint32_t *MPdescriptor = pt_desc + (size_t)k * 8;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int c = 0; c < nc; c++) {
            const int idx = cand[c];
            if (taken && taken[idx]) continue;  // if(F.mvpMapPoints[idx]) if(F.mvpMapPoints[idx]->Observations()>0) continue;
            int dist = 0;                        // (descriptor1 ^ descriptor2).count()
            for (int w = 0; w < 8; w++) dist += __builtin_popcount(MPdescriptor[w] ^ feat[idx].desc[w]);
            const int octave = 0;
            if (dist < bestDist) {
                bestDist2 = bestDist;
                bestDist = dist;
                bestLevel2 = bestLevel;
                bestLevel = octave;
                bestIdx = idx;
            } else if (dist < bestDist2) {
                bestLevel2 = octave;
                bestDist2 = dist;
            }
        }
        if (bestDist <= prm->th_high) {
            if (bestLevel == bestLevel2 && (float)bestDist > prm->nn_ratio * (float)bestDist2) continue;
            prop[(size_t)k] = bestIdx;
            pt_dist[k] = bestDist;
        }
    }
    // a keypoint chosen by several map points: the smallest (distance, point index) holds it
    int nmatches = 0;
    for (int k = 0; k < n_pts; k++) {
        const int f = prop[(size_t)k];
        if (f < 0) continue;
        const int cur = feat_match[f];
        if (cur < 0 || pt_dist[k] < pt_dist[cur]) feat_match[f] = k;  // ascending k: ties keep the earlier point
    }
    for (int k = 0; k < n_pts; k++) {
        const int f = prop[(size_t)k];
        if (f >= 0 && feat_match[f] == k) {
            pt_match[k] = f;
            nmatches++;
        }
    }
    return nmatches;
}

// Tracking::UpdateLocalPoints (src/Tracking.cc:1171-1198) on index lists: `idx` is the concatenation of the local keyframes'
// GetMapPointMatches() in the order the reference walks them (mvpLocalKeyFrames reversed, each list in order; a NULL entry is -1),
// `store` the stream's map points by index. mnTrackReferenceForFrame is modelled by a visited set. Returns the number of points
// written to `out` (at most `capacity`); *n_from_first = how many of them came from the first n_first list entries.
extern "C" int orc_update_local_points(const movfe_map_point *store, int n_store, const int32_t *idx, int n_idx, int n_first,
                                       movfe_map_point *out, int capacity, int32_t *n_from_first) {
    std::vector<char> seen((size_t)std::max(n_store, 1), 0);
    int n = 0, first = 0;
    for (int j = 0; j < n_idx; j++) {
        const int i = idx[j];
        if (i < 0 || i >= n_store) continue;               // if (!pMP) continue;                       :1187-1188
        if (seen[(size_t)i]) continue;                     // mnTrackReferenceForFrame == current frame   :1189-1190
        if (store[i].flags & MOVFE_MP_BAD) continue;       // if (!pMP->isBad())                         :1191
        seen[(size_t)i] = 1;                               // mnTrackReferenceForFrame = mCurrentFrame.mnId :1195
        if (n < capacity) {
            out[n] = store[i];                             // mvpLocalMapPoints.push_back(pMP)            :1194
            if (j < n_first) first = n + 1;
        }
        n++;
    }
    if (n_from_first) *n_from_first = first;
    return n < capacity ? n : capacity;
}
