// MOVMatcher_movfe.h — drop-in for the tracking-thread functions of include/MOVMatcher.h:35-137: same static signatures,
// the track-id joins run through movfe_join. Eligibility of a map point is evaluated on the host exactly as the reference
// does (isBad(), mbTrackInView, bFarPoints && mTrackDepth > thFarPoints): those getters take per-object mutexes.
// SearchForTriangulation / Fuse (LocalMapping) stay on the reference's code (out of scope).
#pragma once
#include <cstring>

#include "movfe_shim.h"

namespace MOV_SLAM {
class MOVMatcher {
    static int join(Frame &F, const std::vector<int32_t> &probe_ids, const std::vector<uint8_t> &probe_ok, std::vector<int32_t> &match) {
        movfe_ctx *ctx = movfe_shim::operator_context();
        if (!ctx) return 0;
        std::vector<int32_t> tid(F.mvVF.size());
        for (size_t i = 0; i < tid.size(); i++) tid[i] = F.mvVF[i].trackId;
        const int32_t toff[2] = {0, (int32_t)tid.size()}, poff[2] = {0, (int32_t)probe_ids.size()};
        int32_t n = 0;
        if (movfe_join(ctx, 1, tid.data(), toff, probe_ids.data(), probe_ok.data(), poff, match.data(), &n) != MOVFE_OK) {
            movfe_shim::fail(ctx, "join");
            return 0;
        }
        return n;
    }

public:
    static int SearchByVideoFeature(Frame &F, const vector<MapPoint *> &vpMapPoints, const bool bFarPoints, const float thFarPoints) {
        std::vector<int32_t> ids(vpMapPoints.size());
        std::vector<uint8_t> ok(vpMapPoints.size());
        for (size_t i = 0; i < vpMapPoints.size(); i++) {
            MapPoint *pMP = vpMapPoints[i];
            ids[i] = pMP->mTrackId;
            ok[i] = !(bFarPoints && pMP->mTrackDepth > thFarPoints) && !pMP->isBad() && pMP->mbTrackInView;  // :43-49
        }
        std::vector<int32_t> match(F.mvVF.size(), -1);
        const int n = join(F, ids, ok, match);
        for (size_t t = 0; t < match.size(); t++)
            if (match[t] >= 0) F.mvpMapPoints[t] = vpMapPoints[match[t]];  // entries not hit keep their value
        return n;
    }

    static int SearchByVideoFeature(KeyFrame *pKF, Frame &F, vector<MapPoint *> &vpMapPointMatches) {
        const vector<MapPoint *> vpMapPointsKF = pKF->GetMapPointMatches();
        vpMapPointMatches = vector<MapPoint *>(F.N, static_cast<MapPoint *>(NULL));  // :73
        std::vector<int32_t> ids(vpMapPointsKF.size());
        std::vector<uint8_t> ok(vpMapPointsKF.size());
        for (size_t i = 0; i < vpMapPointsKF.size(); i++) {
            MapPoint *pMP = vpMapPointsKF[i];
            ok[i] = pMP && !pMP->isBad();  // :82-86
            ids[i] = pMP ? pMP->mTrackId : 0;
        }
        std::vector<int32_t> match(F.mvVF.size(), -1);
        const int n = join(F, ids, ok, match);
        for (size_t t = 0; t < match.size() && t < vpMapPointMatches.size(); t++)
            if (match[t] >= 0) vpMapPointMatches[t] = vpMapPointsKF[match[t]];
        return n;
    }

    // Grid-bucketed search by projection over the frame's bucket grid (Frame.cc:356-388, 602-668) - an ADDITION: the reference's
    // matcher has no such function (its grid is built and never queried). ORB-SLAM3's SearchByProjection(Frame&, vector<MapPoint*>&,
    // th, bFarPoints, thFarPoints) on MoV-SLAM's types, batched per call through movfe_search_by_projection (include/movfe.h says
    // how conflicts are settled). Keypoints that already hold an observed map point are passed over.
    static int SearchByProjection(Frame &F, const vector<MapPoint *> &vpMapPoints, const float th = 3, const bool bFarPoints = false,
                                  const float thFarPoints = 50.0f, const int TH_HIGH = 100, const float mfNNratio = 0.8f) {
        movfe_ctx *ctx = movfe_shim::frame_operator_context(F.imageCols, F.imageRows);
        if (!ctx) return 0;
        const size_t n = F.mvVF.size(), m = vpMapPoints.size();
        std::vector<movfe_track> feat(n);
        std::vector<uint8_t> taken(n, 0);
        for (size_t i = 0; i < n; i++) {
            feat[i] = movfe_shim::pack(F.mvVF[i]);
            feat[i].pt_x = F.mvKeysUn[i].pt.x;
            feat[i].pt_y = F.mvKeysUn[i].pt.y;
            if (i < F.mvpMapPoints.size() && F.mvpMapPoints[i]) taken[i] = F.mvpMapPoints[i]->Observations() > 0;
        }
        std::vector<movfe_map_point> pts(m);
        std::vector<movfe_projection> proj(m);
        std::vector<uint32_t> desc(m * 8, 0u);
        for (size_t k = 0; k < m; k++) {
            MapPoint *pMP = vpMapPoints[k];
            memset(&pts[k], 0, sizeof pts[k]);
            memset(&proj[k], 0, sizeof proj[k]);
            if (!pMP) {
                pts[k].flags = MOVFE_MP_NULL;
                continue;
            }
            pts[k].flags = pMP->isBad() ? MOVFE_MP_BAD : 0u;
            pts[k].track_id = pMP->mTrackId;
            proj[k].in_view = pMP->mbTrackInView;
            proj[k].u = pMP->mTrackProjX;
            proj[k].v = pMP->mTrackProjY;
            proj[k].depth = pMP->mTrackDepth;
            proj[k].view_cos = pMP->mTrackViewCos;
            const std::bitset<256> d = pMP->GetDescriptor();
            for (int b = 0; b < 256; b++)
                if (d[b]) desc[k * 8 + (b >> 5)] |= 1u << (b & 31);
        }
        const int32_t foff[2] = {0, (int32_t)n}, poff[2] = {0, (int32_t)m};
        const movfe_projection_search_params prm = {th, bFarPoints ? 1 : 0, thFarPoints, TH_HIGH, mfNNratio};
        std::vector<int32_t> fm(n ? n : 1), pm(m ? m : 1), pd(m ? m : 1);
        int32_t nm = 0;
        if (movfe_search_by_projection(ctx, 1, feat.data(), taken.data(), foff, pts.data(), proj.data(), desc.data(), poff, &prm, fm.data(),
                                       pm.data(), pd.data(), &nm) != MOVFE_OK) {
            movfe_shim::fail(ctx, "search_by_projection");
            return 0;
        }
        for (size_t i = 0; i < n && i < F.mvpMapPoints.size(); i++)
            if (fm[i] >= 0) F.mvpMapPoints[i] = vpMapPoints[fm[i]];
        return nm;
    }

    static int SearchForInitialization(Frame &F1, Frame &F2, vector<cv::Point2f> &vbPrevMatched, vector<int> &vnMatches12, int windowSize) {
        (void)windowSize;  // unused by the reference as well (:105-137)
        vnMatches12 = vector<int>(F1.mvKeysUn.size(), -1);
        std::vector<int32_t> ids(F2.mvVF.size());
        std::vector<uint8_t> ok(F2.mvVF.size(), 1);
        for (size_t i = 0; i < ids.size(); i++) ids[i] = F2.mvVF[i].trackId;
        std::vector<int32_t> match(F1.mvVF.size(), -1);
        const int n = join(F1, ids, ok, match);
        for (size_t i1 = 0; i1 < vnMatches12.size() && i1 < match.size(); i1++) vnMatches12[i1] = match[i1];
        for (size_t i1 = 0; i1 < vnMatches12.size(); i1++)  // :131-134
            if (vnMatches12[i1] >= 0) vbPrevMatched[i1] = F2.mvKeysUn[vnMatches12[i1]].pt;
        return n;
    }
};
}  // namespace MOV_SLAM
