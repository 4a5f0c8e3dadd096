// PoseOptimization_movfe.cc — drop-in for Optimizer::PoseOptimization (include/Optimizer.h:55, src/Optimizer.cc:397-459):
// unchanged 7-argument signature, gather / <4 rule / SetPose / mvbOutlier convention / return value as the reference; the
// solver is the Huber-robust pose-only Gauss-Newton of north_star (movfe_pose_optimize) instead of cv::solvePnPRansac.
// `confidence` and `algorithm` are accepted and unused.
#include "movfe_shim.h"
#ifdef MOVFE_IN_TREE
#include "Optimizer.h"
#endif

namespace MOV_SLAM {
#ifndef MOVFE_IN_TREE
class Optimizer {
public:
    static int PoseOptimization(Frame *pFrame, const bool isLost, const int iterationCount = 50, const double reprojectionError = 5.0,
                                const double reprojectErrorLost = 8.0, const double confidence = 0.95, const int algorithm = 38);
};
#endif

int Optimizer::PoseOptimization(Frame *pFrame, const bool isLost, const int iterationCount, const double reprojectionError,
                                const double reprojectErrorLost, const double confidence, const int algorithm) {
    std::vector<float> pts, obs;
    std::vector<int> indx;
    for (size_t i = 0; i < pFrame->mvpMapPoints.size(); i++) {  // :404-413: distorted mvKeys, GetWorldPos()
        if (pFrame->mvpMapPoints[i]) {
            const Eigen::Vector3f wp = pFrame->mvpMapPoints[i]->GetWorldPos();
            obs.push_back(pFrame->mvKeys[i].pt.x);
            obs.push_back(pFrame->mvKeys[i].pt.y);
            pts.push_back(wp.x());
            pts.push_back(wp.y());
            pts.push_back(wp.z());
            indx.push_back((int)i);
        }
    }
    if (indx.size() < 4) return 0;  // :415-418, frame untouched
    movfe_ctx *ctx = movfe_shim::operator_context();
    if (!ctx) return 0;
    const movfe_camera cam = movfe_shim::pack(pFrame->mpCamera);
    movfe_pose_params pp;
    pp.is_lost = isLost;
    pp.iteration_count = iterationCount;
    pp.reprojection_error = reprojectionError;
    pp.reprojection_error_lost = reprojectErrorLost;
    pp.confidence = confidence;
    pp.algorithm = algorithm;
    pp._pad = 0;
    movfe_pose pose;  // initial estimate = the frame's current pose (the previous frame's, Tracking.cc:807)
    const Sophus::SE3<float> T0 = pFrame->GetPose();
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) pose.R[r * 3 + c] = T0.rotationMatrix()(r, c);
        pose.t[r] = T0.translation()(r);
    }
    const int32_t off[2] = {0, (int32_t)indx.size()};
    std::vector<uint8_t> outlier(indx.size());
    int32_t n_inliers = 0;
    if (movfe_pose_optimize(ctx, 1, &cam, &pp, pts.data(), obs.data(), off, &pose, outlier.data(), &n_inliers, nullptr) != MOVFE_OK) {
        movfe_shim::fail(ctx, "pose_optimize");
        return 0;  // solver failure leaves the frame untouched (:442-445)
    }
    Eigen::Matrix3f R1;
    Eigen::Vector3f t1;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) R1(r, c) = (float)pose.R[r * 3 + c];
        t1(r) = (float)pose.t[r];
    }
    pFrame->SetPose(Sophus::SE3<float>(R1, t1));                 // :451
    pFrame->mvbOutlier = std::vector<bool>(pFrame->N, true);     // :452
    for (size_t k = 0; k < indx.size(); k++)
        if (!outlier[k]) pFrame->mvbOutlier[indx[k]] = false;    // :453-456
    return n_inliers;
}
}  // namespace MOV_SLAM
