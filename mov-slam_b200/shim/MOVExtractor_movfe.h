// Same declaration as the reference's include/MOVExtractor.h:27-41 (used when built outside the tree).
#pragma once
#include "movfe_shim.h"

namespace MOV_SLAM {
class MOVExtractor {
public:
    MOVExtractor(int threshold = 20, double coverageThreshold = 0.60, double relocalizationDistance = 0.25);
    ~MOVExtractor() {}

    int operator()(const shared_ptr<MotionVectorImage> &_smv, std::vector<cv::KeyPoint> &_keypoints, std::vector<VideoFeature> &_vf,
                   std::map<int, int> &_vfmap, std::vector<std::bitset<256>> &descriptors, Frame *_prev_frame);

    int mCurrentId;
    int mThreshold;
    double mCoverageThreshold;
    double mRelocalizationDistance;
};
}  // namespace MOV_SLAM
