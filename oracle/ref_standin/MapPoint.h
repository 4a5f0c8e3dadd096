// stand-in for include/MapPoint.h: the members the front-end sources compiled by oracle/Makefile (_ref) read or call
// (src/MOVExtractor.cc:169-192, include/MOVMatcher.h:35-277), with the reference's names and types. Storage only.
#pragma once
#include <map>
#include "Frame.h"
#include "geom_standin.h"
namespace MOV_SLAM {
class KeyFrame;
class MapPoint {
public:
    int mTrackId = -1;
    bool mbTrackInView = false;
    float mTrackProjX = 0.f, mTrackProjY = 0.f, mTrackDepth = 0.f, mTrackViewCos = 0.f;
    bool mbBad = false;
    Eigen::Vector3f mWorldPos, mNormal;
    float mfMinDistance = 0.f, mfMaxDistance = 0.f;
    int nObs = 0;
    bool isBad() { return mbBad; }
    Eigen::Vector3f GetWorldPos() { return mWorldPos; }
    Eigen::Vector3f GetNormal() { return mNormal; }
    float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }
    float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }
    bool IsInKeyFrame(KeyFrame *) { return false; }
    int Observations() { return nObs; }
    void Replace(MapPoint *) {}
    void AddObservation(KeyFrame *, int) {}
};
}  // namespace MOV_SLAM
