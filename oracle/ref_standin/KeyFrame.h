// stand-in for include/KeyFrame.h (+ the GeometricCamera interface it brings in): the members the front-end sources
// compiled by oracle/Makefile (_ref) read or call (src/MOVExtractor.cc:163-196, include/MOVMatcher.h:70-277).
#pragma once
#include <vector>
#include "Frame.h"
#include "MapPoint.h"
namespace MOV_SLAM {
class GeometricCamera {   // include/CameraModels/GeometricCamera.h: the members the compiled sources and the shims touch
public:
    virtual ~GeometricCamera() {}
    virtual Eigen::Vector2f project(const Eigen::Vector3f &) { return Eigen::Vector2f(); }
    float getParameter(const int i) { return mvParameters[i]; }
    size_t size() { return mvParameters.size(); }
    unsigned int GetType() { return mnType; }
    const static unsigned int CAM_PINHOLE = 0;
    const static unsigned int CAM_FISHEYE = 1;
    std::vector<float> mvParameters;
    unsigned int mnType = 0;
};
class KeyFrame {
public:
    std::vector<MapPoint *> mvpMapPoints;
    std::vector<VideoFeature> mvVF;
    std::vector<cv::KeyPoint> mvKeysUn;
    cv::Mat mImage;
    GeometricCamera *mpCamera = nullptr;
    Sophus::SE3f mTcw;
    Eigen::Vector3f mOw;
    std::vector<MapPoint *> GetMapPointMatches() { return mvpMapPoints; }
    MapPoint *GetMapPoint(size_t i) { return mvpMapPoints[i]; }
    void AddMapPoint(MapPoint *p, size_t i) { mvpMapPoints[i] = p; }
    Sophus::SE3f GetPose() { return mTcw; }
    Eigen::Vector3f GetCameraCenter() { return mOw; }
    bool IsInImage(float, float) const { return true; }
};
}  // namespace MOV_SLAM
