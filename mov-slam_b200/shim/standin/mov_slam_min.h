// mov_slam_min.h — STAND-IN declarations, used only when the shims are built outside the MoV-SLAM tree (this repo's
// tests). They declare just the members the shims touch, with the reference's names and types, so that the same shim
// sources compile against the real headers (-DMOVFE_IN_TREE: include/Frame.h, MapPoint.h, KeyFrame.h,
// CameraModels/GeometricCamera.h, OpenCV, Eigen, Sophus) and against these. Nothing here is reference code: no method has
// behaviour beyond storing a value.
#pragma once
#include <bitset>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <vector>

// OpenCV's type codes for the three element types the front-end uses
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32SC4 28

namespace cv {
struct Point2f {
    float x = 0, y = 0;
    Point2f() {}
    Point2f(float x_, float y_) : x(x_), y(y_) {}
};
struct Rect {
    int x = 0, y = 0, width = 0, height = 0;
    Rect() {}
    Rect(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {}
};
struct KeyPoint {
    Point2f pt;
    float size = 0;
    KeyPoint() {}
    KeyPoint(Point2f p, float s) : pt(p), size(s) {}
};
// dense 2-D array: only data / rows / cols / step / empty() / create-like constructor are used
class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;  // bytes per row
    unsigned char *data = nullptr;
    Mat() {}
    Mat(int r, int c, int type) : rows(r), cols(c) {
        const size_t elem_bytes = type == CV_32SC4 ? 16 : type == CV_8UC3 ? 3 : 1;
        step = (size_t)c * elem_bytes;
        buf.reset(new std::vector<unsigned char>((size_t)r * step));
        data = buf->data();
    }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
private:
    std::shared_ptr<std::vector<unsigned char>> buf;
};
}  // namespace cv

namespace Eigen {
struct Vector3f {
    float v[3] = {0, 0, 0};
    float &operator()(int i) { return v[i]; }
    float operator()(int i) const { return v[i]; }
    float x() const { return v[0]; }
    float y() const { return v[1]; }
    float z() const { return v[2]; }
};
struct Matrix3f {
    float m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    float &operator()(int r, int c) { return m[r * 3 + c]; }
    float operator()(int r, int c) const { return m[r * 3 + c]; }
};
}  // namespace Eigen

namespace Sophus {
template <typename T>
class SE3 {
public:
    SE3() {}
    SE3(const Eigen::Matrix3f &R, const Eigen::Vector3f &t) : R_(R), t_(t) {}
    const Eigen::Matrix3f &rotationMatrix() const { return R_; }
    const Eigen::Vector3f &translation() const { return t_; }
private:
    Eigen::Matrix3f R_;
    Eigen::Vector3f t_;
};
using SE3f = SE3<float>;
}  // namespace Sophus

namespace MOV_SLAM {
using std::bitset;
using std::map;
using std::shared_ptr;
using std::vector;

enum class FrameType { I_FRAME, P_FRAME };  // include/Frame.h:49-53

struct MotionVector {  // include/Frame.h:55-77
    int indx = -1;
    bool occupied = false;
    cv::Point2f pt;
    int dIndx = -1;
    cv::Rect mb;
};

struct VideoFeature {  // include/Frame.h:79-107
    int trackId = -1;
    int qIndx = -1;
    int dIndx = -1;
    cv::Point2f pt;
    cv::Rect mb;
    int age = 0;
    bitset<256> desc;
    bool coverage = false;
};

struct VideoImage {  // include/Frame.h:109-156
    cv::Mat imGray, imRGB, mvi;
    vector<cv::Rect> kps;
    vector<MotionVector> mvs;
    FrameType ft = FrameType::I_FRAME;
    double coverageArea = 0.0;
    int frame = 0;
    VideoImage(int width, int height) : mvi(height, width, CV_32SC4) { std::memset(mvi.data, 0xff, (size_t)width * height * 16); }
};
typedef VideoImage MotionVectorImage;

class GeometricCamera {  // include/CameraModels/GeometricCamera.h:61-101
public:
    const static unsigned int CAM_PINHOLE = 0;
    const static unsigned int CAM_FISHEYE = 1;
    GeometricCamera(const vector<float> &params, unsigned type) : mvParameters(params), mnType(type) {}
    float getParameter(const int i) { return mvParameters[i]; }
    size_t size() { return mvParameters.size(); }
    unsigned int GetType() { return mnType; }
protected:
    vector<float> mvParameters;
    unsigned int mnType;
};

class MapPoint {  // include/MapPoint.h (members read by the front-end)
public:
    Eigen::Vector3f GetWorldPos() { return mWorldPos; }
    bool isBad() { return mbBad; }
    int mTrackId = -1;
    bool mbTrackInView = false;
    float mTrackProjX = 0.f, mTrackProjY = 0.f;
    float mTrackDepth = 0.f;
    float mTrackViewCos = 0.f;
    std::bitset<256> GetDescriptor() { return mDescriptor; }   // include/MapPoint.h:112
    int Observations() { return nObs; }                        // include/MapPoint.h:90
    std::bitset<256> mDescriptor;
    int nObs = 0;
    Eigen::Vector3f mWorldPos;
    bool mbBad = false;
};

class KeyFrame {
public:
    vector<MapPoint *> GetMapPointMatches() { return mvpMapPoints; }
    vector<MapPoint *> mvpMapPoints;
    cv::Mat mImage;
};

class Frame {  // include/Frame.h:158-466 (members the shims touch)
public:
    int N = 0;
    vector<cv::KeyPoint> mvKeys, mvKeysUn;
    vector<VideoFeature> mvVF;
    map<int, int> mvVFMap;
    vector<MapPoint *> mvpMapPoints;
    vector<bitset<256>> mDescriptors;
    vector<bool> mvbOutlier;
    GeometricCamera *mpCamera = nullptr;
    bool mLost = false;
    KeyFrame *mpReferenceKF = nullptr;
    cv::Mat imgLeft;
    int imageCols = 0, imageRows = 0;
    void SetPose(const Sophus::SE3<float> &Tcw) { mTcw = Tcw; }
    Sophus::SE3<float> GetPose() const { return mTcw; }
private:
    Sophus::SE3<float> mTcw;
};
}  // namespace MOV_SLAM
