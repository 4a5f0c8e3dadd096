// Compile-only check of KannalaBrandt8_movfe.h against the reference's GeometricCamera interface: the abstract class below
// restates include/CameraModels/GeometricCamera.h:61-101 (signatures only) over stand-in cv / Eigen / Sophus types, so a
// missing or mistyped override fails to compile (the class would stay abstract). Numbers: tests/test_shim.py compares the
// projection and Jacobian with the oracle's camera model.
#include <cstdio>
#include <vector>
#include "../../../oracle/ref_standin/cv_standin.h"
#include "../../../oracle/ref_standin/geom_standin.h"
namespace cv {
typedef struct Point3f_ { float x, y, z; Point3f_() : x(0), y(0), z(0) {} Point3f_(float a, float b, float c) : x(a), y(b), z(c) {} } Point3f;
}
#define CV_32F 5
namespace MOV_SLAM {
class GeometricCamera {
public:
    GeometricCamera() {}
    GeometricCamera(const std::vector<float> &_vParameters) : mvParameters(_vParameters) {}
    virtual ~GeometricCamera() {}
    virtual cv::Point2f project(const cv::Point3f &p3D) = 0;
    virtual Eigen::Vector2d project(const Eigen::Vector3d &v3D) = 0;
    virtual Eigen::Vector2f project(const Eigen::Vector3f &v3D) = 0;
    virtual Eigen::Vector2f projectMat(const cv::Point3f &p3D) = 0;
    virtual float uncertainty2(const Eigen::Matrix<double, 2, 1> &p2D) = 0;
    virtual Eigen::Vector3f unprojectEig(const cv::Point2f &p2D) = 0;
    virtual cv::Point3f unproject(const cv::Point2f &p2D) = 0;
    virtual Eigen::Matrix<double, 2, 3> projectJac(const Eigen::Vector3d &v3D) = 0;
    virtual bool ReconstructWithTwoViews(const std::vector<cv::KeyPoint> &vKeys1, const std::vector<cv::KeyPoint> &vKeys2, const std::vector<int> &vMatches12,
                                         Sophus::SE3f &T21, std::vector<cv::Point3f> &vP3D, std::vector<bool> &vbTriangulated) = 0;
    virtual cv::Mat toK() = 0;
    virtual Eigen::Matrix3f toK_() = 0;
    virtual bool epipolarConstrain(GeometricCamera *otherCamera, const cv::KeyPoint &kp1, const cv::KeyPoint &kp2, const Eigen::Matrix3f &R12,
                                   const Eigen::Vector3f &t12, const float sigmaLevel, const float unc) = 0;
    virtual bool matchAndtriangulate(const cv::KeyPoint &kp1, const cv::KeyPoint &kp2, GeometricCamera *pOther, Sophus::SE3f &Tcw1, Sophus::SE3f &Tcw2,
                                     const float sigmaLevel1, const float sigmaLevel2, Eigen::Vector3f &x3Dtriangulated) = 0;
    float getParameter(const int i) { return mvParameters[i]; }
    size_t size() { return mvParameters.size(); }
    unsigned int GetType() { return mnType; }
    const static unsigned int CAM_PINHOLE = 0;
    const static unsigned int CAM_FISHEYE = 1;
    static long unsigned int nNextId;
protected:
    std::vector<float> mvParameters;
    unsigned int mnId;
    unsigned int mnType;
};
long unsigned int GeometricCamera::nNextId = 0;
}  // namespace MOV_SLAM
#include "../KannalaBrandt8_movfe.h"

// usage: camera_iface_check fx fy cx cy k1 k2 k3 k4 x y z  -> prints u v and the six Jacobian entries
int main(int argc, char **argv) {
    if (argc < 12) return 2;
    std::vector<float> p;
    for (int i = 1; i <= 8; i++) p.push_back((float)atof(argv[i]));
    MOV_SLAM::KannalaBrandt8 cam(p);
    MOV_SLAM::GeometricCamera *g = &cam;
    Eigen::Vector3d X(atof(argv[9]), atof(argv[10]), atof(argv[11]));
    const Eigen::Vector2d uv = g->project(X);
    const Eigen::Matrix<double, 2, 3> J = g->projectJac(X);
    const cv::Point3f back = g->unproject(cv::Point2f((float)uv(0), (float)uv(1)));
    printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.9g %.9g %u\n", uv(0), uv(1), J(0, 0), J(0, 1), J(0, 2), J(1, 0), J(1, 1), J(1, 2),
           back.x, back.y, g->GetType());
    return 0;
}
