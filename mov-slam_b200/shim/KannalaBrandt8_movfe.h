// KannalaBrandt8_movfe.h — the fisheye camera model behind the reference's GeometricCamera interface
// (include/CameraModels/GeometricCamera.h:61-101). The MoV-SLAM tree carries only the type constant CAM_FISHEYE (:96) and the
// Settings enum value (include/Settings.h:48); Settings::readCamera1 exits on anything but "PinHole"/"Rectified". north_star
// names the model, so it is supplied here for Frame / Tracking to hold through a GeometricCamera*: the drop-in shims read
// mnType and mvParameters = [fx, fy, cx, cy, k1, k2, k3, k4] (movfe_shim::pack) and the CUDA path projects with the same
// formulae (pose.cu: project_d / project_jac_d). Formulae: equidistant model with a 9th-order odd polynomial in theta
// (ORB-SLAM3 lineage); parity unpinned - there is no reference implementation to compare with.
// Built inside the MoV-SLAM tree (real Eigen / OpenCV / Sophus); tests/test_shim.py compiles it against stand-in headers.
#pragma once
#include <cmath>
#include <vector>
#ifdef MOVFE_IN_TREE
#include "CameraModels/GeometricCamera.h"
#endif

namespace MOV_SLAM {

class KannalaBrandt8 : public GeometricCamera {
public:
    KannalaBrandt8() : precision(1e-6f) {
        mvParameters.resize(8);
        mnId = nNextId++;
        mnType = CAM_FISHEYE;
    }
    explicit KannalaBrandt8(const std::vector<float> &_vParameters) : GeometricCamera(_vParameters), precision(1e-6f) {
        mnId = nNextId++;
        mnType = CAM_FISHEYE;
    }

    cv::Point2f project(const cv::Point3f &p3D) {
        const Eigen::Vector2d uv = project(Eigen::Vector3d(p3D.x, p3D.y, p3D.z));
        return cv::Point2f((float)uv(0), (float)uv(1));
    }
    Eigen::Vector2d project(const Eigen::Vector3d &v) {
        const double x = v(0), y = v(1), z = v(2);
        const double r = std::sqrt(x * x + y * y);
        const double theta = std::atan2(r, z);
        const double t2 = theta * theta;
        const double thetad = theta * (1.0 + t2 * (k(0) + t2 * (k(1) + t2 * (k(2) + t2 * k(3)))));
        const double s = r > 1e-12 ? thetad / r : 1.0;   // r -> 0: theta_d / r -> 1 / z * z = 1 on the optical axis
        return Eigen::Vector2d(mvParameters[0] * s * x + mvParameters[2], mvParameters[1] * s * y + mvParameters[3]);
    }
    Eigen::Vector2f project(const Eigen::Vector3f &v) {
        const Eigen::Vector2d uv = project(Eigen::Vector3d(v(0), v(1), v(2)));
        return Eigen::Vector2f((float)uv(0), (float)uv(1));
    }
    Eigen::Vector2f projectMat(const cv::Point3f &p3D) {
        const cv::Point2f p = project(p3D);
        return Eigen::Vector2f(p.x, p.y);
    }
    float uncertainty2(const Eigen::Matrix<double, 2, 1> &) { return 1.f; }

    // inverse of theta_d(theta) by Newton's method, then the ray (sin(theta) cos(psi), sin(theta) sin(psi), cos(theta)) / cos(theta)
    cv::Point3f unproject(const cv::Point2f &p2D) {
        const double mx = (p2D.x - mvParameters[2]) / mvParameters[0], my = (p2D.y - mvParameters[3]) / mvParameters[1];
        const double thetad = std::sqrt(mx * mx + my * my);
        double theta = std::fmin(std::fmax(thetad, -M_PI / 2), M_PI / 2);
        if (thetad > 1e-8) {
            for (int it = 0; it < 10; it++) {
                const double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t4 * t4;
                const double f = theta * (1 + k(0) * t2 + k(1) * t4 + k(2) * t6 + k(3) * t8) - thetad;
                const double fd = 1 + 3 * k(0) * t2 + 5 * k(1) * t4 + 7 * k(2) * t6 + 9 * k(3) * t8;
                const double step = f / fd;
                theta -= step;
                if (std::fabs(step) < precision) break;
            }
            const double scale = std::tan(theta) / thetad;
            return cv::Point3f((float)(mx * scale), (float)(my * scale), 1.f);
        }
        return cv::Point3f((float)mx, (float)my, 1.f);
    }
    Eigen::Vector3f unprojectEig(const cv::Point2f &p2D) {
        const cv::Point3f r = unproject(p2D);
        return Eigen::Vector3f(r.x, r.y, r.z);
    }

    // d(u,v)/d(x,y,z), SURVEY.md App. A.6
    Eigen::Matrix<double, 2, 3> projectJac(const Eigen::Vector3d &v) {
        const double x = v(0), y = v(1), z = v(2);
        const double r2 = x * x + y * y, r = std::sqrt(r2), D = r2 + z * z;
        const double theta = std::atan2(r, z), t2 = theta * theta;
        const double f = theta * (1.0 + t2 * (k(0) + t2 * (k(1) + t2 * (k(2) + t2 * k(3)))));
        const double fd = 1.0 + t2 * (3 * k(0) + t2 * (5 * k(1) + t2 * (7 * k(2) + t2 * 9 * k(3))));
        Eigen::Matrix<double, 2, 3> J;
        const double fx = mvParameters[0], fy = mvParameters[1];
        if (r < 1e-12) {  // on the axis the model is locally a pinhole of focal length f / z
            J(0, 0) = fx / z, J(0, 1) = 0, J(0, 2) = 0;
            J(1, 0) = 0, J(1, 1) = fy / z, J(1, 2) = 0;
            return J;
        }
        const double r3 = r2 * r;
        J(0, 0) = fx * (fd * z * x * x / (r2 * D) + f * y * y / r3);
        J(0, 1) = fx * (fd * z * x * y / (r2 * D) - f * x * y / r3);
        J(0, 2) = -fx * fd * x / D;
        J(1, 0) = fy * (fd * z * x * y / (r2 * D) - f * x * y / r3);
        J(1, 1) = fy * (fd * z * y * y / (r2 * D) + f * x * x / r3);
        J(1, 2) = -fy * fd * y / D;
        return J;
    }

    cv::Mat toK() {
        cv::Mat K(3, 3, CV_32F, cv::Scalar(0));
        K.at<float>(0, 0) = mvParameters[0];
        K.at<float>(1, 1) = mvParameters[1];
        K.at<float>(0, 2) = mvParameters[2];
        K.at<float>(1, 2) = mvParameters[3];
        K.at<float>(2, 2) = 1.f;
        return K;
    }
    Eigen::Matrix3f toK_() {
        Eigen::Matrix3f K;
        K(0, 0) = mvParameters[0], K(0, 1) = 0, K(0, 2) = mvParameters[2];
        K(1, 0) = 0, K(1, 1) = mvParameters[1], K(1, 2) = mvParameters[3];
        K(2, 0) = 0, K(2, 1) = 0, K(2, 2) = 1;
        return K;
    }

    // Initialisation geometry and triangulation stay with the reference's mapping code (out of scope):
    // a fisheye rig would undistort to the normalised plane first (unproject above) and reuse TwoViewReconstruction.
    bool ReconstructWithTwoViews(const std::vector<cv::KeyPoint> &, const std::vector<cv::KeyPoint> &, const std::vector<int> &, Sophus::SE3f &,
                                 std::vector<cv::Point3f> &, std::vector<bool> &) { return false; }
    bool epipolarConstrain(GeometricCamera *, const cv::KeyPoint &, const cv::KeyPoint &, const Eigen::Matrix3f &, const Eigen::Vector3f &, const float,
                           const float) { return false; }
    bool matchAndtriangulate(const cv::KeyPoint &, const cv::KeyPoint &, GeometricCamera *, Sophus::SE3f &, Sophus::SE3f &, const float, const float,
                             Eigen::Vector3f &) { return false; }

private:
    double k(int i) const { return mvParameters[4 + i]; }
    const float precision;
};

}  // namespace MOV_SLAM
