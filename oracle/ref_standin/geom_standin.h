// geom_standin.h — TEST INFRASTRUCTURE. Minimal stand-ins for the Eigen / Sophus types that appear in the reference
// headers compiled by oracle/Makefile (_ref): include/Frame.h holds poses as Sophus::SE3<float> and Eigen 3-vectors,
// include/MOVMatcher.h (Fuse, out of scope but compiled with the header) multiplies a pose by a point. Only storage and
// the handful of operations those headers spell out are provided; nothing on the parity path goes through here.
#pragma once
#include <cmath>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

namespace Eigen {
template <typename T, int R, int C>
struct Matrix {
    T m[R * C];
    Matrix() {
        for (int i = 0; i < R * C; i++) m[i] = T(0);
    }
    Matrix(T a, T b) {
        static_assert(R * C == 2, "two-element constructor");
        m[0] = a, m[1] = b;
    }
    Matrix(T a, T b, T c) {
        static_assert(R * C == 3, "three-element constructor");
        m[0] = a, m[1] = b, m[2] = c;
    }
    T &operator()(int i) { return m[i]; }
    const T &operator()(int i) const { return m[i]; }
    T &operator()(int r, int c) { return m[r * C + c]; }
    const T &operator()(int r, int c) const { return m[r * C + c]; }
    Matrix operator-(const Matrix &o) const {
        Matrix x;
        for (int i = 0; i < R * C; i++) x.m[i] = m[i] - o.m[i];
        return x;
    }
    Matrix operator+(const Matrix &o) const {
        Matrix x;
        for (int i = 0; i < R * C; i++) x.m[i] = m[i] + o.m[i];
        return x;
    }
    T dot(const Matrix &o) const {
        T s = T(0);
        for (int i = 0; i < R * C; i++) s += m[i] * o.m[i];
        return s;
    }
    T norm() const { return std::sqrt(dot(*this)); }
};
template <typename T, int R, int K, int C>
static inline Matrix<T, R, C> operator*(const Matrix<T, R, K> &a, const Matrix<T, K, C> &b) {
    Matrix<T, R, C> x;
    for (int r = 0; r < R; r++)
        for (int c = 0; c < C; c++) {
            T s = T(0);
            for (int k = 0; k < K; k++) s += a(r, k) * b(k, c);
            x(r, c) = s;
        }
    return x;
}
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 3> Matrix3d;
}  // namespace Eigen

namespace Sophus {
template <typename T>
class SE3 {
public:
    SE3() {
        for (int i = 0; i < 3; i++) R_(i, i) = T(1);
    }
    SE3(const Eigen::Matrix<T, 3, 3> &R, const Eigen::Matrix<T, 3, 1> &t) : R_(R), t_(t) {}
    const Eigen::Matrix<T, 3, 3> &rotationMatrix() const { return R_; }
    const Eigen::Matrix<T, 3, 1> &translation() const { return t_; }
    Eigen::Matrix<T, 3, 1> operator*(const Eigen::Matrix<T, 3, 1> &p) const { return R_ * p + t_; }

private:
    Eigen::Matrix<T, 3, 3> R_;
    Eigen::Matrix<T, 3, 1> t_;
};
typedef SE3<float> SE3f;
typedef SE3<double> SE3d;
}  // namespace Sophus
