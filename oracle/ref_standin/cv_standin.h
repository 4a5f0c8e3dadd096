// cv_standin.h — TEST INFRASTRUCTURE. A minimal stand-in for the part of OpenCV's C++ API that the reference's
// front-end sources touch (src/VideoDecoder.cc, src/MOVExtractor.cc, include/EXPRESS.h, include/MOVMatcher.h,
// include/Frame.h, include/VideoBase.h), so that those sources compile UNMODIFIED in this image, where OpenCV C++ is
// absent (oracle/Makefile, target _ref). Nothing here is reference code. Semantics restated from OpenCV's published
// API (4.x): cv::Mat is a reference-counted view (data / step / rows / cols; operator()(Rect) shares the buffer;
// clone() copies), saturate_cast<int>(double) rounds to nearest-even (cvRound), Rect_<int>(float...) truncates through
// the implicit float->int conversion of the call.
//
// cv::calcOpticalFlowPyrLK is NOT restated (third-party arithmetic): it returns whatever the test
// driver injected through cvstub::lk_queue() — exactly how the C-ABI receives host LK results.
#pragma once

#include <algorithm>
#include <bitset>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_32S 4
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32SC4 CV_MAKETYPE(CV_32S, 4)

namespace cv {

template <typename T>
static inline T saturate_cast(double v);
template <>
inline int saturate_cast<int>(double v) {  // cvRound: round half to even
    return (int)std::lrint(v);
}
template <>
inline float saturate_cast<float>(double v) {
    return (float)v;
}
template <>
inline double saturate_cast<double>(double v) {
    return v;
}

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename T2>
    operator Point_<T2>() const {
        return Point_<T2>(saturate_cast<T2>((double)x), saturate_cast<T2>((double)y));
    }
};
template <typename T>
static inline Point_<T> operator+(const Point_<T> &a, const Point_<T> &b) {
    return Point_<T>((T)(a.x + b.x), (T)(a.y + b.y));
}
template <typename T>
static inline Point_<T> operator-(const Point_<T> &a, const Point_<T> &b) {
    return Point_<T>((T)(a.x - b.x), (T)(a.y - b.y));
}
template <typename T>
static inline Point_<T> operator*(const Point_<T> &a, double b) {
    return Point_<T>(saturate_cast<T>(a.x * b), saturate_cast<T>(a.y * b));
}
template <typename T>
static inline double norm(const Point_<T> &p) {
    return std::sqrt((double)p.x * p.x + (double)p.y * p.y);
}
typedef Point_<int> Point;
typedef Point_<float> Point2f;

template <typename T>
struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

template <typename T>
struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w_, T h_) : x(x_), y(y_), width(w_), height(h_) {}
    T area() const { return width * height; }
    Point_<T> tl() const { return Point_<T>(x, y); }
    Point_<T> br() const { return Point_<T>(x + width, y + height); }
};
typedef Rect_<int> Rect;

struct KeyPoint {
    Point2f pt;
    float size;
    float angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f pt_, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
        : pt(pt_), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};

template <typename T, int N>
struct Vec {
    T val[N];
    T &operator[](int i) { return val[i]; }
    const T &operator[](int i) const { return val[i]; }
};
typedef Vec<int, 4> Vec4i;

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a, val[1] = b, val[2] = c, val[3] = d; }
};

struct TermCriteria {
    enum { COUNT = 1, MAX_ITER = COUNT, EPS = 2 };
    int type, maxCount;
    double epsilon;
    TermCriteria() : type(0), maxCount(0), epsilon(0) {}
    TermCriteria(int t, int c, double e) : type(t), maxCount(c), epsilon(e) {}
};

enum { OPTFLOW_USE_INITIAL_FLOW = 4, OPTFLOW_LK_GET_MIN_EIGENVALS = 8 };

// dense 2-D array view, the subset the front-end sources use
class Mat {
public:
    struct MSize {
        int p[2];
        int &operator[](int i) { return p[i]; }
        const int &operator[](int i) const { return p[i]; }
        Size operator()() const { return Size(p[1], p[0]); }
    };
    struct MStep {
        size_t p[2];
        size_t &operator[](int i) { return p[i]; }
        const size_t &operator[](int i) const { return p[i]; }
        operator size_t() const { return p[0]; }
    };

    int flags_type = 0;
    int rows = 0, cols = 0;
    uchar *data = nullptr;
    MSize size{{0, 0}};
    MStep step{{0, 0}};

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, const Scalar &s) {
        create(r, c, type);
        setTo(s);
    }
    // user-data constructor (no ownership)
    Mat(int r, int c, int type, void *ptr, size_t step_bytes = 0) {
        flags_type = type;
        rows = r, cols = c;
        size[0] = r, size[1] = c;
        step[1] = elemSize();
        step[0] = step_bytes ? step_bytes : (size_t)c * elemSize();
        data = (uchar *)ptr;
    }

    void create(int r, int c, int type) {
        flags_type = type;
        rows = r, cols = c;
        size[0] = r, size[1] = c;
        step[1] = elemSize();
        step[0] = (size_t)c * elemSize();
        buf_ = std::make_shared<std::vector<uchar>>((size_t)r * step[0]);
        data = buf_->data();
    }
    int type() const { return flags_type; }
    int channels() const { return (flags_type >> 3) + 1; }
    size_t elemSize1() const {
        const int d = flags_type & 7;
        return d == CV_8U ? 1 : 4;
    }
    size_t elemSize() const { return elemSize1() * channels(); }
    size_t step1(int i = 0) const { return step[i] / elemSize1(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    void release() {
        buf_.reset();
        data = nullptr;
        rows = cols = 0;
        size[0] = size[1] = 0;
    }
    void updateContinuityFlag() {}

    uchar *ptr(int r = 0) { return data + step[0] * r; }
    const uchar *ptr(int r = 0) const { return data + step[0] * r; }
    template <typename T>
    T *ptr(int r = 0) {
        return (T *)(data + step[0] * r);
    }
    template <typename T>
    const T *ptr(int r = 0) const {
        return (const T *)(data + step[0] * r);
    }
    template <typename T>
    T &at(int r, int c) {
        return ((T *)(data + step[0] * r))[c];
    }
    template <typename T>
    const T &at(int r, int c) const {
        return ((const T *)(data + step[0] * r))[c];
    }

    Mat operator()(const Rect &roi) const {
        Mat m = *this;  // shares the buffer
        m.data = data + step[0] * roi.y + elemSize() * roi.x;
        m.rows = roi.height, m.cols = roi.width;
        m.size[0] = roi.height, m.size[1] = roi.width;
        return m;
    }
    Mat clone() const {
        Mat m;
        if (!data) return m;
        m.create(rows, cols, flags_type);
        const size_t row_bytes = (size_t)cols * elemSize();
        for (int r = 0; r < rows; r++) std::memcpy(m.data + m.step[0] * r, data + step[0] * r, row_bytes);
        return m;
    }
    Mat &setTo(const Scalar &s) {
        const int cn = channels();
        for (int r = 0; r < rows; r++) {
            if ((flags_type & 7) == CV_8U) {
                uchar *p = ptr(r);
                for (int c = 0; c < cols; c++)
                    for (int k = 0; k < cn; k++) p[c * cn + k] = (uchar)saturate_cast<int>(s.val[k]);
            } else {
                int *p = ptr<int>(r);
                for (int c = 0; c < cols; c++)
                    for (int k = 0; k < cn; k++) p[c * cn + k] = saturate_cast<int>(s.val[k]);
            }
        }
        return *this;
    }

private:
    std::shared_ptr<std::vector<uchar>> buf_;
};

class BFMatcher {};

static inline bool imwrite(const std::string &, const Mat &) { return false; }

}  // namespace cv

// ---- injected LK results -----------------------------------------------------------------------------------------
namespace cvstub {
struct LkCall {
    std::vector<uchar> status;
    std::vector<cv::Point2f> pts;
};
// results handed out by successive cv::calcOpticalFlowPyrLK calls (front first); n_calls counts the calls made;
// last_prev_pts records the points of the most recent call (so the driver can check what the reference asked for)
std::vector<LkCall> &lk_queue();
int &lk_calls();
std::vector<cv::Point2f> &lk_last_prev_pts();
}  // namespace cvstub

namespace cv {
static inline void calcOpticalFlowPyrLK(const Mat &, const Mat &, const std::vector<Point2f> &prevPts, std::vector<Point2f> &nextPts,
                                        std::vector<uchar> &status, std::vector<float> &err, Size = Size(21, 21), int = 3,
                                        TermCriteria = TermCriteria(), int = 0, double = 1e-4) {
    cvstub::lk_calls()++;
    cvstub::lk_last_prev_pts() = prevPts;
    const size_t n = prevPts.size();
    nextPts.assign(n, Point2f());
    status.assign(n, 0);
    err.assign(n, 0.f);
    auto &q = cvstub::lk_queue();
    if (q.empty()) return;  // nothing injected: every point is lost
    const cvstub::LkCall c = q.front();
    q.erase(q.begin());
    for (size_t i = 0; i < n && i < c.status.size(); i++) {
        status[i] = c.status[i];
        nextPts[i] = c.pts[i];
    }
}
}  // namespace cv
