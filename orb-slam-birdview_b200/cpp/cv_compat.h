// Minimal stand-ins for the OpenCV types the ORBextractor / ORBmatcher signatures and bodies use, ONLY for builds where the
// OpenCV C++ SDK is absent (this image).  With OpenCV installed, compile with -DORBB200_HAVE_OPENCV and the real headers are
// used instead; the shim code in ORBextractor.h / ORBmatcher_b200.cc is identical in both cases.
//
// cv::Mat here is a reference-counted header with views (row / col / rowRange / colRange share the buffer) and the handful of
// CV_32F expressions src/ORBmatcher.cc writes on 3x3 / 3x1 / 4x4 matrices, evaluated the way OpenCV 4.x evaluates them for
// these shapes (pinned on cv2.gemm / cv2.norm by tests/ref_py/frustum_py_ref.py, see oracle/match_oracle.cpp):
//   A*B        float products summed left to right in float (gemm's small-matrix path)
//   A*B + C    the same sum, then (float)((double)sum + (double)c)   (MatExpr folds it into one gemm with beta = 1)
//   A - B      float subtraction (cv::subtract)
//   A/s, s*A   x * (float)alpha with alpha = 1/s or s in double (MatExpr -> convertTo with a float scale)
//   -A         exact negation;  A.t() transposition;  A.dot(B) and cv::norm(A) accumulate in double
#pragma once
#ifdef ORBB200_HAVE_OPENCV
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#else
#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5

typedef unsigned char uchar;

namespace cv {

template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<float> Point2f;
typedef Point_<int> Point2i;
typedef Point2i Point;
template <class T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
typedef Point3_<float> Point3f;

struct KeyPoint {          // same 28-byte layout as cv::KeyPoint
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
};

class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    unsigned char* data = nullptr;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* ext, size_t stp = 0) : rows(r), cols(c), data((unsigned char*)ext), type_(type) { step = stp ? stp : (size_t)c * elemSize(); }
    void create(int r, int c, int type)
    {
        if (r == rows && c == cols && type == type_ && buf_) return;
        rows = r; cols = c; type_ = type; step = (size_t)c * elemSize();
        buf_.reset(new std::vector<unsigned char>((size_t)r * step));
        data = buf_->data();
    }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    size_t elemSize() const { return type_ == CV_32F ? 4 : 1; }
    unsigned char* ptr(int y = 0) { return data + (size_t)y * step; }
    const unsigned char* ptr(int y = 0) const { return data + (size_t)y * step; }
    template <class T> T* ptr(int y = 0) { return reinterpret_cast<T*>(data + (size_t)y * step); }
    template <class T> const T* ptr(int y = 0) const { return reinterpret_cast<const T*>(data + (size_t)y * step); }
    template <class T> T& at(int r, int c) { return reinterpret_cast<T*>(data + (size_t)r * step)[c]; }
    template <class T> const T& at(int r, int c) const { return reinterpret_cast<const T*>(data + (size_t)r * step)[c]; }
    template <class T> T& at(int i) { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }
    template <class T> const T& at(int i) const { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }
    Mat getMat() const { return *this; }

    // views sharing the buffer
    Mat rowRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * step; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * elemSize(); m.cols = b - a; return m; }
    Mat row(int y) const { return rowRange(y, y + 1); }
    Mat col(int x) const { return colRange(x, x + 1); }
    Mat clone() const
    {
        Mat m;
        if (!empty()) {
            m.create(rows, cols, type_);
            for (int y = 0; y < rows; y++) memcpy(m.ptr(y), ptr(y), (size_t)cols * elemSize());
        }
        return m;
    }
    void copyTo(Mat& dst) const
    {
        if (dst.rows != rows || dst.cols != cols || dst.type_ != type_ || !dst.data) dst.create(rows, cols, type_);
        for (int y = 0; y < rows; y++) memmove(dst.ptr(y), ptr(y), (size_t)cols * elemSize());
    }
    void copyTo(Mat&& dst) const { Mat& d = dst; copyTo(d); }          // into a temporary view (rowRange / colRange)
    static Mat eye(int r, int c, int type)
    {
        Mat m = zeros(r, c, type);
        for (int i = 0; i < r && i < c; i++) m.at<float>(i, i) = 1.f;
        return m;
    }
    static Mat zeros(int r, int c, int type)
    {
        Mat m(r, c, type);
        memset(m.data, 0, (size_t)r * m.step);
        return m;
    }
    Mat t() const
    {
        Mat m(cols, rows, type_);
        for (int i = 0; i < rows; i++)
            for (int j = 0; j < cols; j++) m.at<float>(j, i) = at<float>(i, j);
        return m;
    }
    double dot(const Mat& b) const
    {
        double s = 0;
        for (int i = 0; i < rows; i++)
            for (int j = 0; j < cols; j++) s += (double)at<float>(i, j) * b.at<float>(i, j);
        return s;
    }
private:
    std::shared_ptr<std::vector<unsigned char> > buf_;
    int type_ = CV_8U;
};

inline Mat operator*(const Mat& a, const Mat& b)
{
    Mat c(a.rows, b.cols, CV_32F);
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < b.cols; j++) {
            float s = 0.f;
            for (int k = 0; k < a.cols; k++) {
                const float p = a.at<float>(i, k) * b.at<float>(k, j);
                s = k == 0 ? p : s + p;
            }
            c.at<float>(i, j) = s;
        }
    return c;
}
inline Mat operator+(const Mat& a, const Mat& b)
{
    Mat c(a.rows, a.cols, CV_32F);
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < a.cols; j++) c.at<float>(i, j) = (float)((double)a.at<float>(i, j) + (double)b.at<float>(i, j));
    return c;
}
inline Mat operator-(const Mat& a, const Mat& b)
{
    Mat c(a.rows, a.cols, CV_32F);
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < a.cols; j++) c.at<float>(i, j) = a.at<float>(i, j) - b.at<float>(i, j);
    return c;
}
inline Mat operator-(const Mat& a)
{
    Mat c(a.rows, a.cols, CV_32F);
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < a.cols; j++) c.at<float>(i, j) = -a.at<float>(i, j);
    return c;
}
inline Mat scaled(const Mat& a, double alpha)
{
    const float f = (float)alpha;
    Mat c(a.rows, a.cols, CV_32F);
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < a.cols; j++) c.at<float>(i, j) = a.at<float>(i, j) * f;
    return c;
}
inline Mat operator/(const Mat& a, double s) { return scaled(a, 1.0 / s); }
inline Mat operator*(double s, const Mat& a) { return scaled(a, s); }
inline Mat operator*(const Mat& a, double s) { return scaled(a, s); }
inline double norm(const Mat& a) { return std::sqrt(a.dot(a)); }

typedef const Mat& InputArray;
typedef Mat& OutputArray;

}  // namespace cv
#endif
