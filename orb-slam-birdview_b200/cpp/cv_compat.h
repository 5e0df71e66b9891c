// Minimal stand-ins for the OpenCV types in ORBextractor's signature, used ONLY where the OpenCV C++ SDK is
// absent (this image).  With OpenCV installed, compile with -DORBB200_HAVE_OPENCV and the real headers are
// used instead; the shim code in ORBextractor.h is identical in both cases.
#pragma once
#ifdef ORBB200_HAVE_OPENCV
#include <opencv2/core/core.hpp>
#include <opencv2/features2d/features2d.hpp>
#else
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5

namespace cv {

struct Point2f { float x = 0, y = 0; };

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
    Mat(int r, int c, int type, void* ext, size_t stp) : rows(r), cols(c), step(stp), data((unsigned char*)ext), type_(type) {}
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
    template <class T> T& at(int r, int c) { return reinterpret_cast<T*>(data + (size_t)r * step)[c]; }
    template <class T> const T& at(int r, int c) const { return reinterpret_cast<const T*>(data + (size_t)r * step)[c]; }
    template <class T> T& at(int i) { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }
    template <class T> const T& at(int i) const { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }
    Mat getMat() const { return *this; }
private:
    std::shared_ptr<std::vector<unsigned char> > buf_;
    int type_ = CV_8U;
};

// CV_32F products and sums of the small matrices the adapters use (R*x + t), evaluated like OpenCV's gemm small-matrix
// path: float products summed left to right in float (oracle/match_oracle.cpp documents the pin against cv2.gemm)
inline Mat operator*(const Mat& a, const Mat& b)
{
    Mat c(a.rows, b.cols, CV_32F);
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < b.cols; j++) {
            float s = 0.f;
            for (int k = 0; k < a.cols; k++) s += a.at<float>(i, k) * b.at<float>(k, j);
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

typedef const Mat& InputArray;
typedef Mat& OutputArray;

}  // namespace cv
#endif
