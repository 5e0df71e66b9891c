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
    Mat(int r, int c, int /*type*/) { create(r, c, CV_8U); }
    Mat(int r, int c, int /*type*/, void* ext, size_t stp) : rows(r), cols(c), step(stp), data((unsigned char*)ext) {}
    void create(int r, int c, int /*type*/)
    {
        if (r == rows && c == cols && buf_) return;
        rows = r; cols = c; step = (size_t)c;
        buf_.reset(new std::vector<unsigned char>((size_t)r * c));
        data = buf_->data();
    }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return CV_8UC1; }
    unsigned char* ptr(int y = 0) { return data + (size_t)y * step; }
    const unsigned char* ptr(int y = 0) const { return data + (size_t)y * step; }
    Mat getMat() const { return *this; }
private:
    std::shared_ptr<std::vector<unsigned char> > buf_;
};

typedef const Mat& InputArray;
typedef Mat& OutputArray;

}  // namespace cv
#endif
