/* ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * Stand-in for the slice of the OpenCV C++ API that the reference's front-end sources touch, so that
 * /root/reference/src/ORBextractor.cc (and, with ref_matcher_capi.cpp, src/ORBmatcher.cc + src/Frame.cc members)
 * compile UNMODIFIED in an image that has no OpenCV SDK.  The reference sources are compiled where they lie
 * (oracle/ref_standin/Makefile); only this header set and the small C wrappers are ours.
 *
 * Arithmetic primitives (cv::FAST, cv::resize, cv::GaussianBlur, cv::fastAtan2) are NOT re-derived here:
 * cv_standin.cpp forwards them to the restatements in oracle/orb_oracle.cpp, which tests/test_oracle.py
 * pins bit-exactly on cv2 4.13.  Container semantics that the reference relies on are reproduced:
 *   - cv::Mat is a reference-counted header; operator()(Rect)/rowRange/colRange share the buffer;
 *   - OutputArray::create / Mat::create keep the buffer when size and type already match
 *     (ORBextractor.cc:1120 resizes INTO a ROI of `temp`, :1122 borders in place);
 *   - `m = Mat::zeros(r,c,t)` on a matching header zero-fills IN PLACE (MatExpr assignment;
 *     ORBextractor.cc:1037 relies on it to write descriptors through a rowRange view).
 */
#ifndef ORACLE_CV_STANDIN_HPP
#define ORACLE_CV_STANDIN_HPP

#include <assert.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <emmintrin.h>
#include <algorithm>
#include <vector>

#define CV_PI 3.1415926535897932384626433832795
#define CV_CN_SHIFT 3
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

typedef unsigned char uchar;
typedef unsigned short ushort;

/* cvRound = round half to even (SSE cvtss2si / cvtsd2si), like OpenCV's x86 build */
static inline int cvRound(float v) { return _mm_cvtss_si32(_mm_set_ss(v)); }
static inline int cvRound(double v) { return _mm_cvtsd_si32(_mm_set_sd(v)); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }

namespace cv {

using ::uchar;
using ::ushort;

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4,
       BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template <typename U> explicit Point_(const Point_<U>& p) : x((T)p.x), y((T)p.y) {}
};
template <typename T> static inline Point_<T>& operator*=(Point_<T>& a, float b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
template <typename T> static inline Point_<T>& operator*=(Point_<T>& a, double b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
template <typename T> static inline Point_<T>& operator*=(Point_<T>& a, int b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
template <typename T> static inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <typename T> static inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T _x, T _y, T _z) : x(_x), y(_y), z(_z) {}
};
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T _x, T _y, T w, T h) : x(_x), y(_y), width(w), height(h) {}
};
typedef Rect_<int> Rect;

template <typename T> struct Scalar_ {
    T val[4];
    Scalar_() { val[0] = val[1] = val[2] = val[3] = 0; }
    Scalar_(T v0, T v1 = 0, T v2 = 0, T v3 = 0) { val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
};
typedef Scalar_<double> Scalar;

struct Range { int start, end; Range() : start(0), end(0) {} Range(int s, int e) : start(s), end(e) {} };

/* binary layout of the real cv::KeyPoint (28 bytes) */
class KeyPoint {
public:
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f _pt, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(_pt), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};

class Mat;
struct MatZerosExpr { int rows, cols, type; };

class Mat {
    struct Buf { int refs; };   /* header of a malloc'ed block: [Buf][pad to 64][pixels] */
    Buf* buf_;
public:
    struct Step {
        size_t v;
        Step() : v(0) {}
        operator size_t() const { return v; }
        Step& operator=(size_t s) { v = s; return *this; }
    };
    int flags;          /* = type */
    int dims;
    int rows, cols;
    uchar* data;
    Step step;

    Mat() : buf_(0), flags(0), dims(2), rows(0), cols(0), data(0) {}
    Mat(int r, int c, int t) : buf_(0), flags(0), dims(2), rows(0), cols(0), data(0) { create(r, c, t); }
    Mat(Size s, int t) : buf_(0), flags(0), dims(2), rows(0), cols(0), data(0) { create(s.height, s.width, t); }
    /* header over user data, no ownership */
    Mat(int r, int c, int t, void* d, size_t st = 0) : buf_(0), flags(t), dims(2), rows(r), cols(c), data((uchar*)d)
    { step = st ? st : (size_t)c * elemSize(); }
    Mat(const Mat& m) : buf_(m.buf_), flags(m.flags), dims(2), rows(m.rows), cols(m.cols), data(m.data), step(m.step)
    { if (buf_) __sync_fetch_and_add(&buf_->refs, 1); }
    Mat(const MatZerosExpr& e) : buf_(0), flags(0), dims(2), rows(0), cols(0), data(0) { *this = e; }
    ~Mat() { release(); }
    Mat& operator=(const Mat& m)
    {
        if (this != &m) {
            if (m.buf_) __sync_fetch_and_add(&m.buf_->refs, 1);
            release();
            buf_ = m.buf_; flags = m.flags; rows = m.rows; cols = m.cols; data = m.data; step = m.step;
        }
        return *this;
    }
    /* MatExpr assignment of Mat::zeros: create() (no-op on a matching header) then fill in place */
    Mat& operator=(const MatZerosExpr& e)
    {
        create(e.rows, e.cols, e.type);
        for (int y = 0; y < rows; y++) memset(data + (size_t)y * step, 0, (size_t)cols * elemSize());
        return *this;
    }
    static MatZerosExpr zeros(int r, int c, int t) { MatZerosExpr e = {r, c, t}; return e; }
    static MatZerosExpr zeros(Size s, int t) { MatZerosExpr e = {s.height, s.width, t}; return e; }

    void release()
    {
        if (buf_ && __sync_fetch_and_add(&buf_->refs, -1) == 1) free(buf_);
        buf_ = 0; data = 0; rows = cols = 0; step = 0;
    }
    void create(int r, int c, int t)
    {
        if (data && r == rows && c == cols && t == flags) return;
        release();
        flags = t; rows = r; cols = c;
        step = (size_t)c * elemSize();
        size_t bytes = (size_t)r * step;
        buf_ = (Buf*)malloc(64 + bytes + 64);       /* malloc, not operator new: see ref_capi.cpp's allocator switch */
        buf_->refs = 1;
        data = (uchar*)buf_ + 64;
    }
    void create(Size s, int t) { create(s.height, s.width, t); }

    int type() const { return flags; }
    int depth() const { return flags & 7; }
    int channels() const { return (flags >> CV_CN_SHIFT) + 1; }
    size_t elemSize1() const { static const int sz[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return sz[flags & 7]; }
    size_t elemSize() const { return elemSize1() * channels(); }
    size_t step1() const { return step.v / elemSize1(); }
    bool empty() const { return data == 0 || rows == 0 || cols == 0; }
    Size size() const { return Size(cols, rows); }
    size_t total() const { return (size_t)rows * cols; }
    bool isContinuous() const { return step.v == (size_t)cols * elemSize(); }

    Mat operator()(const Rect& r) const
    {
        assert(r.x >= 0 && r.y >= 0 && r.x + r.width <= cols && r.y + r.height <= rows);
        Mat m(*this);
        m.data = data + (size_t)r.y * step + (size_t)r.x * elemSize();
        m.rows = r.height; m.cols = r.width;
        return m;
    }
    Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
    Mat row(int y) const { return rowRange(y, y + 1); }
    Mat col(int x) const { return colRange(x, x + 1); }
    Mat clone() const
    {
        Mat m;
        if (!empty()) {
            m.create(rows, cols, flags);
            for (int y = 0; y < rows; y++) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * elemSize());
        }
        return m;
    }
    void copyTo(Mat& dst) const
    {
        dst.create(rows, cols, flags);
        for (int y = 0; y < rows; y++) memmove(dst.data + (size_t)y * dst.step, data + (size_t)y * step, (size_t)cols * elemSize());
    }

    uchar* ptr(int y = 0) { return data + (size_t)y * step; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
    template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
    template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
    template <typename T> T& at(int y, int x) { return ((T*)(data + (size_t)y * step))[x]; }
    template <typename T> const T& at(int y, int x) const { return ((const T*)(data + (size_t)y * step))[x]; }
    /* single-index access on a row or column vector */
    template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
};

class _InputArray {
    const Mat* m_;
public:
    _InputArray() : m_(0) {}
    _InputArray(const Mat& m) : m_(&m) {}
    bool empty() const { return !m_ || m_->empty(); }
    Mat getMat() const { return m_ ? *m_ : Mat(); }
};
class _OutputArray {
    Mat* m_;
public:
    _OutputArray() : m_(0) {}
    _OutputArray(Mat& m) : m_(&m) {}
    bool needed() const { return m_ != 0; }
    void create(int r, int c, int t) const { m_->create(r, c, t); }
    void create(Size s, int t) const { m_->create(s.height, s.width, t); }
    void release() const { if (m_) m_->release(); }
    Mat getMat() const { return *m_; }
    Mat& getMatRef() const { return *m_; }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
static inline InputArray noArray() { static _InputArray none; return none; }

/* ---- primitives, implemented in cv_standin.cpp on the cv2-pinned restatements of oracle/orb_oracle.cpp ---- */
void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType,
                    const Scalar& value = Scalar());
float fastAtan2(float y, float x);

/* only reached from ORBextractor::ComputeKeyPointsOld, which operator() never calls (ORBextractor.cc:1057) */
struct KeyPointsFilter { static void retainBest(std::vector<KeyPoint>& keypoints, int npoints); };

}  // namespace cv

#endif
