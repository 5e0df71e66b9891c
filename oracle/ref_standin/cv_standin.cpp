/* ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * The OpenCV primitives the reference's ORBextractor.cc calls, forwarded to the restatements in
 * oracle/orb_oracle.cpp that tests/test_oracle.py pins bit-exactly on cv2 4.13 (resize INTER_LINEAR, GaussianBlur
 * 7x7 sigma 2, FAST 9/16 + NMS, fastAtan2).  copyMakeBorder is written here: it is pure copying. */
#include "cv_standin.hpp"
#include "../orb_oracle.h"

namespace cv {

/* cv::FAST(image, keypoints, threshold, nms) -- ORBextractor.cc:809,814; KeyPoint(x, y, 7.f, -1, score) in row-major order */
void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression)
{
    Mat img = image.getMat();
    assert(img.type() == CV_8UC1);
    keypoints.clear();
    if (img.empty()) return;
    int cap = ((img.cols + 1) / 2) * ((img.rows + 1) / 2) + 16;
    if (!nonmaxSuppression) cap = img.cols * img.rows;
    std::vector<int32_t> xyr((size_t)cap * 3);
    int n = oracle_fast9(img.data, img.cols, img.rows, img.step, threshold, nonmaxSuppression ? 1 : 0, xyr.data(), cap);
    assert(n <= cap);
    keypoints.reserve(n);
    for (int i = 0; i < n; i++)
        keypoints.push_back(KeyPoint((float)xyr[3 * i], (float)xyr[3 * i + 1], 7.f, -1.f, (float)xyr[3 * i + 2]));
}

/* cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR) on CV_8UC1 -- ORBextractor.cc:1120.  dst.create() keeps a matching ROI. */
void resize(InputArray _src, OutputArray _dst, Size dsize, double fx, double fy, int interpolation)
{
    Mat src = _src.getMat();
    assert(src.type() == CV_8UC1 && interpolation == INTER_LINEAR && dsize.width > 0 && dsize.height > 0);
    (void)fx; (void)fy;
    _dst.create(dsize.height, dsize.width, src.type());
    Mat dst = _dst.getMat();
    assert(dst.data != src.data);
    oracle_resize_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}

/* cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) on CV_8UC1 -- ORBextractor.cc:1086 (in place there) */
void GaussianBlur(InputArray _src, OutputArray _dst, Size ksize, double sigmaX, double sigmaY, int borderType)
{
    Mat src = _src.getMat();
    assert(src.type() == CV_8UC1 && ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && sigmaY == 2);
    assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
    Mat tmp = src.clone();                       /* in-place call: filter from a copy */
    _dst.create(src.rows, src.cols, src.type());
    Mat dst = _dst.getMat();
    oracle_gauss7_u8(tmp.data, tmp.cols, tmp.rows, tmp.step, dst.data, dst.step);
}

static inline int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

/* cv::copyMakeBorder(src, dst, t, b, l, r, BORDER_REFLECT_101 [+ BORDER_ISOLATED]) -- ORBextractor.cc:1122,1127.
 * dst.create() keeps a matching buffer, so when src is the interior ROI of dst the interior is left where it is and
 * only the border is filled (what OpenCV does too). */
void copyMakeBorder(InputArray _src, OutputArray _dst, int top, int bottom, int left, int right, int borderType, const Scalar&)
{
    Mat src = _src.getMat();
    assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
    const int es = (int)src.elemSize();
    _dst.create(src.rows + top + bottom, src.cols + left + right, src.type());
    Mat dst = _dst.getMat();
    /* interior */
    for (int y = 0; y < src.rows; y++) {
        uchar* d = dst.ptr(y + top) + (size_t)left * es;
        const uchar* s = src.ptr(y);
        if (d != s) memmove(d, s, (size_t)src.cols * es);
    }
    /* left/right of the interior rows, from the interior itself */
    for (int y = 0; y < src.rows; y++) {
        uchar* row = dst.ptr(y + top);
        uchar* in = row + (size_t)left * es;
        for (int x = 0; x < left; x++) memcpy(row + (size_t)x * es, in + (size_t)reflect101(x - left, src.cols) * es, es);
        for (int x = 0; x < right; x++) memcpy(in + (size_t)(src.cols + x) * es, in + (size_t)reflect101(src.cols + x, src.cols) * es, es);
    }
    /* top/bottom: whole bordered rows */
    for (int y = 0; y < top; y++) memcpy(dst.ptr(y), dst.ptr(top + reflect101(y - top, src.rows)), (size_t)dst.cols * es);
    for (int y = 0; y < bottom; y++)
        memcpy(dst.ptr(top + src.rows + y), dst.ptr(top + reflect101(src.rows + y, src.rows)), (size_t)dst.cols * es);
}

float fastAtan2(float y, float x) { return oracle_fast_atan2(y, x); }

/* cv::KeyPointsFilter::retainBest -- dead code on the path (ComputeKeyPointsOld); OpenCV's algorithm in brief */
void KeyPointsFilter::retainBest(std::vector<KeyPoint>& kps, int n)
{
    if (n >= 0 && kps.size() > (size_t)n) {
        if (n == 0) { kps.clear(); return; }
        std::nth_element(kps.begin(), kps.begin() + n - 1, kps.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
        float ambiguous = kps[n - 1].response;
        auto e = std::partition(kps.begin() + n, kps.end(), [ambiguous](const KeyPoint& k) { return k.response >= ambiguous; });
        kps.resize(e - kps.begin());
    }
}

}  // namespace cv
