// Replacement for the birdview feature block of Frame::Frame (/root/reference/src/Frame.cc:328-342):
//
//     cv::Ptr<cv::ORB> extractorBird = cv::ORB::create(2000);
//     extractorBird->detect(mBirdviewImg, mvKeysBird, mBirdviewMask);
//     ... cv::cornerSubPix(mBirdviewImg, vKeysBird, cv::Size(5,5), cv::Size(-1,-1), criteria(40, 0.001)) ...
//     extractorBird->compute(mBirdviewImg, mvKeysBird, mDescriptorsBird);
//
// as one call into liborbb200.so (include/orbb200.h: orbb200_bird_extract).  The three OpenCV calls are also exposed one
// by one (detect / cornerSubPix / compute) with cv::ORB's and cv::cornerSubPix's semantics.  Header-only.
#ifndef ORBB200_BIRDVIEW_EXTRACTOR_H
#define ORBB200_BIRDVIEW_EXTRACTOR_H

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cv_compat.h"
#include "../../include/orbb200.h"

namespace ORB_SLAM2
{

class BirdviewExtractor
{
public:
    // ctx: a context of the calling thread (e.g. ORBextractor::Context()); nfeatures as in cv::ORB::create(nfeatures)
    explicit BirdviewExtractor(orbb200_ctx* ctx_, int nfeatures_ = 2000) : ctx(ctx_), nfeatures(nfeatures_) {}

    // detect(mask) + cornerSubPix(5x5, 40 it, 1e-3) + compute: fills mvKeysBird / mDescriptorsBird
    void operator()(const cv::Mat& image, const cv::Mat& mask, std::vector<cv::KeyPoint>& keypoints, cv::Mat& descriptors)
    {
        static_assert(sizeof(cv::KeyPoint) == sizeof(orbb200_kp_t), "cv::KeyPoint layout");
        if (image.empty()) { keypoints.clear(); descriptors.release(); return; }
        const int cap = capacity(image);
        keypoints.resize(cap);
        buf.resize((size_t)cap * 32);
        int n = 0;
        check(orbb200_bird_extract(ctx, image.ptr(0), mask.empty() ? nullptr : mask.ptr(0), image.cols, image.rows, image.step,
                                   mask.empty() ? 0 : mask.step, nfeatures, reinterpret_cast<orbb200_kp_t*>(keypoints.data()), buf.data(),
                                   cap, &n));
        keypoints.resize(n);
        fill(descriptors, n);
    }

    // cv::ORB::detect(image, keypoints, mask)
    void detect(const cv::Mat& image, std::vector<cv::KeyPoint>& keypoints, const cv::Mat& mask)
    {
        const int cap = capacity(image);
        keypoints.resize(cap);
        int n = 0;
        check(orbb200_bird_detect(ctx, image.ptr(0), mask.empty() ? nullptr : mask.ptr(0), image.cols, image.rows, image.step,
                                  mask.empty() ? 0 : mask.step, nfeatures, reinterpret_cast<orbb200_kp_t*>(keypoints.data()), cap, &n));
        keypoints.resize(n);
    }

    // cv::cornerSubPix(image, corners, Size(win, win), Size(-1,-1), TermCriteria(EPS + MAX_ITER, maxCount, epsilon)); corners = x0,y0,x1,y1,...
    void cornerSubPix(const cv::Mat& image, std::vector<float>& corners, int win = 5, int maxCount = 40, double epsilon = 0.001)
    {
        check(orbb200_corner_subpix(ctx, image.ptr(0), image.cols, image.rows, image.step, corners.data(), (int)(corners.size() / 2), win, win,
                                    maxCount, epsilon));
    }

    // cv::ORB::compute(image, keypoints, descriptors)
    void compute(const cv::Mat& image, std::vector<cv::KeyPoint>& keypoints, cv::Mat& descriptors)
    {
        buf.resize(keypoints.size() * 32 + 32);
        int n = 0;
        check(orbb200_bird_compute(ctx, image.ptr(0), image.cols, image.rows, image.step, reinterpret_cast<orbb200_kp_t*>(keypoints.data()),
                                   (int)keypoints.size(), buf.data(), &n));
        keypoints.resize(n);
        fill(descriptors, n);
    }

private:
    int capacity(const cv::Mat& image)
    {
        const int cap = orbb200_bird_max_keypoints(ctx, image.cols, image.rows, nfeatures);
        if (cap <= 0) die(orbb200_last_error(ctx));
        return cap;
    }
    void fill(cv::Mat& descriptors, int n)
    {
        if (n == 0) { descriptors.release(); return; }
        descriptors.create(n, 32, CV_8U);
        for (int i = 0; i < n; i++) memcpy(descriptors.ptr(i), &buf[(size_t)i * 32], 32);
    }
    void check(int rc) { if (rc != ORBB200_OK) die(orbb200_last_error(ctx)); }
    void die(const char* what) { fprintf(stderr, "BirdviewExtractor (orbb200): %s\n", what); abort(); }

    orbb200_ctx* ctx;
    int nfeatures;
    std::vector<unsigned char> buf;
};

}  // namespace ORB_SLAM2
#endif
