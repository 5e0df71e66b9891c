// Drop-in replacement for the reference's include/ORBextractor.h (ORB_SLAM2::ORBextractor,
// /root/reference/include/ORBextractor.h:44-111): same constructor, operator() signature, getters and public
// mvImagePyramid, implemented over the C ABI of liborbb200.so (include/orbb200.h).  Frame.cc / Tracking.cc
// compile against it unchanged.  Header-only; link with -lorbb200.
//
// Differences a maintainer should know:
//  * no CPU fallback: if the CUDA library reports an error the call prints it and aborts, like the
//    reference's assert()/exit(-1) style error handling (src/ORBextractor.cc:1050, src/System.cc:59-84);
//  * the device context is created on the first call (it needs the image size); a larger image re-creates it;
//  * mvImagePyramid is refreshed after every call because Frame::ComputeStereoMatches reads it
//    (src/Frame.cc:669-776): one device-to-host copy of the pyramid block into a pinned mirror, the levels are cv::Mat headers
//    over it and stay valid until the next call (Frame::ComputeStereoMatches runs before it).  SetDownloadPyramid(false) skips
//    the copy (monocular use, or ComputeStereoMatches replaced by orbb200_compute_stereo_matches).
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <cstdio>
#include <cstdlib>
#include <list>
#include <vector>

#include "cv_compat.h"
#include "../../include/orbb200.h"

namespace ORB_SLAM2
{

class ORBextractor
{
public:

    enum {HARRIS_SCORE=0, FAST_SCORE=1 };

    ORBextractor(int nfeatures_, float scaleFactor_, int nlevels_, int iniThFAST_, int minThFAST_)
        : nfeatures(nfeatures_), scaleFactor(scaleFactor_), nlevels(nlevels_), iniThFAST(iniThFAST_), minThFAST(minThFAST_)
    {
        // scale tables as in src/ORBextractor.cc:415-431 (available before the first image arrives)
        mvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels);
        mvInvScaleFactor.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
        mvScaleFactor[0] = 1.0f; mvLevelSigma2[0] = 1.0f;
        for (int i = 1; i < nlevels; i++) {
            mvScaleFactor[i] = (float)(mvScaleFactor[i - 1] * scaleFactor);
            mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
        }
        for (int i = 0; i < nlevels; i++) {
            mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i];
            mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i];
        }
        mvImagePyramid.resize(nlevels);
    }

    ~ORBextractor() { if (ctx) orbb200_destroy(ctx); }

    // Compute the ORB features and descriptors on an image.
    // Mask is ignored, as in the reference (include/ORBextractor.h:58).
    void operator()( cv::InputArray _image, cv::InputArray /*_mask*/,
      std::vector<cv::KeyPoint>& _keypoints,
      cv::OutputArray _descriptors)
    {
        if (_image.empty())
            return;
        cv::Mat image = _image.getMat();
        if (image.type() != CV_8UC1) die("image.type() == CV_8UC1");
        ensureContext(image.cols, image.rows);
        const int cap = orbb200_max_keypoints(ctx);
        static_assert(sizeof(cv::KeyPoint) == sizeof(orbb200_kp_t), "cv::KeyPoint layout");
        kpbuf.resize(cap);
        descbuf.resize((size_t)cap * 32);
        int n = 0;
        if (bDownloadPyramid != bMirrorSet) { check(orbb200_set_pyramid_mirror(ctx, bDownloadPyramid ? 1 : 0)); bMirrorSet = bDownloadPyramid; }
        check(orbb200_extract(ctx, image.ptr(0), image.cols, image.rows, image.step,
                              reinterpret_cast<orbb200_kp_t*>(kpbuf.data()), descbuf.data(), cap, &n));
        if (n == 0)
            _descriptors.release();
        else {
            _descriptors.create(n, 32, CV_8U);
            cv::Mat d = _descriptors.getMat();
            for (int i = 0; i < n; i++) memcpy(d.ptr(i), &descbuf[(size_t)i * 32], 32);
        }
        _keypoints.assign(kpbuf.begin(), kpbuf.begin() + n);
        if (bDownloadPyramid) {
            // one copy of the whole pyramid block into the context's pinned mirror; mvImagePyramid[l] are headers over it (like the
            // reference's ROIs into its bordered temporaries, src/ORBextractor.cc:1116-1128), valid until the next call
            const uint8_t* lp[16]; size_t pitch[16]; int lw[16], lh[16];
            check(orbb200_pyramid_mirror(ctx, 0, 0, lp, pitch, lw, lh));
            for (int l = 0; l < nlevels; l++) {
                if (!lp[l]) { mvImagePyramid[l].release(); continue; }
                mvImagePyramid[l] = cv::Mat(lh[l], lw[l], CV_8UC1, const_cast<unsigned char*>(lp[l]), pitch[l]);
            }
        }
    }

    int inline GetLevels(){
        return nlevels;}

    float inline GetScaleFactor(){
        return scaleFactor;}

    std::vector<float> inline GetScaleFactors(){
        return mvScaleFactor;
    }

    std::vector<float> inline GetInverseScaleFactors(){
        return mvInvScaleFactor;
    }

    std::vector<float> inline GetScaleSigmaSquares(){
        return mvLevelSigma2;
    }

    std::vector<float> inline GetInverseScaleSigmaSquares(){
        return mvInvLevelSigma2;
    }

    std::vector<cv::Mat> mvImagePyramid;

    // --- additions (not in the reference) ---
    void SetDownloadPyramid(bool b) { bDownloadPyramid = b; }
    void SetDevice(int d) { device = d; }
    orbb200_ctx* Context() { return ctx; }     // for ORBmatcher shims that keep descriptors on the device

protected:

    void ensureContext(int w, int h)
    {
        if (ctx && w <= ctxW && h <= ctxH) return;
        if (ctx) orbb200_destroy(ctx);
        ctx = nullptr;
        ctxW = w > ctxW ? w : ctxW; ctxH = h > ctxH ? h : ctxH;
        if (orbb200_create(&ctx, device, nfeatures, (float)scaleFactor, nlevels, iniThFAST, minThFAST, ctxW, ctxH, 1) != ORBB200_OK) {
            fprintf(stderr, "ORBextractor (orbb200): %s\n", orbb200_last_error(nullptr));
            exit(-1);
        }
        bMirrorSet = false;
    }
    void check(int rc) { if (rc != ORBB200_OK) die(orbb200_last_error(ctx)); }
    void die(const char* what) { fprintf(stderr, "ORBextractor (orbb200): %s\n", what); abort(); }

    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;

    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;

    orbb200_ctx* ctx = nullptr;
    int ctxW = 0, ctxH = 0, device = 0;
    bool bDownloadPyramid = true;
    bool bMirrorSet = false;         // what the context was last told (orbb200_set_pyramid_mirror); a new context starts with "off"
    std::vector<cv::KeyPoint> kpbuf;
    std::vector<unsigned char> descbuf;
};

} //namespace ORB_SLAM

#endif
