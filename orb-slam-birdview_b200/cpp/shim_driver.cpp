// Test driver for the C++ shim: runs ORB_SLAM2::ORBextractor::operator() exactly as Frame::ExtractORB does
// (src/Frame.cc:414-420) on a raw gray image file and dumps keypoints + descriptors.
//   shim_driver <in.raw> <w> <h> <nfeatures> <iniTh> <minTh> <out.bin>
// With two more arguments it then runs the birdview block of Frame::Frame (src/Frame.cc:328-342) through
// BirdviewExtractor.h on a second image + mask and appends those keypoints + descriptors:
//   shim_driver ... <out.bin> <bird.raw> <bird_mask.raw> <bw> <bh>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "BirdviewExtractor.h"
#include "ORBextractor.h"

int main(int argc, char** argv)
{
    if (argc < 8) { fprintf(stderr, "usage\n"); return 2; }
    const int w = atoi(argv[2]), h = atoi(argv[3]);
    std::vector<unsigned char> buf((size_t)w * h);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(buf.data(), 1, buf.size(), f) != buf.size()) { fprintf(stderr, "cannot read image\n"); return 2; }
    fclose(f);
    cv::Mat im(h, w, CV_8UC1, buf.data(), (size_t)w);
    ORB_SLAM2::ORBextractor extractor(atoi(argv[4]), 1.2f, 8, atoi(argv[5]), atoi(argv[6]));
    std::vector<cv::KeyPoint> keys;
    cv::Mat desc;
    extractor(im, cv::Mat(), keys, desc);
    extractor(im, cv::Mat(), keys, desc);      // second call: vectors are cleared and refilled (:1072-1073)
    FILE* o = fopen(argv[7], "wb");
    const int n = (int)keys.size();
    fwrite(&n, 4, 1, o);
    fwrite(keys.data(), sizeof(cv::KeyPoint), n, o);
    for (int i = 0; i < n; i++) fwrite(desc.ptr(i), 1, 32, o);
    const int pw = extractor.mvImagePyramid[1].cols, ph = extractor.mvImagePyramid[1].rows;
    fwrite(&pw, 4, 1, o); fwrite(&ph, 4, 1, o);
    for (int y = 0; y < ph; y++) fwrite(extractor.mvImagePyramid[1].ptr(y), 1, pw, o);
    // ORBB200_SHIM_DUMP_PYRAMID=<file>: every level of mvImagePyramid as {w, h, rows} (the pinned mirror the extraction's kernels fill)
    if (const char* pf = getenv("ORBB200_SHIM_DUMP_PYRAMID")) {
        FILE* po = fopen(pf, "wb");
        const int nl = (int)extractor.mvImagePyramid.size();
        fwrite(&nl, 4, 1, po);
        for (int l = 0; l < nl; l++) {
            const cv::Mat& m = extractor.mvImagePyramid[l];
            const int lw = m.cols, lh = m.rows;
            fwrite(&lw, 4, 1, po); fwrite(&lh, 4, 1, po);
            for (int y = 0; y < lh; y++) fwrite(m.ptr(y), 1, lw, po);
        }
        fclose(po);
    }
    if (argc >= 12) {
        const int bw = atoi(argv[10]), bh = atoi(argv[11]);
        std::vector<unsigned char> bimg((size_t)bw * bh), bmask((size_t)bw * bh);
        FILE* fi = fopen(argv[8], "rb");
        FILE* fm = fopen(argv[9], "rb");
        if (!fi || !fm || fread(bimg.data(), 1, bimg.size(), fi) != bimg.size() || fread(bmask.data(), 1, bmask.size(), fm) != bmask.size()) {
            fprintf(stderr, "cannot read birdview image\n"); return 2;
        }
        fclose(fi); fclose(fm);
        cv::Mat mBirdviewImg(bh, bw, CV_8UC1, bimg.data(), (size_t)bw), mBirdviewMask(bh, bw, CV_8UC1, bmask.data(), (size_t)bw);
        ORB_SLAM2::BirdviewExtractor extractorBird(extractor.Context(), 2000);
        std::vector<cv::KeyPoint> mvKeysBird;
        cv::Mat mDescriptorsBird;
        extractorBird(mBirdviewImg, mBirdviewMask, mvKeysBird, mDescriptorsBird);
        const int nb = (int)mvKeysBird.size();
        fwrite(&nb, 4, 1, o);
        fwrite(mvKeysBird.data(), sizeof(cv::KeyPoint), nb, o);
        for (int i = 0; i < nb; i++) fwrite(mDescriptorsBird.ptr(i), 1, 32, o);
        // the same through the three separate calls
        std::vector<cv::KeyPoint> k2;
        extractorBird.detect(mBirdviewImg, k2, mBirdviewMask);
        std::vector<float> pts(2 * k2.size());
        for (size_t i = 0; i < k2.size(); i++) { pts[2 * i] = k2[i].pt.x; pts[2 * i + 1] = k2[i].pt.y; }
        extractorBird.cornerSubPix(mBirdviewImg, pts);
        for (size_t i = 0; i < k2.size(); i++) { k2[i].pt.x = pts[2 * i]; k2[i].pt.y = pts[2 * i + 1]; }
        cv::Mat d2;
        extractorBird.compute(mBirdviewImg, k2, d2);
        int same = k2.size() == mvKeysBird.size();
        for (int i = 0; same && i < nb; i++) same = memcmp(&k2[i], &mvKeysBird[i], sizeof(cv::KeyPoint)) == 0 && memcmp(d2.ptr(i), mDescriptorsBird.ptr(i), 32) == 0;
        fwrite(&same, 4, 1, o);
        printf("%d birdview keypoints, separate calls agree: %d\n", nb, same);
    }
    fclose(o);
    printf("%d keypoints, levels %d, scale[1] %.6f\n", n, extractor.GetLevels(), extractor.GetScaleFactors()[1]);
    // ORBB200_SHIM_TIME=N: wall time of N calls of operator() as Frame::ExtractORB makes them (pageable cv::Mat in, vectors out),
    // with and without the mvImagePyramid refresh -- the drop-in latency bench.py reports as configs.per_frame.cpp_shim
    if (const char* e = getenv("ORBB200_SHIM_TIME")) {
        const int reps = atoi(e) > 0 ? atoi(e) : 100;
        double ms[2];
        for (int mode = 0; mode < 2; mode++) {
            extractor.SetDownloadPyramid(mode == 0);
            extractor(im, cv::Mat(), keys, desc);
            const auto t0 = std::chrono::steady_clock::now();
            for (int i = 0; i < reps; i++) extractor(im, cv::Mat(), keys, desc);
            ms[mode] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / reps;
        }
        printf("TIMING {\"w\": %d, \"h\": %d, \"keypoints\": %d, \"operator_call_ms\": %.4f, \"operator_call_without_pyramid_refresh_ms\": %.4f}\n",
               w, h, (int)keys.size(), ms[0], ms[1]);
    }
    return 0;
}
