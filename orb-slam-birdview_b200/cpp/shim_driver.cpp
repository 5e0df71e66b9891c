// Test driver for the C++ shim: runs ORB_SLAM2::ORBextractor::operator() exactly as Frame::ExtractORB does
// (src/Frame.cc:414-420) on a raw gray image file and dumps keypoints + descriptors.
//   shim_driver <in.raw> <w> <h> <nfeatures> <iniTh> <minTh> <out.bin>
#include <cstdio>
#include <cstdlib>
#include <vector>

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
    fclose(o);
    printf("%d keypoints, levels %d, scale[1] %.6f\n", n, extractor.GetLevels(), extractor.GetScaleFactors()[1]);
    return 0;
}
