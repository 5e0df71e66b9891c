// Device helpers shared by the extraction translation units: the rBRIEF pattern, the IC_Angle row table and
// cv::fastAtan2.  (Each translation unit gets its own copy of the small constant tables.)
#pragma once
#include <cuda_runtime.h>

#include "../../include/orbb200_pattern.inc"

namespace orbb200 {

static __constant__ signed char c_patX[512] = {ORBB200_PATTERN_X_INIT};
static __constant__ signed char c_patY[512] = {ORBB200_PATTERN_Y_INIT};
// reference src/ORBextractor.cc:454-469; cv::ORB (orb.cpp computeKeyPoints) builds the same table for patchSize 31
static __constant__ int c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    // cv::fastAtan2, float32 without FMA contraction
    constexpr float k180pi = (float)(180.0 / 3.14159265358979323846);
    constexpr float p1 = 0.9997878412794807f * k180pi, p3 = -0.3258083974640975f * k180pi;
    constexpr float p5 = 0.1555786518463281f * k180pi, p7 = -0.04432655554792128f * k180pi;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    const float eps = 2.22044605e-16f;   // (float)DBL_EPSILON
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

}  // namespace orbb200
