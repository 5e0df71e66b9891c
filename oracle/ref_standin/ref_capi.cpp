/* ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * C wrapper around the REFERENCE's own ORB_SLAM2::ORBextractor, compiled unmodified from
 * /root/reference/src/ORBextractor.cc against the stand-in OpenCV headers of this directory
 * (oracle/ref_standin/Makefile -> oracle/_ref/libref_orbextractor*.so).  Used by tests/test_oracle_ref.py to pin
 * oracle/orb_oracle.cpp on the reference's code, and by bench.py's reference arm (cpu_baseline.kind "reference").
 *
 * Two builds of the same sources:
 *   libref_orbextractor.so       glibc malloc -- the reference binary's real behaviour, including the heap-address
 *                                tie-break of sort(vector<pair<int,ExtractorNode*>>) (src/ORBextractor.cc:684);
 *   libref_orbextractor_bump.so  -DREF_BUMP_ALLOC: operator new hands out monotonically increasing addresses and
 *                                never reuses one, which makes that tie-break "most recently created node first",
 *                                the rule oracle/orb_oracle.cpp and the CUDA octree implement.  With it the
 *                                unmodified reference must equal the oracle on EVERY output byte.
 */
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <new>
#include <vector>

#include "ORBextractor.h"   /* the reference's include/ORBextractor.h */

#ifdef REF_BUMP_ALLOC
#include <sys/mman.h>
/* One big virtual reservation, bump pointer, no reuse.  ref_extract() rewinds to its entry mark when it returns
 * (everything allocated inside a call is dead by then: cv::Mat buffers use malloc, not operator new).
 * Single-threaded by construction: this build is only used by the parity tests. */
static char* g_base = 0;
static size_t g_off = 0;
static const size_t ARENA = (size_t)24 << 30;
static void* bump(size_t n)
{
    if (!g_base) {
        g_base = (char*)mmap(0, ARENA, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (g_base == (char*)MAP_FAILED) { fprintf(stderr, "ref_capi: arena mmap failed\n"); abort(); }
    }
    n = (n + 15) & ~(size_t)15;
    if (g_off + n > ARENA) { fprintf(stderr, "ref_capi: arena exhausted\n"); abort(); }
    void* p = g_base + g_off;
    g_off += n;
    return p;
}
void* operator new(size_t n) { return bump(n); }
void* operator new[](size_t n) { return bump(n); }
void operator delete(void*) noexcept {}
void operator delete[](void*) noexcept {}
void operator delete(void*, size_t) noexcept {}
void operator delete[](void*, size_t) noexcept {}
#define ARENA_MARK() size_t arena_mark = g_off
static void arena_rewind(size_t mark)
{
    size_t lo = (mark + 4095) & ~(size_t)4095;      /* whole pages above the mark: give them back */
    if (g_off > lo && g_off - lo > ((size_t)64 << 20)) madvise(g_base + lo, (g_off - lo) & ~(size_t)4095, MADV_DONTNEED);
    g_off = mark;
}
#define ARENA_REWIND() arena_rewind(arena_mark)
#else
#define ARENA_MARK() (void)0
#define ARENA_REWIND() (void)0
#endif

namespace {
struct RefKp { float x, y, size, angle, response; int32_t octave, class_id; };
static_assert(sizeof(cv::KeyPoint) == 28 && sizeof(RefKp) == 28, "cv::KeyPoint layout");

/* protected members of the reference class, reached the C++ way */
struct Access : public ORB_SLAM2::ORBextractor {
    using ORB_SLAM2::ORBextractor::ORBextractor;
    const std::vector<int>& quotas() const { return mnFeaturesPerLevel; }
    const std::vector<int>& um() const { return umax; }
    std::vector<cv::KeyPoint> octree(const std::vector<cv::KeyPoint>& k, int minX, int maxX, int minY, int maxY, int N)
    { return DistributeOctTree(k, minX, maxX, minY, maxY, N, 0); }
};
}  // namespace

extern "C" {

int ref_is_bump_alloc(void)
{
#ifdef REF_BUMP_ALLOC
    return 1;
#else
    return 0;
#endif
}

void* ref_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST)
{ return new Access(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST); }

void ref_extractor_destroy(void* e) { delete (Access*)e; }

/* ORBextractor::operator()(image, mask, keypoints, descriptors); returns N or -1 when cap is too small */
int ref_extract(void* e, const uint8_t* img, int w, int h, size_t step, void* kps_out, uint8_t* desc_out, int cap)
{
    Access* E = (Access*)e;
    int n;
    ARENA_MARK();
    {
        cv::Mat image(h, w, CV_8UC1, (void*)img, step), descriptors;
        std::vector<cv::KeyPoint> kps;
        (*E)(image, cv::Mat(), kps, descriptors);
        n = (int)kps.size();
        if (n <= cap) {
            if (n) memcpy(kps_out, kps.data(), (size_t)n * 28);
            for (int i = 0; i < n; i++) memcpy(desc_out + (size_t)i * 32, descriptors.ptr(i), 32);
        } else
            n = -1;
    }
    ARENA_REWIND();
    return n;
}

int ref_extractor_levels(void* e) { return ((Access*)e)->GetLevels(); }

/* which: 0 GetScaleFactors, 1 GetInverseScaleFactors, 2 GetScaleSigmaSquares, 3 GetInverseScaleSigmaSquares */
int ref_extractor_scale_table(void* e, int which, float* out)
{
    Access* E = (Access*)e;
    std::vector<float> v = which == 0 ? E->GetScaleFactors() : which == 1 ? E->GetInverseScaleFactors()
                         : which == 2 ? E->GetScaleSigmaSquares() : E->GetInverseScaleSigmaSquares();
    for (size_t i = 0; i < v.size(); i++) out[i] = v[i];
    return (int)v.size();
}

int ref_extractor_features_per_level(void* e, int32_t* out)
{
    const std::vector<int>& q = ((Access*)e)->quotas();
    for (size_t i = 0; i < q.size(); i++) out[i] = q[i];
    return (int)q.size();
}

int ref_extractor_umax(void* e, int32_t* out)
{
    const std::vector<int>& u = ((Access*)e)->um();
    for (size_t i = 0; i < u.size(); i++) out[i] = u[i];
    return (int)u.size();
}

/* mvImagePyramid[level] of the last call (the public member Frame::ComputeStereoMatches reads) */
int ref_extractor_level_size(void* e, int level, int* w, int* h)
{
    Access* E = (Access*)e;
    if (level < 0 || level >= (int)E->mvImagePyramid.size()) return -1;
    *w = E->mvImagePyramid[level].cols;
    *h = E->mvImagePyramid[level].rows;
    return 0;
}

/* border > 0 also returns the copyMakeBorder margin around the level (at most EDGE_THRESHOLD = 19) */
int ref_extractor_level_image(void* e, int level, int border, uint8_t* dst, size_t dstep)
{
    Access* E = (Access*)e;
    if (level < 0 || level >= (int)E->mvImagePyramid.size() || border < 0 || border > 19) return -1;
    const cv::Mat& m = E->mvImagePyramid[level];
    for (int y = -border; y < m.rows + border; y++)
        memcpy(dst + (size_t)(y + border) * dstep, m.data + (ptrdiff_t)y * (ptrdiff_t)(size_t)m.step - border, (size_t)m.cols + 2 * border);
    return 0;
}

/* ORBextractor::DistributeOctTree on (x, y, response) candidates in region coordinates; returns the count */
int ref_distribute_octree(void* e, const int32_t* xyr, int n, int minX, int maxX, int minY, int maxY, int N, int32_t* out_xyr, int cap)
{
    Access* E = (Access*)e;
    int m;
    ARENA_MARK();
    {
        std::vector<cv::KeyPoint> in;
        in.reserve(n);
        for (int i = 0; i < n; i++) in.push_back(cv::KeyPoint((float)xyr[3 * i], (float)xyr[3 * i + 1], 7.f, -1.f, (float)xyr[3 * i + 2]));
        std::vector<cv::KeyPoint> out = E->octree(in, minX, maxX, minY, maxY, N);
        m = (int)out.size();
        for (int i = 0; i < m && i < cap; i++) {
            out_xyr[3 * i] = (int)out[i].pt.x; out_xyr[3 * i + 1] = (int)out[i].pt.y; out_xyr[3 * i + 2] = (int)out[i].response;
        }
    }
    ARENA_REWIND();
    return m;
}

}  // extern "C"
