// Birdview front-end of the reference (src/Frame.cc:328-342):
//     cv::ORB::create(2000)->detect(img, kps, mask);  cv::cornerSubPix(img, pts, (5,5), (-1,-1), (EPS+ITER, 40, 1e-3));
//     ->compute(img, kps, desc)
// as CUDA kernels.  cv::ORB / cv::cornerSubPix are OpenCV code (not under the reference tree); the arithmetic followed
// here is OpenCV 4.x's own C++ (features2d/src/orb.cpp, keypoint.cpp, imgproc resize.cpp [INTER_LINEAR_EXACT],
// cornersubpix.cpp, samplers.cpp, filter sepFilter2D), each step restated in the test suite and checked bit-exactly
// against cv2 4.13 there.
//
// Device layout: every pyramid level of every image is its own plane with a 32-pixel reflect-101 margin (what the packed
// cv::ORB buffer provides around each layer); the mask pyramid uses the same geometry with a zero margin.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <map>
#include <tuple>

#include "ctx.cuh"
#include "device_math.cuh"

namespace orbb200 {

namespace {

constexpr int BV_LEVELS = 8;
constexpr int BV_MARGIN = 32;            // max(edgeThreshold 31, ceil(15*sqrt2), 9/2) + 1   (orb.cpp)
constexpr int BV_EDGE = 31;              // edgeThreshold
constexpr int BV_FAST_TH = 20;
constexpr int BV_SORT_CAP = 16384;       // FAST corners per level that the selection kernel can hold
constexpr int BV_MAX_WIN = 7;            // cornerSubPix half window limit
constexpr int BV_TILE = 30;              // emit tile of the whole-level FAST

struct BirdLevel {
    int w, h, pitch;
    unsigned off;            // byte offset of pixel (0,0) inside one image's plane block
    float scale, invScale;
    int quota;
    unsigned candOff;
    int candCap;
    int tabX, tabY;          // offsets into the resize tables
    int kpOff, kpCap;        // per-level output slots
};

struct BirdGeom {
    int w, h, nfeatures;
    unsigned planeBytes;     // bytes per image of one pyramid (image / mask / blurred use the same geometry)
    unsigned candPerImg;
    int kpPerImg;
    BirdLevel lv[BV_LEVELS];
};

struct BirdPlan {
    BirdGeom g{};
    int2* d_tab = nullptr;           // INTER_LINEAR_EXACT tables: {source index, c1} (c0 = 256 - c1)
    int4* d_cells = nullptr;
    int nCells = 0;
    FastSmem need;
    int batch = 0;
    // per-batch pools
    uint8_t* d_pyr = nullptr; uint8_t* d_mask = nullptr; uint8_t* d_blur = nullptr;
    uint8_t* d_maskShared = nullptr;  // one mask pyramid used for every image (orbb200_bird_set_mask), or null
    uint32_t* d_cand = nullptr; int32_t* d_candCount = nullptr;
    float4* d_lvlKp = nullptr; int32_t* d_lvlCount = nullptr;          // {x, y, harris, -} per level slot
    orbb200_kp_t* d_kps = nullptr; orbb200_kp_t* d_kps2 = nullptr; uint8_t* d_desc = nullptr;
    int32_t* d_counts = nullptr; int32_t* d_counts2 = nullptr;
    float* d_pts = nullptr;
    int* d_slow = nullptr;           // corners bird_subpix5_kernel hands to the generic kernel ([batch][kpPerImg] slots)
    int* d_list[2] = {nullptr, nullptr};   // corners continuing in the next phase (ping-pong)
    float* d_pts0 = nullptr;         // their start points
    int* d_iters = nullptr;          // and iteration counts
    // frame-to-frame birdview matching inside the batched frame step: query arrays of "frame i-1" for job i, slot 0 = the
    // last frame of the previous step (carry)
    float* d_qx = nullptr; float* d_qy = nullptr; float* d_qangle = nullptr; int32_t* d_qlevel = nullptr; uint8_t* d_qvalid = nullptr;   // [batch+1][kpPerImg]
    orbb200_kp_t* d_carryKps = nullptr; uint8_t* d_carryDesc = nullptr; int32_t* d_carryCount = nullptr;
    bool carryValid = false;
    std::vector<uint8_t> hostMask;   // packed copy of the mask whose pyramid image 0's slot of d_mask holds (orbb200_bird_extract of one image)
    bool maskCacheValid = false;
};

struct BirdState {
    std::map<std::tuple<int, int, int>, BirdPlan*> plans;
    float* d_winMask = nullptr;      // cornerSubPix window weights of the last (win_w, win_h)
    double* d_winMaskD = nullptr;    // the same values widened to double (what the reference multiplies with)
    int* d_work = nullptr;           // work counter of the generic persistent cornerSubPix kernel
    int* d_work5 = nullptr;          // counters of the phased 5x5 form: [phases] heads, [phases] list lengths, hand-over count, generic head
    size_t workInts = 0;
    int winW = -1, winH = -1;
};

inline int cvRoundF(float v) { return (int)lrintf(v); }
inline int cvFloorD(double v) { int i = (int)v; return i - (i > v); }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------
// cv::resize(prev, cur, INTER_LINEAR_EXACT) (resize.cpp bit-exact path): horizontal pass in 8.8, vertical pass rounded
// from 16.16.  isMask: followed by threshold(254, THRESH_TOZERO) (orb.cpp).
__global__ void bird_resize_kernel(uint8_t* __restrict__ pyr, unsigned planeBytes, BirdLevel S, BirdLevel D, const int2* __restrict__ tab,
                                   int isMask)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, img = blockIdx.z;
    if (x >= D.w) return;
    const int2 tx = __ldg(tab + D.tabX + x), ty = __ldg(tab + D.tabY + y);
    const uint8_t* s = pyr + (size_t)img * planeBytes + S.off;
    const int x0 = tx.x, x1 = min(x0 + 1, S.w - 1), y0 = ty.x, y1 = min(y0 + 1, S.h - 1);
    const int a1 = tx.y, a0 = 256 - a1, b1 = ty.y, b0 = 256 - b1;
    const uint32_t h0 = (uint32_t)(s[(size_t)y0 * S.pitch + x0] * a0 + s[(size_t)y0 * S.pitch + x1] * a1) & 0xffffu;
    const uint32_t h1 = (uint32_t)(s[(size_t)y1 * S.pitch + x0] * a0 + s[(size_t)y1 * S.pitch + x1] * a1) & 0xffffu;
    uint32_t v = min((h0 * b0 + h1 * b1 + 32768u) >> 16, 255u);
    if (isMask && v <= 254u) v = 0;
    pyr[(size_t)img * planeBytes + D.off + (size_t)y * D.pitch + x] = (uint8_t)v;
}

// copyMakeBorder(BORDER_REFLECT_101) of BV_MARGIN pixels around every level
__global__ void bird_border_kernel(uint8_t* __restrict__ pyr, unsigned planeBytes, BirdGeom g)
{
    const int level = blockIdx.y, img = blockIdx.z;
    const BirdLevel L = g.lv[level];
    const int W = L.w + 2 * BV_MARGIN, H = L.h + 2 * BV_MARGIN;
    uint8_t* B = pyr + (size_t)img * planeBytes + L.off;
    // only the margin is visited: BV_MARGIN full rows above and below, then 2 * BV_MARGIN pixels beside every image row
    const int bandPx = BV_MARGIN * W, total = 2 * bandPx + L.h * 2 * BV_MARGIN;
    (void)H;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int x, y;
        if (i < 2 * bandPx) {
            const int band = i >= bandPx, j = i - band * bandPx;
            const int yy = j / W;
            x = j - yy * W - BV_MARGIN;
            y = band ? L.h + yy : yy - BV_MARGIN;
        } else {
            const int j = i - 2 * bandPx;
            const int cc = j & (2 * BV_MARGIN - 1);
            y = j / (2 * BV_MARGIN);
            x = cc < BV_MARGIN ? cc - BV_MARGIN : L.w + cc - BV_MARGIN;
        }
        int sx = x < 0 ? -x : (x >= L.w ? 2 * (L.w - 1) - x : x);
        int sy = y < 0 ? -y : (y >= L.h ? 2 * (L.h - 1) - y : y);
        sx = min(max(sx, 0), L.w - 1); sy = min(max(sy, 0), L.h - 1);      // (levels narrower than the margin)
        B[(ptrdiff_t)y * L.pitch + x] = B[(ptrdiff_t)sy * L.pitch + sx];
    }
}

// The whole pyramid of one image by one CTA (batched path): levels 1..7 in turn (each from the previous one, the reference's
// order), then the reflect-101 margins of all levels.  Seven dependent resize launches + the border launch over a batch were
// 0.35 ms per 128 images, mostly launch-to-launch latency of many small CTAs; here an image's 0.5 MP stay in one SM's L1/L2
// and 128 images fill 128 SMs.  Same arithmetic as bird_resize_kernel / bird_border_kernel.
__global__ void __launch_bounds__(1024) bird_pyramid_kernel(uint8_t* __restrict__ pyr, unsigned planeBytes, BirdGeom g, const int2* __restrict__ tab,
                                                            int nLevels)
{
    uint8_t* P = pyr + (size_t)blockIdx.x * planeBytes;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int l = 1; l < nLevels; l++) {
        const BirdLevel S = g.lv[l - 1], D = g.lv[l];
        const uint8_t* s = P + S.off;
        uint8_t* d = P + D.off;
        for (int y = wid; y < D.h; y += 32) {
            const int2 ty = tab[D.tabY + y];
            const int y0 = ty.x, y1 = min(y0 + 1, S.h - 1), b1 = ty.y, b0 = 256 - b1;
            const uint8_t* r0 = s + (size_t)y0 * S.pitch;
            const uint8_t* r1 = s + (size_t)y1 * S.pitch;
            // four output pixels per lane and trip, every load of the four requested before the first store: source and
            // destination live in the same allocation, so the compiler keeps loads behind earlier stores, and a lane with one pixel
            // in flight waits out two dependent load latencies (table entry, then the four taps) per pixel
            for (int xb = lane; xb < D.w; xb += 128) {
                int2 tx[4];
                uint32_t p00[4], p01[4], p10[4], p11[4];
#pragma unroll
                for (int k = 0; k < 4; k++) tx[k] = tab[D.tabX + min(xb + 32 * k, D.w - 1)];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int x0 = tx[k].x, x1 = min(x0 + 1, S.w - 1);
                    p00[k] = r0[x0]; p01[k] = r0[x1]; p10[k] = r1[x0]; p11[k] = r1[x1];
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int x = xb + 32 * k, a1 = tx[k].y, a0 = 256 - a1;
                    const uint32_t h0 = (uint32_t)(p00[k] * a0 + p01[k] * a1) & 0xffffu;
                    const uint32_t h1 = (uint32_t)(p10[k] * a0 + p11[k] * a1) & 0xffffu;
                    if (x < D.w) d[(size_t)y * D.pitch + x] = (uint8_t)min((h0 * b0 + h1 * b1 + 32768u) >> 16, 255u);
                }
            }
        }
        __syncthreads();                    // level l complete (global writes of this CTA) before level l+1 reads it
    }
    for (int l = 0; l < nLevels; l++) {
        const BirdLevel L = g.lv[l];
        uint8_t* B = P + L.off;
        const int W = L.w + 2 * BV_MARGIN;
        // left / right margins of the image rows first (the top / bottom bands copy whole bordered rows afterwards)
        for (int i = threadIdx.x; i < L.h * 2 * BV_MARGIN; i += 1024) {
            const int y = i / (2 * BV_MARGIN), cc = i & (2 * BV_MARGIN - 1);
            const int x = cc < BV_MARGIN ? cc - BV_MARGIN : L.w + cc - BV_MARGIN;
            int sx = x < 0 ? -x : 2 * (L.w - 1) - x;
            sx = min(max(sx, 0), L.w - 1);
            B[(ptrdiff_t)y * L.pitch + x] = B[(ptrdiff_t)y * L.pitch + sx];
        }
        __syncthreads();
        // top / bottom bands: bordered row y <- bordered row reflect(y), byte-wise over W (rows start 32 bytes left of x = 0)
        for (int i = threadIdx.x; i < 2 * BV_MARGIN * W; i += 1024) {
            const int band = i >= BV_MARGIN * W, j = i - band * BV_MARGIN * W;
            const int yy = j / W, x = j - yy * W - BV_MARGIN;
            const int y = band ? L.h + yy : yy - BV_MARGIN;
            int sy = y < 0 ? -y : 2 * (L.h - 1) - y;
            sy = min(max(sy, 0), L.h - 1);
            B[(ptrdiff_t)y * L.pitch + x] = B[(ptrdiff_t)sy * L.pitch + x];
        }
    }
}

// ---- KeyPointsFilter::retainBest (keypoint.cpp) = libstdc++ std::nth_element + std::partition, run literally by one
//      on (response, payload) arrays in shared memory: the ORDER of the survivors is part of cv::ORB's output.
//      Sequential pieces (median of three, final insertion sort of <= 3 elements, the heap-select fallback) run on one
//      thread; the partition steps, where the time goes, run on the whole CTA (below) ----
struct RB {
    float* r; uint32_t* p;
    __device__ __forceinline__ void swp(int i, int j) { const float a = r[i]; r[i] = r[j]; r[j] = a; const uint32_t b = p[i]; p[i] = p[j]; p[j] = b; }
    __device__ __forceinline__ bool gt(int i, int j) const { return r[i] > r[j]; }
};

__device__ void rb_move_median_to_first(RB v, int result, int a, int b, int c)
{
    if (v.gt(a, b)) {
        if (v.gt(b, c)) v.swp(result, b);
        else if (v.gt(a, c)) v.swp(result, c);
        else v.swp(result, a);
    } else if (v.gt(a, c)) v.swp(result, a);
    else if (v.gt(b, c)) v.swp(result, c);
    else v.swp(result, b);
}

__device__ void rb_insertion_sort(RB v, int first, int last)
{
    if (first == last) return;
    for (int i = first + 1; i < last; i++) {
        const float vr = v.r[i]; const uint32_t vp = v.p[i];
        if (vr > v.r[first]) {
            for (int j = i; j > first; j--) { v.r[j] = v.r[j - 1]; v.p[j] = v.p[j - 1]; }
            v.r[first] = vr; v.p[first] = vp;
        } else {
            int j = i;
            while (vr > v.r[j - 1]) { v.r[j] = v.r[j - 1]; v.p[j] = v.p[j - 1]; j--; }
            v.r[j] = vr; v.p[j] = vp;
        }
    }
}

__device__ void rb_adjust_heap(RB v, int base, int holeIndex, int len, float valr, uint32_t valp)
{
    const int topIndex = holeIndex;
    int secondChild = holeIndex;
    while (secondChild < (len - 1) / 2) {
        secondChild = 2 * (secondChild + 1);
        if (v.r[base + secondChild] > v.r[base + secondChild - 1]) secondChild--;
        v.r[base + holeIndex] = v.r[base + secondChild]; v.p[base + holeIndex] = v.p[base + secondChild];
        holeIndex = secondChild;
    }
    if ((len & 1) == 0 && secondChild == (len - 2) / 2) {
        secondChild = 2 * (secondChild + 1);
        v.r[base + holeIndex] = v.r[base + secondChild - 1]; v.p[base + holeIndex] = v.p[base + secondChild - 1];
        holeIndex = secondChild - 1;
    }
    int parent = (holeIndex - 1) / 2;
    while (holeIndex > topIndex && v.r[base + parent] > valr) {
        v.r[base + holeIndex] = v.r[base + parent]; v.p[base + holeIndex] = v.p[base + parent];
        holeIndex = parent;
        parent = (holeIndex - 1) / 2;
    }
    v.r[base + holeIndex] = valr; v.p[base + holeIndex] = valp;
}

__device__ void rb_heap_select(RB v, int first, int middle, int last)
{
    const int len = middle - first;
    if (len >= 2)
        for (int parent = (len - 2) / 2;; parent--) {
            rb_adjust_heap(v, first, parent, len, v.r[first + parent], v.p[first + parent]);
            if (parent == 0) break;
        }
    for (int i = middle; i < last; i++)
        if (v.r[i] > v.r[first]) {
            const float vr = v.r[i]; const uint32_t vp = v.p[i];
            v.r[i] = v.r[first]; v.p[i] = v.p[first];
            rb_adjust_heap(v, first, 0, len, vr, vp);
        }
}

// ---- partition steps with the whole CTA: the control flow of std::nth_element stays sequential (and identical in
//      every thread), but each partition step is done in parallel.  A Hoare partition swaps the k-th "left stopper"
//      (ascending indices that do not belong left of the pivot) with the k-th "right stopper" (descending indices that
//      do not belong right of it) while the former lies before the latter; the stoppers are a property of the
//      untouched data, so two stream compactions give both lists, a count gives the number of swaps K, and the K
//      disjoint swaps and the returned cut point follow -- the exact element order of the sequential code. ----
#ifndef ORBB200_SEL_THREADS
#define ORBB200_SEL_THREADS 256
#endif
constexpr int SEL_THREADS = ORBB200_SEL_THREADS;       // batches: many (image, level) CTAs in flight
constexpr int SEL_THREADS_FEW = 1024;                  // one or two images: the level-0 CTA is the critical path of the whole frame

struct RBPar { float* r; uint32_t* p; uint16_t* A; uint16_t* B; int* sc; };   // sc: [NT / 32 + 4] ints of scratch

// exclusive block scan of one int per thread (all threads call); total returned through sc
template <int NT>
__device__ __forceinline__ int rb_block_scan(int v, int* sc, int& total)
{
    constexpr int SEL_THREADS = NT;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) sc[wid] = x;
    __syncthreads();
    if (wid == 0) {
        int t = lane < SEL_THREADS / 32 ? sc[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, t, o);
            if (lane >= o) t += y;
        }
        if (lane < SEL_THREADS / 32) sc[lane] = t;
    }
    __syncthreads();
    const int base = wid ? sc[wid - 1] : 0;
    total = sc[SEL_THREADS / 32 - 1];
    __syncthreads();
    return base + x - v;
}

// ge == false: std::__unguarded_partition(lo, hi, pivot value pv) with comp = greater  -> cut
// ge == true : std::partition(lo, hi, response >= pv)                                   -> first element of the false group
template <int NT>
__device__ int rb_partition_block(RBPar v, int lo, int hi, float pv, bool ge)
{
    constexpr int SEL_THREADS = NT;
    const int tid = threadIdx.x;
    const int len = hi - lo;
    if (len <= 0) return lo;
    const int chunk = (len + SEL_THREADS - 1) / SEL_THREADS;
    const int b = min(lo + tid * chunk, hi), e = min(b + chunk, hi);
    int cnt = 0;                                        // low 16 bits: left stoppers, high 16: right stoppers
    for (int i = b; i < e; i++) {
        const float x = v.r[i];
        const bool isL = ge ? !(x >= pv) : !(x > pv);
        const bool isR = ge ? (x >= pv) : !(pv > x);
        cnt += (int)isL + ((int)isR << 16);
    }
    int total;
    const int base = rb_block_scan<NT>(cnt, v.sc, total);
    const int nL = total & 0xffff, nR = total >> 16;
    int kL = base & 0xffff, kR = base >> 16;
    for (int i = b; i < e; i++) {
        const float x = v.r[i];
        const bool isL = ge ? !(x >= pv) : !(x > pv);
        const bool isR = ge ? (x >= pv) : !(pv > x);
        if (isL) v.A[kL++] = (uint16_t)i;
        if (isR) v.B[nR - 1 - kR++] = (uint16_t)i;      // descending
    }
    int* sK = v.sc + SEL_THREADS / 32;
    if (tid == 0) *sK = 0;
    __syncthreads();
    const int m = min(nL, nR);
    int local = 0;
    for (int k = tid; k < m; k += SEL_THREADS) local += v.A[k] < v.B[k];
    if (local) atomicAdd(sK, local);
    __syncthreads();
    const int K = *sK;
    for (int k = tid; k < K; k += SEL_THREADS) {
        const int i = v.A[k], j = v.B[k];
        const float a = v.r[i]; v.r[i] = v.r[j]; v.r[j] = a;
        const uint32_t q = v.p[i]; v.p[i] = v.p[j]; v.p[j] = q;
    }
    int cut;
    if (ge) cut = lo + nR;
    else {
        const int lim = K ? (int)v.B[K - 1] : hi;
        cut = (K < nL && (int)v.A[K] < lim) ? (int)v.A[K] : lim;
    }
    __syncthreads();
    return cut;
}

// KeyPointsFilter::retainBest, all threads of the CTA call; returns the new count
template <int NT>
__device__ int rb_retain_best_block(RBPar v, int n, int n_points)
{
    if (!(n_points >= 0 && n > n_points)) return n;
    if (n_points == 0) return 0;
    const int tid = threadIdx.x;
    RB s{v.r, v.p};
    const int nth = n_points - 1;
    int first = 0, last = n, depth = 0;
    for (int k = n; k > 1; k >>= 1) depth++;
    depth *= 2;
    bool done = false;
    while (last - first > 3) {
        if (depth == 0) {
            if (tid == 0) { rb_heap_select(s, first, nth + 1, last); s.swp(first, nth); }
            done = true;
            break;
        }
        --depth;
        const int mid = first + (last - first) / 2;
        if (tid == 0) rb_move_median_to_first(s, first, first + 1, mid, last - 1);
        __syncthreads();
        const int cut = rb_partition_block<NT>(v, first + 1, last, v.r[first], false);
        if (cut <= nth) first = cut; else last = cut;
    }
    if (!done && tid == 0) rb_insertion_sort(s, first, last);
    __syncthreads();
    return rb_partition_block<NT>(v, n_points, n, v.r[n_points - 1], true);
}

// HarrisResponses (orb.cpp), blockSize 7, k 0.04
__device__ float bird_harris(const uint8_t* img, int pitch, int x0, int y0)
{
    // the 7x7 block of Sobel responses reads a 9x9 pixel patch: its rows stream through three register rows (81 byte loads
    // instead of 8 per block pixel), the integer sums are the reference's
    const uint8_t* ptr0 = img + (ptrdiff_t)(y0 - 4) * pitch + (x0 - 4);
    int a = 0, b = 0, c = 0;
    int r0[9], r1[9], r2[9];
#pragma unroll
    for (int j = 0; j < 9; j++) { r0[j] = ptr0[j]; r1[j] = ptr0[pitch + j]; }
#pragma unroll
    for (int i = 0; i < 7; i++) {
        const uint8_t* pr = ptr0 + (i + 2) * pitch;
#pragma unroll
        for (int j = 0; j < 9; j++) r2[j] = pr[j];
#pragma unroll
        for (int j = 0; j < 7; j++) {
            const int Ix = (r1[j + 2] - r1[j]) * 2 + (r0[j + 2] - r0[j]) + (r2[j + 2] - r2[j]);
            const int Iy = (r2[j + 1] - r0[j + 1]) * 2 + (r2[j] - r0[j]) + (r2[j + 2] - r0[j + 2]);
            a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
        }
#pragma unroll
        for (int j = 0; j < 9; j++) { r0[j] = r1[j]; r1[j] = r2[j]; }
    }
    const float scale = __fdiv_rn(1.f, (float)(4 * 7) * 255.f);
    const float ssq = __fmul_rn(__fmul_rn(__fmul_rn(scale, scale), scale), scale);
    const float fa = (float)a, fb = (float)b, fc = (float)c;
    const float sum = __fadd_rn(fa, fb);
    const float t = __fsub_rn(__fsub_rn(__fmul_rn(fa, fb), __fmul_rn(fc, fc)), __fmul_rn(__fmul_rn(0.04f, sum), sum));
    return __fmul_rn(t, ssq);
}

// Per (image, level): FAST corners -> mask filter -> row-major order (cv::FAST's) -> retainBest(2 * quota) on the FAST
// score -> Harris responses -> retainBest(quota) (orb.cpp computeKeyPoints).  Writes the level's survivors in order.
template <int NT>
__global__ void __launch_bounds__(NT) bird_select_kernel(BirdGeom g, const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ mpyr,
                                                                  unsigned maskPlaneBytes,      // 0: one mask pyramid for every image
                                                                  const uint32_t* __restrict__ cand, const int32_t* __restrict__ candCount,
                                                                  float4* __restrict__ lvlKp, int32_t* __restrict__ lvlCount,
                                                                  int32_t* __restrict__ status,
                                                                  int capKey, int capN,     // this launch: sort keys (power of two) and corners its shared memory holds
                                                                  int prevKey, int prevN)   // the previous (smaller) launch's, 0 for the first
{
    // The kernel is a chain of block-wide steps (latency-bound), so residency matters: it is launched in tiers of growing
    // shared memory -- 24 KB (levels of <= 2048 corners, 9 CTAs per SM), 72 KB (<= 5120, three per SM; a 400x400 level 0 holds
    // 3-5 k), 196 KB (the NMS bound) -- and a CTA returns at once when its level belongs to another tier.
    constexpr int SEL_THREADS = NT;
    extern __shared__ uint32_t selSmem[];
    uint32_t* key = selSmem;                                              // [capKey]
    float* resp = reinterpret_cast<float*>(selSmem + capKey);             // [capN]
    uint16_t* stopA = reinterpret_cast<uint16_t*>(selSmem + capKey + capN);     // [capN] left stoppers of a partition step
    uint16_t* stopB = stopA + capN;                                       // [capN] right stoppers
    __shared__ int sCount, sScratch[SEL_THREADS / 32 + 4];
    const int level = blockIdx.y, img = blockIdx.x, tid = threadIdx.x;      // x = image: the long level-0 CTAs of every image are dispatched first
    const BirdLevel L = g.lv[level];
    int n = candCount[img * MAX_LEVELS + level];            // fast_cells_kernel's layout
    if (n > L.candCap || n > BV_SORT_CAP) {
        if (tid == 0 && prevKey == 0) { atomicExch(status, 3); lvlCount[img * BV_LEVELS + level] = 0; }
        return;
    }
    int P2 = 1;
    while (P2 < n) P2 <<= 1;
    if (P2 > capKey || n > capN) return;                    // a later tier's level
    if (prevKey > 0 && P2 <= prevKey && n <= prevN) return; // an earlier tier did it
    const uint32_t* C = cand + (size_t)img * g.candPerImg + L.candOff;
    const uint8_t* M = mpyr ? mpyr + (size_t)img * maskPlaneBytes + L.off : nullptr;
    // key = (y << 20) | (x << 8) | score : ascending == cv::FAST's row-major output order
    auto make_key = [&](int i) {
        const uint32_t v = C[i];
        const int x = (int)(v & 0xfff) + FAST_BORDER, y = (int)((v >> 12) & 0xfff) + FAST_BORDER;
        // KeyPointsFilter::runByPixelsMask: mask((int)(y + 0.5f), (int)(x + 0.5f)) == 0 -> dropped
        return (!M || M[(size_t)y * L.pitch + x] != 0) ? (((uint32_t)y << 20) | ((uint32_t)x << 8) | (v >> 24)) : 0xffffffffu;
    };
    const int nRows = L.h + 1;
    if (nRows <= capN) {
        // Row-major order without a comparison sort: the keys are unique in (y, x), so a counting sort by row (histogram, scan,
        // scatter through per-row cursors) followed by a rank inside each row (a dozen corners, one warp per row) gives the same
        // permutation as sorting the keys -- in ~6 block-wide steps instead of the ~90 of a bitonic network.  The row cursors alias
        // the stopper lists (not in use before the partitions), the scattered keys the response array.
        int* rowPos = reinterpret_cast<int*>(stopA);
        uint32_t* tmp = reinterpret_cast<uint32_t*>(resp);
        for (int i = tid; i < nRows; i += SEL_THREADS) rowPos[i] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += SEL_THREADS) {
            const uint32_t k = make_key(i);
            key[i] = k;
            if (k != 0xffffffffu) atomicAdd(&rowPos[k >> 20], 1);
        }
        __syncthreads();
        const int chunk = (nRows + SEL_THREADS - 1) / SEL_THREADS;
        const int rb = min(tid * chunk, nRows), re = min(rb + chunk, nRows);
        int sum = 0;
        for (int i = rb; i < re; i++) sum += rowPos[i];
        int total;
        int basePos = rb_block_scan<NT>(sum, sScratch, total);
        for (int i = rb; i < re; i++) { const int cnt = rowPos[i]; rowPos[i] = basePos; basePos += cnt; }
        __syncthreads();
        for (int i = tid; i < n; i += SEL_THREADS) {
            const uint32_t k = key[i];
            if (k != 0xffffffffu) tmp[atomicAdd(&rowPos[k >> 20], 1)] = k;      // afterwards rowPos[r] = end of row r = start of row r + 1
        }
        __syncthreads();
        n = total;                                           // masked-out corners are gone
        const int lane = tid & 31;
        for (int r = tid >> 5; r < nRows; r += SEL_THREADS / 32) {
            const int s0 = r ? rowPos[r - 1] : 0, e0 = rowPos[r];
            for (int idx = s0 + lane; idx < e0; idx += 32) {
                const uint32_t k = tmp[idx];
                int rank = 0;
                for (int j = s0; j < e0; j++) rank += tmp[j] < k;
                key[s0 + rank] = k;
            }
        }
        __syncthreads();
    } else {
        for (int i = tid; i < P2; i += SEL_THREADS) key[i] = i < n ? make_key(i) : 0xffffffffu;
        __syncthreads();
        for (int k = 2; k <= P2; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < P2; i += SEL_THREADS) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const uint32_t a = key[i], b = key[ixj];
                        if (((i & k) == 0) == (a > b)) { key[i] = b; key[ixj] = a; }
                    }
                }
                __syncthreads();
            }
        // dropped entries sorted to the end: count the survivors
        if (tid == 0) sCount = 0;
        __syncthreads();
        int local = 0;
        for (int i = tid; i < n; i += SEL_THREADS) local += key[i] != 0xffffffffu;
        if (local) atomicAdd(&sCount, local);
        __syncthreads();
        n = sCount;
    }
    for (int i = tid; i < n; i += SEL_THREADS) resp[i] = (float)(key[i] & 0xffu);
    __syncthreads();
    RBPar v{resp, key, stopA, stopB, sScratch};
    n = rb_retain_best_block<NT>(v, n, 2 * L.quota);
    const uint8_t* I = pyr + (size_t)img * g.planeBytes + L.off;
    for (int i = tid; i < n; i += SEL_THREADS) resp[i] = bird_harris(I, L.pitch, (int)((key[i] >> 8) & 0xfff), (int)(key[i] >> 20));
    __syncthreads();
    n = rb_retain_best_block<NT>(v, n, L.quota);
    if (n > L.kpCap) {
        if (tid == 0) atomicExch(status, 4);
        n = L.kpCap;
    }
    float4* out = lvlKp + (size_t)img * g.kpPerImg + L.kpOff;
    for (int i = tid; i < n; i += SEL_THREADS) out[i] = make_float4((float)((key[i] >> 8) & 0xfff), (float)(key[i] >> 20), resp[i], 0.f);
    if (tid == 0) lvlCount[img * BV_LEVELS + level] = n;
}

// ICAngles (orb.cpp) + the final keypoint record of computeKeyPoints (pt *= scale, size = patchSize * scale); one warp per
// keypoint, output order = level-major, retainBest order inside a level.
__global__ void __launch_bounds__(256) bird_finish_kernel(BirdGeom g, const uint8_t* __restrict__ pyr, const float4* __restrict__ lvlKp,
                                                          const int32_t* __restrict__ lvlCount, orbb200_kp_t* __restrict__ kps,
                                                          int32_t* __restrict__ counts, float* __restrict__ pts)      // pts != nullptr: also the (x, y) list cornerSubPix refines
{
    const int img = blockIdx.y, lane = threadIdx.x & 31;
    const int gk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int level = -1, off = 0, total = 0;
    for (int l = 0; l < BV_LEVELS; l++) {
        const int c = lvlCount[img * BV_LEVELS + l];
        if (level < 0 && gk < total + c) { level = l; off = total; }
        total += c;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) counts[img] = min(total, g.kpPerImg);
    if (level < 0 || gk >= g.kpPerImg) return;
    const BirdLevel L = g.lv[level];
    const float4 k = lvlKp[(size_t)img * g.kpPerImg + L.kpOff + (gk - off)];
    const int x = (int)k.x, y = (int)k.y;
    const uint8_t* center = pyr + (size_t)img * g.planeBytes + L.off + (size_t)y * L.pitch + x;
    int m01 = 0, m10 = 0;
    const int u = lane - HALF_PATCH;
    if (lane < 31) {
        m10 = u * center[u];
#pragma unroll
        for (int vv = 1; vv <= HALF_PATCH; vv++)
            if (abs(u) <= c_umax[vv]) {
                const int vp = center[u + vv * L.pitch], vm = center[u - vv * L.pitch];
                m01 += vv * (vp - vm);
                m10 += u * (vp + vm);
            }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
    }
    if (lane == 0) {
        orbb200_kp_t kp;
        kp.x = __fmul_rn(k.x, L.scale); kp.y = __fmul_rn(k.y, L.scale);
        kp.size = __fmul_rn((float)PATCH_SIZE, L.scale);
        kp.angle = fast_atan2_deg((float)m01, (float)m10);
        kp.response = k.z;
        kp.octave = level;
        kp.class_id = -1;
        kps[(size_t)img * g.kpPerImg + gk] = kp;
        if (pts) { pts[2 * ((size_t)img * g.kpPerImg + gk)] = kp.x; pts[2 * ((size_t)img * g.kpPerImg + gk) + 1] = kp.y; }
    }
}

// ---- cv::getRectSubPix CV_8U -> CV_32F (samplers.cpp: getRectSubPix_8u32f; windows leaving the image go through
//      getRectSubPix_Cn_ + adjustRect = replicated border), one warp per window: every output sample only depends on
//      its own taps (the "prev" carry of the 8u32f path is (float)(t[j-1] * s), a function of the previous column's
//      taps), so the lanes fill the window independently with the reference's operation order per sample ----
__device__ void bird_get_rect_sub_pix_warp(const uint8_t* src, int src_step, int src_w, int src_h, float* dst, int win_w, int win_h,
                                           float cx, float cy, int lane)
{
    const float centerx = __fsub_rn(cx, __fmul_rn((float)(win_w - 1), 0.5f));
    const float centery = __fsub_rn(cy, __fmul_rn((float)(win_h - 1), 0.5f));
    const int ipx = (int)floorf(centerx), ipy = (int)floorf(centery);
    const int total = win_w * win_h;
    if (0 <= ipx && ipx + win_w < src_w && 0 <= ipy && ipy + win_h < src_h) {
        float a = __fsub_rn(centerx, (float)ipx);
        const float b = __fsub_rn(centery, (float)ipy);
        a = fmaxf(a, 0.0001f);
        const float b1 = __fsub_rn(1.f, b), b2 = b;
        const float a12 = __fmul_rn(a, b1), a22 = __fmul_rn(a, b);
        const float oma = __fsub_rn(1.f, a);
        const double s = __ddiv_rn(__dsub_rn(1.0, (double)a), (double)a);
        const uint8_t* p0 = src + (ptrdiff_t)ipy * src_step + ipx;
        for (int e = lane; e < total; e += 32) {
            const int i = e / win_w, j = e - i * win_w;
            const uint8_t* p = p0 + (ptrdiff_t)i * src_step;
            const float t = __fadd_rn(__fmul_rn(a12, (float)p[j + 1]), __fmul_rn(a22, (float)p[j + 1 + src_step]));
            float prev;
            if (j == 0) prev = __fmul_rn(oma, __fadd_rn(__fmul_rn(b1, (float)p[0]), __fmul_rn(b2, (float)p[src_step])));
            else {
                const float tp = __fadd_rn(__fmul_rn(a12, (float)p[j]), __fmul_rn(a22, (float)p[j + src_step]));
                prev = (float)__dmul_rn((double)tp, s);
            }
            dst[e] = __fadd_rn(prev, t);
        }
        return;
    }
    const float a = __fsub_rn(centerx, (float)ipx), b = __fsub_rn(centery, (float)ipy);
    const float oma = __fsub_rn(1.f, a), omb = __fsub_rn(1.f, b);
    const float a11 = __fmul_rn(oma, omb), a12 = __fmul_rn(a, omb), a21 = __fmul_rn(oma, b), a22 = __fmul_rn(a, b);
    const float b1 = omb, b2 = b;
    auto tap4 = [&](const uint8_t* r0, const uint8_t* r1, int j) {
        return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn((float)r0[j], a11), __fmul_rn((float)r0[j + 1], a12)), __fmul_rn((float)r1[j], a21)),
                         __fmul_rn((float)r1[j + 1], a22));
    };
    if (0 <= ipx && ipx < src_w - win_w && 0 <= ipy && ipy < src_h - win_h) {
        const uint8_t* p0 = src + (ptrdiff_t)ipy * src_step + ipx;
        for (int e = lane; e < total; e += 32) {
            const int i = e / win_w, j = e - i * win_w;
            const uint8_t* p = p0 + (ptrdiff_t)i * src_step;
            dst[e] = tap4(p, p + src_step, j);
        }
        return;
    }
    // adjustRect: the source pointer of window row i advances only while i < rh, and only from row ry on
    int rx, ry, rw, rh;
    const uint8_t* base = src;
    if (ipx >= 0) { base += ipx; rx = 0; } else { rx = -ipx; if (rx > win_w) rx = win_w; }
    if (ipx < src_w - win_w) rw = win_w;
    else { rw = src_w - ipx - 1; if (rw < 0) { base += rw; rw = 0; } }
    if (ipy >= 0) { base += (ptrdiff_t)ipy * src_step; ry = 0; } else ry = -ipy;
    if (ipy < src_h - win_h) rh = win_h;
    else { rh = src_h - ipy - 1; if (rh < 0) { base += (ptrdiff_t)rh * src_step; rh = 0; } }
    base -= rx;
    for (int e = lane; e < total; e += 32) {
        const int i = e / win_w, j = e - i * win_w;
        // rows advanced before window row i: one per earlier row k with ry <= k < rh
        const int adv = max(min(i, rh) - min(ry, min(i, rh)), 0);
        const uint8_t* p = base + (ptrdiff_t)adv * src_step;
        const uint8_t* p2 = (i < ry || i >= rh) ? p : p + src_step;
        float v;
        if (j < rx) v = __fadd_rn(__fmul_rn((float)p[rx], b1), __fmul_rn((float)p2[rx], b2));
        else if (j >= rw) v = __fadd_rn(__fmul_rn((float)p[rw], b1), __fmul_rn((float)p2[rw], b2));
        else v = tap4(p, p2, j);
        dst[e] = v;
    }
}

// cv::cornerSubPix (cornersubpix.cpp), zeroZone (-1,-1): one warp per corner.  The window samples and the per-pixel
// gradient products are computed by all lanes; the five sums (a, b, c, bb1, bb2) are accumulated in double in the
// reference's element order -- floating-point addition is not associative -- by one lane each, without FMA contraction.
constexpr int SP_WARPS = 4;
constexpr int SP_BUF = (2 * BV_MAX_WIN + 3) * (2 * BV_MAX_WIN + 3);
constexpr int SP_WIN = (2 * BV_MAX_WIN + 1) * (2 * BV_MAX_WIN + 1);

__global__ void __launch_bounds__(SP_WARPS * 32) bird_subpix_kernel(const uint8_t* __restrict__ imgs, size_t imgStrideBytes, int pitch, int cols,
                                                                    int rows, float* __restrict__ pts, size_t ptsPerImg,
                                                                    const int32_t* __restrict__ counts, int nFixed, const float* __restrict__ winMask,
                                                                    int winW, int winH, int maxIters, double eps)
{
    __shared__ float sBuf[SP_WARPS][SP_BUF];
    __shared__ double sG[SP_WARPS][5][SP_WIN];          // gxx, gxy, gyy, (gxx*px + gxy*py), (gxy*px + gyy*py)
    const int img = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int i = blockIdx.x * SP_WARPS + wid;
    const int n = counts ? counts[img] : nFixed;
    if (i >= n) return;
    const uint8_t* src = imgs + (size_t)img * imgStrideBytes;
    float* P = pts + (size_t)img * ptsPerImg * 2 + 2 * (size_t)i;
    const int win_w = winW * 2 + 1, win_h = winH * 2 + 1, bw = win_w + 2, nwin = win_w * win_h;
    float* buf = sBuf[wid];
    const float cTx = P[0], cTy = P[1];
    float cIx = cTx, cIy = cTy;
    int iter = 0;
    double err = 0;
    do {
        bird_get_rect_sub_pix_warp(src, pitch, cols, rows, buf, win_w + 2, win_h + 2, cIx, cIy, lane);
        __syncwarp();
        for (int k = lane; k < nwin; k += 32) {
            const int ii = k / win_w, j = k - ii * win_w;
            const float* subpix = buf + (ii + 1) * bw + 1;
            const double py = (double)(ii - winH), px = (double)(j - winW);
            const double m = (double)winMask[k];
            const double tgx = (double)__fsub_rn(subpix[j + 1], subpix[j - 1]);
            const double tgy = (double)__fsub_rn(subpix[j + bw], subpix[j - bw]);
            const double gxx = __dmul_rn(__dmul_rn(tgx, tgx), m);
            const double gxy = __dmul_rn(__dmul_rn(tgx, tgy), m);
            const double gyy = __dmul_rn(__dmul_rn(tgy, tgy), m);
            sG[wid][0][k] = gxx; sG[wid][1][k] = gxy; sG[wid][2][k] = gyy;
            sG[wid][3][k] = __dadd_rn(__dmul_rn(gxx, px), __dmul_rn(gxy, py));
            sG[wid][4][k] = __dadd_rn(__dmul_rn(gxy, px), __dmul_rn(gyy, py));
        }
        __syncwarp();
        double acc = 0;
        if (lane < 5) {
            const double* g = sG[wid][lane];
            for (int k = 0; k < nwin; k++) acc = __dadd_rn(acc, g[k]);
        }
        const double a = __shfl_sync(0xffffffffu, acc, 0), b = __shfl_sync(0xffffffffu, acc, 1), c = __shfl_sync(0xffffffffu, acc, 2);
        const double bb1 = __shfl_sync(0xffffffffu, acc, 3), bb2 = __shfl_sync(0xffffffffu, acc, 4);
        const double det = __dsub_rn(__dmul_rn(a, c), __dmul_rn(b, b));
        if (fabs(det) <= DBL_EPSILON * DBL_EPSILON) break;
        const double scale = __ddiv_rn(1.0, det);
        const float nx = (float)__dsub_rn(__dadd_rn((double)cIx, __dmul_rn(__dmul_rn(c, scale), bb1)), __dmul_rn(__dmul_rn(b, scale), bb2));
        const float ny = (float)__dadd_rn(__dsub_rn((double)cIy, __dmul_rn(__dmul_rn(b, scale), bb1)), __dmul_rn(__dmul_rn(a, scale), bb2));
        const float dx = __fsub_rn(nx, cIx), dy = __fsub_rn(ny, cIy);
        err = (double)__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        cIx = nx; cIy = ny;
        if (cIx < 0 || cIx >= cols || cIy < 0 || cIy >= rows) break;
    } while (++iter < maxIters && err > eps);
    if (fabsf(__fsub_rn(cIx, cTx)) > winW || fabsf(__fsub_rn(cIy, cTy)) > winH) { cIx = cTx; cIy = cTy; }
    if (lane == 0) { P[0] = cIx; P[1] = cIy; }
}

// The same for the reference's 5x5 half-window (Frame.cc:337), every trip count a constant: the six rounds of window samples and the
// four rounds of gradient products are unrolled, so a round's loads and conversion chains overlap the others' instead of queueing
// behind them.  One or two images leave most warp slots empty, the kernel's duration is (iterations of the slowest corner) x
// (latency of one iteration), and that latency is what this form cuts; a window that touches the image border takes the generic
// sampler, with the same arithmetic.
__global__ void __launch_bounds__(SP_WARPS * 32) bird_subpix_warp5_kernel(const uint8_t* __restrict__ imgs, size_t imgStrideBytes, int pitch, int cols,
                                                                          int rows, float* __restrict__ pts, size_t ptsPerImg,
                                                                          const int32_t* __restrict__ counts, int nFixed, const double* __restrict__ winMaskD,
                                                                          int maxIters, double eps)
{
    constexpr int WIN = 5, WW = 2 * WIN + 1, BW = WW + 2, NWIN = WW * WW, NSAMP = BW * BW;
    __shared__ float sBuf[SP_WARPS][NSAMP + 7];
    __shared__ double sG[SP_WARPS][5][NWIN];            // gxx, gxy, gyy, (gxx*px + gxy*py), (gxy*px + gyy*py)
    __shared__ double sMask[NWIN];
    const int img = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x < NWIN) sMask[threadIdx.x] = winMaskD[threadIdx.x];
    __syncthreads();
    const int i = blockIdx.x * SP_WARPS + wid;
    const int n = counts ? counts[img] : nFixed;
    if (i >= n) return;
    const uint8_t* src = imgs + (size_t)img * imgStrideBytes;
    float* P = pts + (size_t)img * ptsPerImg * 2 + 2 * (size_t)i;
    float* buf = sBuf[wid];
    const float cTx = P[0], cTy = P[1];
    float cIx = cTx, cIy = cTy;
    int iter = 0;
    double err = 0;
    do {
        {   // getRectSubPix_8u32f, window (BW, BW) around (cIx, cIy)
            const float centerx = __fsub_rn(cIx, (float)(BW - 1) * 0.5f), centery = __fsub_rn(cIy, (float)(BW - 1) * 0.5f);
            const int ipx = (int)floorf(centerx), ipy = (int)floorf(centery);
            if (0 <= ipx && ipx + BW < cols && 0 <= ipy && ipy + BW < rows) {
                float a = __fsub_rn(centerx, (float)ipx);
                const float b = __fsub_rn(centery, (float)ipy);
                a = fmaxf(a, 0.0001f);
                const float b1 = __fsub_rn(1.f, b), b2 = b;
                const float a12 = __fmul_rn(a, b1), a22 = __fmul_rn(a, b);
                const float oma = __fsub_rn(1.f, a);
                const double s = __ddiv_rn(__dsub_rn(1.0, (double)a), (double)a);
                const uint8_t* p0 = src + (ptrdiff_t)ipy * pitch + ipx;
#pragma unroll
                for (int r = 0; r < (NSAMP + 31) / 32; r++) {
                    const int e = min(r * 32 + lane, NSAMP - 1);               // (surplus lanes of the last round redo the last sample)
                    const int ii = e / BW, j = e - ii * BW;
                    const uint8_t* p = p0 + ii * pitch + j;
                    const float q0 = (float)p[0], q1 = (float)p[1], r0 = (float)p[pitch], r1 = (float)p[pitch + 1];
                    const float t = __fadd_rn(__fmul_rn(a12, q1), __fmul_rn(a22, r1));
                    const float tp = __fadd_rn(__fmul_rn(a12, q0), __fmul_rn(a22, r0));
                    const float prev = j == 0 ? __fmul_rn(oma, __fadd_rn(__fmul_rn(b1, q0), __fmul_rn(b2, r0))) : (float)__dmul_rn((double)tp, s);
                    buf[e] = __fadd_rn(prev, t);
                }
            } else {
                bird_get_rect_sub_pix_warp(src, pitch, cols, rows, buf, BW, BW, cIx, cIy, lane);
            }
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < (NWIN + 31) / 32; r++) {
            const int k = min(r * 32 + lane, NWIN - 1);
            const int ii = k / WW, j = k - ii * WW;
            const float* subpix = buf + (ii + 1) * BW + 1;
            const double py = (double)(ii - WIN), px = (double)(j - WIN);
            const double m = sMask[k];
            const double tgx = (double)__fsub_rn(subpix[j + 1], subpix[j - 1]);
            const double tgy = (double)__fsub_rn(subpix[j + BW], subpix[j - BW]);
            const double gxx = __dmul_rn(__dmul_rn(tgx, tgx), m);
            const double gxy = __dmul_rn(__dmul_rn(tgx, tgy), m);
            const double gyy = __dmul_rn(__dmul_rn(tgy, tgy), m);
            sG[wid][0][k] = gxx; sG[wid][1][k] = gxy; sG[wid][2][k] = gyy;
            sG[wid][3][k] = __dadd_rn(__dmul_rn(gxx, px), __dmul_rn(gxy, py));
            sG[wid][4][k] = __dadd_rn(__dmul_rn(gxy, px), __dmul_rn(gyy, py));
        }
        __syncwarp();
        double acc = 0;
        if (lane < 5) {
            const double* g = sG[wid][lane];
#pragma unroll
            for (int k = 0; k < NWIN; k++) acc = __dadd_rn(acc, g[k]);
        }
        const double a = __shfl_sync(0xffffffffu, acc, 0), b = __shfl_sync(0xffffffffu, acc, 1), c = __shfl_sync(0xffffffffu, acc, 2);
        const double bb1 = __shfl_sync(0xffffffffu, acc, 3), bb2 = __shfl_sync(0xffffffffu, acc, 4);
        const double det = __dsub_rn(__dmul_rn(a, c), __dmul_rn(b, b));
        if (fabs(det) <= DBL_EPSILON * DBL_EPSILON) break;
        const double scale = __ddiv_rn(1.0, det);
        const float nx = (float)__dsub_rn(__dadd_rn((double)cIx, __dmul_rn(__dmul_rn(c, scale), bb1)), __dmul_rn(__dmul_rn(b, scale), bb2));
        const float ny = (float)__dadd_rn(__dsub_rn((double)cIy, __dmul_rn(__dmul_rn(b, scale), bb1)), __dmul_rn(__dmul_rn(a, scale), bb2));
        const float dx = __fsub_rn(nx, cIx), dy = __fsub_rn(ny, cIy);
        err = (double)__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        cIx = nx; cIy = ny;
        if (cIx < 0 || cIx >= cols || cIy < 0 || cIy >= rows) break;
    } while (++iter < maxIters && err > eps);
    if (fabsf(__fsub_rn(cIx, cTx)) > WIN || fabsf(__fsub_rn(cIy, cTy)) > WIN) { cIx = cTx; cIy = cTy; }
    if (lane == 0) { P[0] = cIx; P[1] = cIy; }
}

// ---- the same sampler run by one thread (throughput form of cornerSubPix below) ----
__device__ void bird_get_rect_sub_pix(const uint8_t* src, int src_step, int src_w, int src_h, float* dst, int win_w, int win_h,
                                      float cx, float cy)
{
    const float centerx = __fsub_rn(cx, __fmul_rn((float)(win_w - 1), 0.5f));
    const float centery = __fsub_rn(cy, __fmul_rn((float)(win_h - 1), 0.5f));
    const int ipx = (int)floorf(centerx), ipy = (int)floorf(centery);
    if (0 <= ipx && ipx + win_w < src_w && 0 <= ipy && ipy + win_h < src_h) {
        float a = __fsub_rn(centerx, (float)ipx);
        const float b = __fsub_rn(centery, (float)ipy);
        a = fmaxf(a, 0.0001f);
        const float b1 = __fsub_rn(1.f, b), b2 = b;
        const float a12 = __fmul_rn(a, b1), a22 = __fmul_rn(a, b);
        const float oma = __fsub_rn(1.f, a);
        const double s = __ddiv_rn(__dsub_rn(1.0, (double)a), (double)a);
        const uint8_t* p = src + (ptrdiff_t)ipy * src_step + ipx;
        for (int i = 0; i < win_h; i++, p += src_step, dst += win_w) {
            float prev = __fmul_rn(oma, __fadd_rn(__fmul_rn(b1, (float)p[0]), __fmul_rn(b2, (float)p[src_step])));
            for (int j = 0; j < win_w; j++) {
                const float t = __fadd_rn(__fmul_rn(a12, (float)p[j + 1]), __fmul_rn(a22, (float)p[j + 1 + src_step]));
                dst[j] = __fadd_rn(prev, t);
                prev = (float)__dmul_rn((double)t, s);
            }
        }
        return;
    }
    const float a = __fsub_rn(centerx, (float)ipx), b = __fsub_rn(centery, (float)ipy);
    const float oma = __fsub_rn(1.f, a), omb = __fsub_rn(1.f, b);
    const float a11 = __fmul_rn(oma, omb), a12 = __fmul_rn(a, omb), a21 = __fmul_rn(oma, b), a22 = __fmul_rn(a, b);
    const float b1 = omb, b2 = b;
    auto tap4 = [&](const uint8_t* r0, const uint8_t* r1, int j) {
        return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn((float)r0[j], a11), __fmul_rn((float)r0[j + 1], a12)), __fmul_rn((float)r1[j], a21)),
                         __fmul_rn((float)r1[j + 1], a22));
    };
    if (0 <= ipx && ipx < src_w - win_w && 0 <= ipy && ipy < src_h - win_h) {
        const uint8_t* p = src + (ptrdiff_t)ipy * src_step + ipx;
        for (int i = 0; i < win_h; i++, p += src_step, dst += win_w)
            for (int j = 0; j < win_w; j++) dst[j] = tap4(p, p + src_step, j);
        return;
    }
    // adjustRect
    int rx, ry, rw, rh;
    const uint8_t* p = src;
    if (ipx >= 0) { p += ipx; rx = 0; } else { rx = -ipx; if (rx > win_w) rx = win_w; }
    if (ipx < src_w - win_w) rw = win_w;
    else { rw = src_w - ipx - 1; if (rw < 0) { p += rw; rw = 0; } }
    if (ipy >= 0) { p += (ptrdiff_t)ipy * src_step; ry = 0; } else ry = -ipy;
    if (ipy < src_h - win_h) rh = win_h;
    else { rh = src_h - ipy - 1; if (rh < 0) { p += (ptrdiff_t)rh * src_step; rh = 0; } }
    p -= rx;
    for (int i = 0; i < win_h; i++, dst += win_w) {
        const uint8_t* p2 = p + src_step;
        if (i < ry || i >= rh) p2 -= src_step;
        float s0 = __fadd_rn(__fmul_rn((float)p[rx], b1), __fmul_rn((float)p2[rx], b2));
        for (int j = 0; j < rx; j++) dst[j] = s0;
        s0 = __fadd_rn(__fmul_rn((float)p[rw], b1), __fmul_rn((float)p2[rw], b2));
        for (int j = rw; j < win_w; j++) dst[j] = s0;
        for (int j = rx; j < rw; j++) dst[j] = tap4(p, p2, j);
        if (i < rh) p = p2;
    }
}

// cornerSubPix, throughput form for batches: one thread per corner, but corners need between 2 and 40 iterations
// (median ~12), so a warp that keeps its 32 corners until the slowest one converges runs at a third of its lanes.  The
// threads are persistent instead: every trip of the loop is ONE iteration of whatever corner the lane currently holds, and
// a lane whose corner has finished stores it and fetches the next corner from a global counter.  Measured, 64 images x
// 1850 corners: fixed assignment 2.9 ms, this form 2.8 ms, one warp per corner 5.6 ms, eight lanes per corner 8.1 ms.  The
// floor is the device's FP64 rate: the reference accumulates in double, ~2060 double operations per corner and iteration,
// 3.4 G per batch, and this GPU sustains ~1.2 T double operations/s (its FP64 pipe issues ~4 lanes/clk/SM).
constexpr int SP_THREADS = 64;

__global__ void __launch_bounds__(SP_THREADS) bird_subpix_thread_kernel(const uint8_t* __restrict__ imgs, size_t imgStrideBytes, int pitch, int cols,
                                                                 int rows, float* __restrict__ pts, int ptsPerImg, int nImages,
                                                                 const int32_t* __restrict__ counts, int nFixed, const float* __restrict__ winMask,
                                                                 int winW, int winH, int maxIters, double eps, int* __restrict__ nextWork,
                                                                 const int* __restrict__ workList, const int* __restrict__ workListCount)
{
    // workList != nullptr: only the listed (image, corner) slots (what bird_subpix5_kernel handed over)
    const int win_w = winW * 2 + 1, win_h = winH * 2 + 1, bw = win_w + 2;
    const int totalSlots = workList ? *workListCount : nImages * ptsPerImg;
    float buf[(2 * BV_MAX_WIN + 3) * (2 * BV_MAX_WIN + 3)];
    const uint8_t* src = nullptr;
    float* P = nullptr;
    float cTx = 0, cTy = 0, cIx = 0, cIy = 0;
    int iter = 0;
    // next valid (image, corner) slot, or false when the work is exhausted
    auto fetch = [&]() -> bool {
        while (true) {
            int w = atomicAdd(nextWork, 1);
            if (w >= totalSlots) return false;
            if (workList) w = workList[w];
            const int img = w / ptsPerImg, i = w - img * ptsPerImg;
            if (i >= (counts ? counts[img] : nFixed)) continue;
            src = imgs + (size_t)img * imgStrideBytes;
            P = pts + ((size_t)img * ptsPerImg + i) * 2;
            cTx = P[0]; cTy = P[1]; cIx = cTx; cIy = cTy; iter = 0;
            return true;
        }
    };
    bool active = fetch();
    while (__any_sync(0xffffffffu, active)) {
        if (!active) continue;
        double a = 0, b = 0, c = 0, bb1 = 0, bb2 = 0;
        bird_get_rect_sub_pix(src, pitch, cols, rows, buf, win_w + 2, win_h + 2, cIx, cIy);
        const float* subpix = buf + bw + 1;
        for (int ii = 0, k = 0; ii < win_h; ii++, subpix += bw) {
            const double py = (double)(ii - winH);
            for (int j = 0; j < win_w; j++, k++) {
                const double m = (double)winMask[k];
                const double tgx = (double)__fsub_rn(subpix[j + 1], subpix[j - 1]);
                const double tgy = (double)__fsub_rn(subpix[j + bw], subpix[j - bw]);
                const double gxx = __dmul_rn(__dmul_rn(tgx, tgx), m);
                const double gxy = __dmul_rn(__dmul_rn(tgx, tgy), m);
                const double gyy = __dmul_rn(__dmul_rn(tgy, tgy), m);
                const double px = (double)(j - winW);
                a = __dadd_rn(a, gxx); b = __dadd_rn(b, gxy); c = __dadd_rn(c, gyy);
                bb1 = __dadd_rn(bb1, __dadd_rn(__dmul_rn(gxx, px), __dmul_rn(gxy, py)));
                bb2 = __dadd_rn(bb2, __dadd_rn(__dmul_rn(gxy, px), __dmul_rn(gyy, py)));
            }
        }
        bool finished = false;
        const double det = __dsub_rn(__dmul_rn(a, c), __dmul_rn(b, b));
        if (fabs(det) <= DBL_EPSILON * DBL_EPSILON) finished = true;
        else {
            const double scale = __ddiv_rn(1.0, det);
            const float nx = (float)__dsub_rn(__dadd_rn((double)cIx, __dmul_rn(__dmul_rn(c, scale), bb1)), __dmul_rn(__dmul_rn(b, scale), bb2));
            const float ny = (float)__dadd_rn(__dsub_rn((double)cIy, __dmul_rn(__dmul_rn(b, scale), bb1)), __dmul_rn(__dmul_rn(a, scale), bb2));
            const float dx = __fsub_rn(nx, cIx), dy = __fsub_rn(ny, cIy);
            const double err = (double)__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
            cIx = nx; cIy = ny;
            if (cIx < 0 || cIx >= cols || cIy < 0 || cIy >= rows) finished = true;
            else finished = !(++iter < maxIters && err > eps);
        }
        if (finished) {
            if (fabsf(__fsub_rn(cIx, cTx)) > winW || fabsf(__fsub_rn(cIy, cTy)) > winH) { cIx = cTx; cIy = cTy; }
            P[0] = cIx; P[1] = cIy;
            active = fetch();
        }
    }
}

// ---- cornerSubPix, fast form for the window the reference uses (Size(5,5): 11x11 gradient window, 13x13 samples) ----
// Measured on B200 (tools/ubench/fp64.cu): DADD / DMUL issue at ~60 lanes/clk/SM (17 T/s), a dependent DADD takes 8 cycles,
// F2F.F64.F32 is free beside them, F2F.F32.F64 runs at ~16 lanes/clk/SM.  One corner-iteration needs 121 x (10 DMUL + 7 DADD)
// + 169 DMUL = 2226 FP64-pipe instructions, i.e. ~1190 SM-cycles per warp of 32 corners: the floor of this kernel.  The
// generic kernel above ran at 21 % of the FP64 pipe (ncu: XU pipe 45 % busy with I2F.U8 conversions, 18 of 32 lanes active,
// long-scoreboard stalls on the thread-local sample window).  This form removes all three:
//   * a 24x28-byte source patch around the current point is staged in shared memory (cooperatively, aligned words); a
//     window that leaves it asks for a new patch around the new point (corners on straight edges jump many pixels per
//     iteration); only a patch that is not inside the image sends the corner to the generic kernel (`slowList`);
//   * bytes become floats with PRMT + FADD (0x4B0000bb - 2^23), not I2F;
//   * the 13x13 sample window is never stored: rows stream through registers (two source rows, three sample rows) while the
//     five double sums advance in the reference's element order;
//   * corners need between 2 and 40 iterations, so a lane keeps a corner for at most `budget` iterations per turn and then
//     queues it for the next launch ("phase"): the tail of a launch is at most `budget` iterations long, and the lanes of a
//     warp stay in step.  ceil(maxIters / budget) launches finish every corner.
#ifndef ORBB200_S5_UNROLL
#define ORBB200_S5_UNROLL 2
#endif
// rows of the 14-row sample loop per unrolled body.  Two: the top / cur source rows swap roles instead of being copied (14 moves
// per row less), 11 KB of loop body; measured per 128 images: 1 -> 2.08 ms, 2 -> 1.94, 3 -> 1.94, 7 -> 2.19, 14 -> 2.39 (the body
// outgrows the instruction cache)
constexpr int S5_UNROLL = ORBB200_S5_UNROLL;
constexpr int S5_THREADS = 128;
constexpr int S5_WIN = 5;                           // half window
#ifndef ORBB200_S5_BUDGET
#define ORBB200_S5_BUDGET 8
#endif
constexpr int S5_BUDGET = ORBB200_S5_BUDGET;
// patch geometry for a reach of R pixels around the point it was placed on: rows [fy-(R+6), +2R+14), columns from
// fx-(R+6) rounded down to a word, 2R+14 (+3 for the rounding) bytes wide.  R = 5: 7 words x 24 rows, 86 KB per CTA, two CTAs
// per SM; R = 3: 6 x 20, 62 KB, three CTAs per SM (12 warps: more latency hiding for a few more re-placements).
template <int R> struct S5Geom {
    static constexpr int PW = (2 * R + 14 + 3 + 3) / 4, PH = 2 * R + 14;
    static constexpr int STRIDE = (PW * PH) | 1;    // words per thread, odd (bank spread); >= PW*PH+1 absorbs the 5th-word over-read
    static constexpr int DXMAX = 4 * PW - 14, DYMAX = PH - 14;
    static_assert(STRIDE > PW * PH && ((DXMAX >> 2) + 4) <= PW, "patch layout");
};

__device__ __forceinline__ float s5_byte(unsigned w, int k)
{
    // exact (float)byte: 0x4B0000bb is 2^23 + bb
    const unsigned sel = k == 0 ? 0x7440u : k == 1 ? 0x7441u : k == 2 ? 0x7442u : 0x7443u;
    return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, sel)), 8388608.f);
}

struct S5Args {
    const uint8_t* imgs; size_t imgStrideBytes; int pitch, cols, rows;
    float* pts;                  // [nImages][ptsPerImg][2]: start point in, current / final point out
    float* pts0;                 // start points (cT) of corners that take more than one turn
    int* iters;                  // iterations done so far by those corners
    int ptsPerImg, nImages;
    const int32_t* counts; int nFixed;
    const double* winMaskD; int maxIters; double eps;
    const int* listIn; const int* listInCount;      // this phase's corners (nullptr: every slot, first phase)
    int* head;                                       // work counter of this launch
    int* listOut; int* listOutCount;                 // corners that used up their budget
    int* slowList; int* slowCount;                   // corners for the generic kernel
};

template <int R, int MINB>
__global__ void __launch_bounds__(S5_THREADS, MINB) bird_subpix5_kernel(const S5Args A)
{
    constexpr int S5_PW = S5Geom<R>::PW, S5_PH = S5Geom<R>::PH, S5_STRIDE = S5Geom<R>::STRIDE;
    extern __shared__ unsigned sPatch[];            // [S5_THREADS][S5_STRIDE]
    __shared__ double sMask[121];
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 121) sMask[tid] = A.winMaskD[tid];
    __syncthreads();
    unsigned* myPatch = sPatch + (size_t)tid * S5_STRIDE;
    unsigned* warpPatch = sPatch + (size_t)(tid - lane) * S5_STRIDE;
    const bool first = A.listIn == nullptr;
    const int totalWork = first ? A.nImages * A.ptsPerImg : *A.listInCount;
    const int pitch = A.pitch, cols = A.cols, rows = A.rows;
    float* P = nullptr;
    float cTx = 0, cTy = 0, cIx = 0, cIy = 0;
    int iter = 0, turn = 0, X0 = 0, Y0 = 0, slot = -1, img = 0;
    size_t srcOff = 0;                              // byte offset of the patch origin from imgs
    bool active = false, exhausted = false, stage = false;

    // new patch around (cIx, cIy); false: it would not lie inside the image
    auto place_patch = [&]() -> bool {
        const float fxf = floorf(cIx), fyf = floorf(cIy);
        // columns [fx-(R+6) rounded down to 4, +4*PW), rows [fy-(R+6), +PH) inside the image
        if (!(fxf >= (float)(R + 9) && fxf <= (float)(cols - 4 * S5_PW + R + 6) && fyf >= (float)(R + 6) && fyf <= (float)(rows - S5_PH + R + 6))) return false;
        X0 = ((int)fxf - (R + 6)) & ~3; Y0 = (int)fyf - (R + 6);
        srcOff = (size_t)img * A.imgStrideBytes + (size_t)Y0 * pitch + X0;
        return true;
    };
    auto to_generic = [&]() {                       // exact fallback: restart this corner from cT in the generic kernel
        P[0] = cTx; P[1] = cTy;
        A.slowList[atomicAdd(A.slowCount, 1)] = slot;
        active = false;
    };

    while (true) {
        // ---- refill: lanes without a corner fetch one; the warp stages the patches asked for ----
        if (!active && !exhausted) {
            while (true) {
                int w = atomicAdd(A.head, 1);
                if (w >= totalWork) { exhausted = true; break; }
                if (!first) w = A.listIn[w];
                img = w / A.ptsPerImg;
                const int i = w - img * A.ptsPerImg;
                if (first && i >= (A.counts ? A.counts[img] : A.nFixed)) continue;
                slot = w;
                P = A.pts + (size_t)w * 2;
                cIx = P[0]; cIy = P[1];
                if (first) { cTx = cIx; cTy = cIy; iter = 0; }
                else { cTx = A.pts0[2 * (size_t)w]; cTy = A.pts0[2 * (size_t)w + 1]; iter = A.iters[w]; }
                turn = 0; active = true;
                if (!place_patch()) { to_generic(); continue; }
                stage = true;
                break;
            }
        }
        unsigned todo = __ballot_sync(0xffffffffu, stage);
        stage = false;
        while (todo) {
            const int l = __ffs(todo) - 1;
            todo &= todo - 1;
            const size_t off = __shfl_sync(0xffffffffu, (unsigned long long)srcOff, l);
            const uint8_t* base = A.imgs + off;
            unsigned* dst = warpPatch + (size_t)l * S5_STRIDE;
#pragma unroll
            for (int k = 0; k < 6; k++) {
                const int idx = lane + 32 * k;
                if (idx < S5_PW * S5_PH) {
                    const int r = idx / S5_PW, wcol = idx - r * S5_PW;
                    dst[idx] = __ldg(reinterpret_cast<const unsigned*>(base + (size_t)r * pitch) + wcol);
                }
            }
        }
        __syncwarp();
        if (!__any_sync(0xffffffffu, active)) break;
        if (!active) continue;

        // ---- one iteration of cornerSubPix for this lane's corner ----
        const float centerx = __fsub_rn(cIx, 6.f), centery = __fsub_rn(cIy, 6.f);      // (13 - 1) * 0.5f
        const float fpx = floorf(centerx), fpy = floorf(centery);
        const int dx = (int)fpx - X0, dy = (int)fpy - Y0;
        if (!(dx >= 0 && dx <= S5Geom<R>::DXMAX && dy >= 0 && dy <= S5Geom<R>::DYMAX)) {      // window left the staged patch (also NaN)
            if (place_patch()) stage = true; else to_generic();
            continue;
        }
        float a = __fsub_rn(centerx, fpx);
        const float b = __fsub_rn(centery, fpy);
        a = fmaxf(a, 0.0001f);
        const float b1 = __fsub_rn(1.f, b), b2 = b;
        const float a12 = __fmul_rn(a, b1), a22 = __fmul_rn(a, b);
        const float oma = __fsub_rn(1.f, a);
        const double s = __ddiv_rn(__dsub_rn(1.0, (double)a), (double)a);
        const unsigned* rowp = myPatch + dy * S5_PW + (dx >> 2);
        const int sh = (dx & 3) * 8;

        double sa = 0, sb = 0, sc = 0, sbb1 = 0, sbb2 = 0;
        float top[14], cur[14], wm2[13], wm1[13], wn[13];
#pragma unroll
        for (int j = 0; j < 13; j++) { wm2[j] = 0.f; wm1[j] = 0.f; }
#pragma unroll
        for (int j = 0; j < 14; j++) top[j] = 0.f;
#pragma unroll S5_UNROLL
        for (int r = 0; r < 14; r++, rowp += S5_PW) {
            {   // source row r: 14 bytes starting at patch byte column dx
                const unsigned w0 = rowp[0], w1 = rowp[1], w2 = rowp[2], w3 = rowp[3], w4 = rowp[4];
                const unsigned v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh), v2 = __funnelshift_r(w2, w3, sh),
                               v3 = __funnelshift_r(w3, w4, sh);
#pragma unroll
                for (int k = 0; k < 4; k++) { cur[k] = s5_byte(v0, k); cur[4 + k] = s5_byte(v1, k); cur[8 + k] = s5_byte(v2, k); }
                cur[12] = s5_byte(v3, 0); cur[13] = s5_byte(v3, 1);
            }
            if (r >= 1) {
                // sample row i = r-1 (getRectSubPix_8u32f): dst[j] = prev + t, prev' = (float)(t * s)
                float prev = __fmul_rn(oma, __fadd_rn(__fmul_rn(b1, top[0]), __fmul_rn(b2, cur[0])));
#pragma unroll
                for (int j = 0; j < 13; j++) {
                    const float t = __fadd_rn(__fmul_rn(a12, top[j + 1]), __fmul_rn(a22, cur[j + 1]));
                    wn[j] = __fadd_rn(prev, t);
                    prev = (float)__dmul_rn((double)t, s);
                }
                if (r >= 3) {
                    // gradient row ii = r-3: samples (ii+1, j+1) of the 13x13 window, rows wm2 / wm1 / wn = ii, ii+1, ii+2
                    const int ii = r - 3;
                    const double py = (double)(ii - S5_WIN);
                    const double* mrow = sMask + ii * 11;
                    // x * k is exact in double for k in {0, +-1, +-2, +-4}; then fma(x, k, y) == x*k + y with its single
                    // rounding, so the reference's "gxx*px + gxy*py" (two products, one sum) needs one multiply less
                    // wherever px (compile time) or py (per row, warp-uniform) is such a k: 87 % of the window.
                    const int apy = abs(ii - S5_WIN);
                    const bool pyExact = apy != 3 && apy != 5;
#pragma unroll
                    for (int j = 0; j < 11; j++) {
                        const double m = mrow[j];
                        const double tgx = (double)__fsub_rn(wm1[j + 2], wm1[j]);
                        const double tgy = (double)__fsub_rn(wn[j + 1], wm2[j + 1]);
                        const double gxx = __dmul_rn(__dmul_rn(tgx, tgx), m);
                        const double gxy = __dmul_rn(__dmul_rn(tgx, tgy), m);
                        const double gyy = __dmul_rn(__dmul_rn(tgy, tgy), m);
                        const double px = (double)(j - S5_WIN);
                        constexpr int dummy = 0; (void)dummy;
                        const bool pxExact = (j - S5_WIN) != 3 && (j - S5_WIN) != -3 && (j - S5_WIN) != 5 && (j - S5_WIN) != -5;
                        sa = __dadd_rn(sa, gxx); sb = __dadd_rn(sb, gxy); sc = __dadd_rn(sc, gyy);
                        double t1, t2;
                        if (pxExact) { t1 = __fma_rn(gxx, px, __dmul_rn(gxy, py)); t2 = __fma_rn(gxy, px, __dmul_rn(gyy, py)); }
                        else if (pyExact) { t1 = __fma_rn(gxy, py, __dmul_rn(gxx, px)); t2 = __fma_rn(gyy, py, __dmul_rn(gxy, px)); }
                        else { t1 = __dadd_rn(__dmul_rn(gxx, px), __dmul_rn(gxy, py)); t2 = __dadd_rn(__dmul_rn(gxy, px), __dmul_rn(gyy, py)); }
                        sbb1 = __dadd_rn(sbb1, t1);
                        sbb2 = __dadd_rn(sbb2, t2);
                    }
                }
#pragma unroll
                for (int j = 0; j < 13; j++) { wm2[j] = wm1[j]; wm1[j] = wn[j]; }
            }
#pragma unroll
            for (int j = 0; j < 14; j++) top[j] = cur[j];
        }
        bool finished = false;
        const double det = __dsub_rn(__dmul_rn(sa, sc), __dmul_rn(sb, sb));
        if (fabs(det) <= DBL_EPSILON * DBL_EPSILON) finished = true;
        else {
            const double scale = __ddiv_rn(1.0, det);
            const float nx = (float)__dsub_rn(__dadd_rn((double)cIx, __dmul_rn(__dmul_rn(sc, scale), sbb1)), __dmul_rn(__dmul_rn(sb, scale), sbb2));
            const float ny = (float)__dadd_rn(__dsub_rn((double)cIy, __dmul_rn(__dmul_rn(sb, scale), sbb1)), __dmul_rn(__dmul_rn(sa, scale), sbb2));
            const float ex = __fsub_rn(nx, cIx), ey = __fsub_rn(ny, cIy);
            const double err = (double)__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
            cIx = nx; cIy = ny;
            if (cIx < 0 || cIx >= cols || cIy < 0 || cIy >= rows) finished = true;
            else finished = !(++iter < A.maxIters && err > A.eps);
        }
        if (finished) {
            if (fabsf(__fsub_rn(cIx, cTx)) > S5_WIN || fabsf(__fsub_rn(cIy, cTy)) > S5_WIN) { cIx = cTx; cIy = cTy; }
            P[0] = cIx; P[1] = cIy;
            active = false;
        }
        else if (++turn >= S5_BUDGET) {              // budget used up: next phase continues from here
            P[0] = cIx; P[1] = cIy;
            A.pts0[2 * (size_t)slot] = cTx; A.pts0[2 * (size_t)slot + 1] = cTy;
            A.iters[slot] = iter;
            A.listOut[atomicAdd(A.listOutCount, 1)] = slot;
            active = false;
        }
    }
}

// n device-resident images (rows of `stride` bytes, image i at imgs + i*imgBytes) into level 0 of the planes
__global__ void __launch_bounds__(128) bird_import_kernel(const uint8_t* __restrict__ imgs, size_t imgBytes, size_t stride, uint8_t* __restrict__ pyr,
                                                          unsigned planeBytes, BirdLevel L0)
{
    const int y = blockIdx.x, img = blockIdx.y;
    const uint8_t* s = imgs + (size_t)img * imgBytes + (size_t)y * stride;
    uint8_t* d = pyr + (size_t)img * planeBytes + L0.off + (size_t)y * L0.pitch;      // 32-byte aligned (margin 32, pitch % 128 == 0)
    if (((uintptr_t)s & 3) == 0) {
        const int nw = L0.w >> 2;
        for (int x = threadIdx.x; x < nw; x += 128) reinterpret_cast<uint32_t*>(d)[x] = __ldg(reinterpret_cast<const uint32_t*>(s) + x);
        for (int x = (nw << 2) + threadIdx.x; x < L0.w; x += 128) d[x] = s[x];
    } else
        for (int x = threadIdx.x; x < L0.w; x += 128) d[x] = s[x];
}

// Query arrays of BirdviewMatch(Last, Cur) (src/ORBmatcher.cc:1788-1899) for every frame of a step: slot 0 = the frame
// carried over from the previous step, slot s = frame s-1 of this step.  Query i1 is keypoint i1 of the "last" frame:
// window centre = its position, level = its octave, angle for the rotation histogram.
__global__ void __launch_bounds__(256) bird_queries_kernel(const orbb200_kp_t* __restrict__ kps, const int32_t* __restrict__ counts,
                                                           const orbb200_kp_t* __restrict__ carryKps, const int32_t* __restrict__ carryCount,
                                                           int kpPerImg, float* __restrict__ qx, float* __restrict__ qy, float* __restrict__ qangle,
                                                           int32_t* __restrict__ qlevel, uint8_t* __restrict__ qvalid)
{
    const int slot = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kpPerImg) return;
    const orbb200_kp_t* K = slot == 0 ? carryKps : kps + (size_t)(slot - 1) * kpPerImg;
    const int n = min(slot == 0 ? *carryCount : counts[slot - 1], kpPerImg);
    const size_t o = (size_t)slot * kpPerImg + i;
    const bool ok = i < n;
    orbb200_kp_t k{};
    if (ok) k = K[i];
    qx[o] = k.x; qy[o] = k.y; qangle[o] = k.angle; qlevel[o] = k.octave; qvalid[o] = ok ? 1 : 0;
}

// ORB::compute on provided keypoints: KeyPointsFilter::runByImageBorder(kps, image size, 31) keeping the order, then a
// stable regrouping by octave when the input is not sorted by level (orb.cpp detectAndCompute).  One CTA per image.
// pts != nullptr: the keypoints' positions are taken from the refined (x, y) list (the fused pipelines skip the copy back into `in`).
// NT: 256 in batches (one CTA per image, many images), 1024 for one or two images (the lone CTA's chunk loops are the latency).
template <int NT>
__global__ void __launch_bounds__(NT) bird_filter_kernel(BirdGeom g, const orbb200_kp_t* __restrict__ in, const int32_t* __restrict__ inCount,
                                                         orbb200_kp_t* __restrict__ out, int32_t* __restrict__ outCount, const float* __restrict__ pts)
{
    __shared__ int sBase[BV_LEVELS + 1], sWarp[NT / 32], sSorted;
    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = min(inCount[img], g.kpPerImg);
    const orbb200_kp_t* K = in + (size_t)img * g.kpPerImg;
    const float* PT = pts ? pts + 2 * (size_t)img * g.kpPerImg : nullptr;
    auto load = [&](int i) { orbb200_kp_t k = K[i]; if (PT) { k.x = PT[2 * i]; k.y = PT[2 * i + 1]; } return k; };
    orbb200_kp_t* O = out + (size_t)img * g.kpPerImg;
    auto keep = [&](const orbb200_kp_t& k) {
        if (g.h <= BV_EDGE * 2 || g.w <= BV_EDGE * 2) return false;
        const int x = __float2int_rn(k.x), y = __float2int_rn(k.y);
        return x >= BV_EDGE && x < g.w - BV_EDGE && y >= BV_EDGE && y < g.h - BV_EDGE && k.octave >= 0 && k.octave < BV_LEVELS;
    };
    // sortedByLevel?
    if (tid == 0) sSorted = 1;
    __syncthreads();
    for (int i = tid + 1; i < n; i += NT) if (K[i].octave < K[i - 1].octave) sSorted = 0;
    __syncthreads();
    const bool sorted = sSorted != 0;
    // pass 1: per-level counts of kept keypoints (one level "all" when already sorted)
    if (tid <= BV_LEVELS) sBase[tid] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += NT) { const orbb200_kp_t k = load(i); if (keep(k)) atomicAdd(&sBase[sorted ? 0 : k.octave], 1); }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int l = 0; l <= BV_LEVELS; l++) { const int c = sBase[l]; sBase[l] = acc; acc += c; }
        outCount[img] = acc;
    }
    __syncthreads();
    // pass 2: stable scatter, chunks of 256 in order
    for (int l = 0; l < (sorted ? 1 : BV_LEVELS); l++) {
        int base = sBase[l];
        for (int c0 = 0; c0 < n; c0 += NT) {
            const int i = c0 + tid;
            orbb200_kp_t k;
            bool f = false;
            if (i < n) { k = load(i); f = keep(k) && (sorted || k.octave == l); }
            const unsigned bal = __ballot_sync(0xffffffffu, f);
            if (lane == 0) sWarp[wid] = __popc(bal);
            __syncthreads();
            int before = 0, totalc = 0;
            for (int w = 0; w < NT / 32; w++) { if (w < wid) before += sWarp[w]; totalc += sWarp[w]; }
            if (f) O[base + before + __popc(bal & ((1u << lane) - 1u))] = k;
            base += totalc;
            __syncthreads();
        }
    }
}

// sepFilter2D(8U -> 8U, float getGaussianKernel(7, 2)) -- what GaussianBlur(7x7, 2, 2, BORDER_REFLECT_101) does on a
// sub-matrix (orb.cpp blurs the layers inside the packed buffer): rows = float sum k[i]*p[i] left to right; columns =
// k[3]*H[y] then += k[3+j]*(H[y+j] + H[y-j]); cvRound + saturate.  One thread per output pixel column segment.
struct GaussK { float k[7]; };

__global__ void __launch_bounds__(128) bird_blur_kernel(const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur, unsigned planeBytes, BirdGeom g,
                                                        GaussK gk, int nLevels)
{
    // One thread per (16-row block, 4 adjacent columns): a source row is three aligned words (12 bytes cover the 10 the four
    // outputs need), converted once and shared by the four horizontal sums; the products and sums keep sepFilter2D's order
    // (row filter: k0*s0 + k1*s1 + ... left to right; column filter: symmetric form), so the pixels are the ones cv::ORB blurs.
    // The first form (one thread per column, seven byte loads and conversions per pixel and row) took 0.21 ms per 128 images.
    const int level = blockIdx.z % BV_LEVELS, img = blockIdx.z / BV_LEVELS;
    if (level >= nLevels) return;
    const BirdLevel L = g.lv[level];
    const int nq = (L.w + 3) >> 2, nrb = (L.h + 15) >> 4;
    const int item = blockIdx.x * 128 + threadIdx.x;
    if (item >= nq * nrb) return;
    const int rb = item / nq, x = 4 * (item - rb * nq), y0 = 16 * rb;
    const uint8_t* S = pyr + (size_t)img * planeBytes + L.off;
    uint8_t* D = blur + (size_t)img * planeBytes + L.off;
    // horizontal sums of the last seven source rows in a register ring (slot = row % 7, the loop is unrolled by 7 so that every
    // slot index is a constant); output row r - 6 leaves as soon as source row r has arrived
    float H[7][4];
    const int rows = min(16, L.h - y0);
    const bool whole = x + 3 < L.w;
    for (int r0 = 0; r0 < rows + 6; r0 += 7) {
#pragma unroll
        for (int u = 0; u < 7; u++) {
            const int r = r0 + u;
            if (r < rows + 6) {
                const uint32_t* s = reinterpret_cast<const uint32_t*>(S + (ptrdiff_t)(y0 + r - 3) * L.pitch + x - 4);   // level rows are 4-byte aligned
                const uint32_t w0 = __ldg(s), w1 = __ldg(s + 1), w2 = __ldg(s + 2);
                float f[10];                                     // columns x-3 .. x+6
                f[0] = (float)((w0 >> 8) & 0xff); f[1] = (float)((w0 >> 16) & 0xff); f[2] = (float)(w0 >> 24);
                f[3] = (float)(w1 & 0xff); f[4] = (float)((w1 >> 8) & 0xff); f[5] = (float)((w1 >> 16) & 0xff); f[6] = (float)(w1 >> 24);
                f[7] = (float)(w2 & 0xff); f[8] = (float)((w2 >> 8) & 0xff); f[9] = (float)((w2 >> 16) & 0xff);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    float acc = __fmul_rn(gk.k[0], f[c]);
#pragma unroll
                    for (int k = 1; k < 7; k++) acc = __fadd_rn(acc, __fmul_rn(gk.k[k], f[c + k]));
                    H[u][c] = acc;
                }
                if (r >= 6) {
                    // rows r-6 .. r sit in slots (u+1)%7 .. u; centre = row r-3 = slot (u+4)%7
                    uint32_t out = 0;
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        float acc = __fmul_rn(gk.k[3], H[(u + 4) % 7][c]);
                        acc = __fadd_rn(acc, __fmul_rn(gk.k[4], __fadd_rn(H[(u + 5) % 7][c], H[(u + 3) % 7][c])));
                        acc = __fadd_rn(acc, __fmul_rn(gk.k[5], __fadd_rn(H[(u + 6) % 7][c], H[(u + 2) % 7][c])));
                        acc = __fadd_rn(acc, __fmul_rn(gk.k[6], __fadd_rn(H[u][c], H[(u + 1) % 7][c])));
                        const int v = __float2int_rn(acc);
                        out |= (uint32_t)min(max(v, 0), 255) << (8 * c);
                    }
                    uint8_t* d = D + (size_t)(y0 + r - 6) * L.pitch + x;
                    if (whole) *reinterpret_cast<uint32_t*>(d) = out;
                    else for (int c = 0; x + c < L.w; c++) d[c] = (uint8_t)(out >> (8 * c));
                }
            }
        }
    }
}

// computeOrbDescriptors (orb.cpp, WTA_K 2): one warp per keypoint, one descriptor byte per lane; the keypoint's angle is
// taken as given, its position is cvRound(pt / layerScale) on the blurred layer.
constexpr int BD_R = 19, BD_ROWS = 2 * BD_R + 1, BD_WORDS = 11;      // descriptor window: 39 rows x 44 bytes

__global__ void __launch_bounds__(256) bird_describe_kernel(BirdGeom g, const uint8_t* __restrict__ blur, const orbb200_kp_t* __restrict__ kps,
                                                            const int32_t* __restrict__ counts, uint8_t* __restrict__ desc)
{
    __shared__ float sPX[16 * 32], sPY[16 * 32];
    __shared__ unsigned sDescPatch[8][BD_ROWS * BD_WORDS];
    const int img = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 512; i += blockDim.x) {
        const int o = (i & 15) * 32 + (i >> 4);
        sPX[o] = (float)c_patX[i]; sPY[o] = (float)c_patY[i];
    }
    __syncthreads();
    // every CTA pays for the pattern table above, so a CTA walks over many keypoints of its image (the grid is a few CTAs
    // per image, not one per eight keypoints: 72 k CTAs, most of them past the image's count, cost 0.6 ms per 128 images)
    const int nk = min(counts[img], g.kpPerImg);
    for (int gk = blockIdx.x * (blockDim.x >> 5) + (tid >> 5); gk < nk; gk += gridDim.x * (blockDim.x >> 5)) {
        const orbb200_kp_t kp = kps[(size_t)img * g.kpPerImg + gk];
        const BirdLevel L = g.lv[kp.octave];
        constexpr float factorPI = (float)(3.14159265358979323846 / 180.0);
        const float ang = __fmul_rn(kp.angle, factorPI);
        double sd, cd;
        sincos((double)ang, &sd, &cd);                    // == host cosf/sinf for all but ~1e-8 of inputs
        const float a = (float)cd, b = (float)sd;
        const int cx = __float2int_rn(__fmul_rn(kp.x, L.invScale)), cy = __float2int_rn(__fmul_rn(kp.y, L.invScale));
        // The rotated pattern stays within 19 pixels of the centre (max radius 18.4): the warp first copies that 39-row window of
        // the blurred level into shared memory with aligned word loads (11 words per row from the column rounded down to 4; the
        // planes carry a 32-pixel margin, so the window is always inside the allocation), then every lane gathers its 16 bytes
        // from there.  512 scattered byte loads from global memory per keypoint were 3.4 ns per keypoint; this is the same
        // staging idea as the front camera's describe_kernel (which uses a TMA box).
        unsigned* patch = sDescPatch[tid >> 5];
        const int X0 = (cx - BD_R) & ~3, dxc = cx - X0;
        const uint8_t* win = blur + (size_t)img * g.planeBytes + L.off + (ptrdiff_t)(cy - BD_R) * L.pitch + X0;
        for (int i = lane; i < BD_ROWS * BD_WORDS; i += 32) {
            const int r = i / BD_WORDS, w = i - r * BD_WORDS;
            patch[i] = __ldg(reinterpret_cast<const unsigned*>(win + (size_t)r * L.pitch) + w);
        }
        __syncwarp();
        const uint8_t* bc = reinterpret_cast<const uint8_t*>(patch) + BD_R * (BD_WORDS * 4) + dxc;
        int val = 0;
    #pragma unroll
        for (int k = 0; k < 8; k++) {
            int t[2];
    #pragma unroll
            for (int e = 0; e < 2; e++) {
                const int idx = (2 * k + e) * 32 + lane;
                const float px = sPX[idx], py = sPY[idx];
                const int xx = __float_as_int(__fadd_rn(__fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b)), 12582912.f)) - 0x4B400000;
                const int yy = __float_as_int(__fadd_rn(__fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a)), 12582912.f)) - 0x4B400000;
                t[e] = bc[yy * (BD_WORDS * 4) + xx];
            }
            val |= (t[0] < t[1]) << k;
        }
        desc[((size_t)img * g.kpPerImg + gk) * 32 + lane] = (uint8_t)val;
        __syncwarp();                       // the next keypoint's window overwrites this warp's patch
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
BirdState& state(Ctx& c)
{
    if (!c.bird) c.bird = new BirdState();
    return *static_cast<BirdState*>(c.bird);
}

void free_plan(BirdPlan* p)
{
    void* ptrs[] = {p->d_tab, p->d_cells, p->d_pyr, p->d_mask, p->d_blur, p->d_cand, p->d_candCount, p->d_lvlKp, p->d_lvlCount,
                    p->d_kps, p->d_kps2, p->d_desc, p->d_counts, p->d_counts2, p->d_pts, p->d_slow, p->d_list[0], p->d_list[1], p->d_pts0, p->d_iters, p->d_maskShared,
                    p->d_qx, p->d_qy, p->d_qangle, p->d_qlevel, p->d_qvalid, p->d_carryKps, p->d_carryDesc, p->d_carryCount};
    for (void* q : ptrs) if (q) cudaFree(q);
    delete p;
}

void exact_table(int ssize, int dsize, std::vector<int2>& tab)
{
    // resize.cpp interpolationLinear<ufixedpoint16>: coordinates in (soft)double, coefficients rounded to 8.8
    const double scale = 1.0 / ((double)dsize / ssize);
    for (int v = 0; v < dsize; v++) {
        const double fval = scale * ((double)v + 0.5) - 0.5;
        const int ival = cvFloorD(fval);
        int2 e = make_int2(0, 0);
        if (ival >= 0 && ssize > 1) {
            if (ival < ssize - 1) { e.x = ival; e.y = (int)std::nearbyint((fval - (double)ival) * 256.0); }
            else e.x = ssize - 1;
        }
        tab.push_back(e);
    }
}

BirdPlan* get_plan(Ctx& c, int w, int h, int nfeatures, int batch)
{
    BirdState& S = state(c);
    const auto key = std::make_tuple(w, h, nfeatures);
    auto it = S.plans.find(key);
    BirdPlan* p = it == S.plans.end() ? nullptr : it->second;
    if (p && p->batch >= batch) return p;
    c.allocEpoch++;
    if (p) { cudaStreamSynchronize(c.stream); free_plan(p); S.plans.erase(key); }
    p = new BirdPlan();
    BirdGeom& g = p->g;
    g.w = w; g.h = h; g.nfeatures = nfeatures;
    // per-level quotas (orb.cpp computeKeyPoints)
    const double scaleFactor = (double)1.2f;
    const float factor = (float)(1.0 / scaleFactor);
    float nd = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)BV_LEVELS));
    int sum = 0;
    size_t off = 0, candOff = 0;
    int kpOff = 0;
    std::vector<int2> tab;
    std::vector<int4> cells;
    for (int l = 0; l < BV_LEVELS; l++) {
        BirdLevel& L = g.lv[l];
        L.scale = (float)std::pow(scaleFactor, (double)l);
        L.invScale = 1.0f / L.scale;
        L.w = cvRoundF(w * L.invScale); L.h = cvRoundF(h * L.invScale);
        if (l < BV_LEVELS - 1) { L.quota = cvRoundF(nd); sum += L.quota; nd *= factor; }
        else L.quota = std::max(nfeatures - sum, 0);
        L.pitch = (int)align_up((size_t)L.w + 2 * BV_MARGIN, 128);
        L.off = (unsigned)(off + (size_t)BV_MARGIN * L.pitch + BV_MARGIN);
        off += align_up((size_t)(L.h + 2 * BV_MARGIN) * L.pitch, 128);
        L.candCap = std::min(std::max((L.w * L.h) / 4, 64), BV_SORT_CAP);
        L.candOff = (unsigned)candOff; candOff += L.candCap;
        L.kpCap = std::min(L.candCap, 2 * L.quota + 64);
        L.kpOff = kpOff; kpOff += L.kpCap;
        if (l > 0) {
            L.tabX = (int)tab.size(); exact_table(g.lv[l - 1].w, L.w, tab);
            L.tabY = (int)tab.size(); exact_table(g.lv[l - 1].h, L.h, tab);
        }
        // whole-level FAST as tiles: a tile's cell image is the emit tile grown by 4 (3 ring pixels + 1 for the NMS
        // neighbourhood), clipped to the level; corners closer than edgeThreshold to the level border are dropped by
        // runByImageBorder anyway, so only [31, w-31) x [31, h-31) is emitted
        if (L.w > 2 * BV_EDGE && L.h > 2 * BV_EDGE)
            for (int ty = BV_EDGE; ty < L.h - BV_EDGE; ty += BV_TILE)
                for (int tx = BV_EDGE; tx < L.w - BV_EDGE; tx += BV_TILE) {
                    const int ex1 = std::min(tx + BV_TILE, L.w - BV_EDGE), ey1 = std::min(ty + BV_TILE, L.h - BV_EDGE);
                    push_fast_cell(cells, p->need, tx - 4, ty - 4, std::min(ex1 + 4, L.w), std::min(ey1 + 4, L.h), l, L.off, L.pitch,
                                   L.candOff, L.candCap, tx, ty, ex1, ey1);
                }
    }
    g.planeBytes = (unsigned)align_up(off, 256);
    g.candPerImg = (unsigned)candOff;
    g.kpPerImg = kpOff;
    p->nCells = (int)(cells.size() / 3);
    p->batch = batch;
    const size_t B = (size_t)batch;
    bool ok = cudaMalloc((void**)&p->d_tab, std::max<size_t>(tab.size(), 1) * sizeof(int2)) == cudaSuccess &&
              cudaMalloc((void**)&p->d_cells, std::max<size_t>(cells.size(), 1) * sizeof(int4)) == cudaSuccess &&
              cudaMalloc((void**)&p->d_pyr, B * g.planeBytes) == cudaSuccess && cudaMalloc((void**)&p->d_mask, B * g.planeBytes) == cudaSuccess &&
              cudaMalloc((void**)&p->d_blur, B * g.planeBytes) == cudaSuccess &&
              cudaMalloc((void**)&p->d_cand, B * g.candPerImg * 4) == cudaSuccess &&
              cudaMalloc((void**)&p->d_candCount, B * MAX_LEVELS * 4) == cudaSuccess &&
              cudaMalloc((void**)&p->d_lvlKp, B * g.kpPerImg * sizeof(float4)) == cudaSuccess &&
              cudaMalloc((void**)&p->d_lvlCount, B * BV_LEVELS * 4) == cudaSuccess &&
              cudaMalloc((void**)&p->d_kps, B * g.kpPerImg * sizeof(orbb200_kp_t)) == cudaSuccess &&
              cudaMalloc((void**)&p->d_kps2, B * g.kpPerImg * sizeof(orbb200_kp_t)) == cudaSuccess &&
              cudaMalloc((void**)&p->d_desc, B * g.kpPerImg * 32) == cudaSuccess &&
              cudaMalloc((void**)&p->d_counts, B * 4) == cudaSuccess && cudaMalloc((void**)&p->d_counts2, B * 4) == cudaSuccess &&
              cudaMalloc((void**)&p->d_pts, B * g.kpPerImg * 8) == cudaSuccess &&
              cudaMalloc((void**)&p->d_slow, B * g.kpPerImg * 4) == cudaSuccess &&
              cudaMalloc((void**)&p->d_list[0], B * g.kpPerImg * 4) == cudaSuccess && cudaMalloc((void**)&p->d_list[1], B * g.kpPerImg * 4) == cudaSuccess &&
              cudaMalloc((void**)&p->d_pts0, B * g.kpPerImg * 8) == cudaSuccess && cudaMalloc((void**)&p->d_iters, B * g.kpPerImg * 4) == cudaSuccess;
    if (ok) {
        ok = cudaMemcpyAsync(p->d_tab, tab.data(), tab.size() * sizeof(int2), cudaMemcpyHostToDevice, c.stream) == cudaSuccess &&
             cudaMemcpyAsync(p->d_cells, cells.data(), cells.size() * sizeof(int4), cudaMemcpyHostToDevice, c.stream) == cudaSuccess &&
             cudaMemsetAsync(p->d_mask, 0, B * g.planeBytes, c.stream) == cudaSuccess &&
             cudaMemsetAsync(p->d_pyr, 0, B * g.planeBytes, c.stream) == cudaSuccess &&
             cudaStreamSynchronize(c.stream) == cudaSuccess;
    }
    if (!ok) { cudaGetLastError(); free_plan(p); c.err = "bird: cudaMalloc failed"; return nullptr; }
    S.plans[key] = p;
    return p;
}

// upload n images (host pointers) into level 0 of the planes; mask optional
int upload_images(Ctx& c, BirdPlan* p, const uint8_t* const* imgs, const uint8_t* const* masks, int n, size_t stride, size_t mstride)
{
    const BirdGeom& g = p->g;
    const BirdLevel& L0 = g.lv[0];
    if (masks) p->maskCacheValid = false;           // image 0's mask slot is about to hold something else (see orbb200_bird_extract_batch)
    for (int i = 0; i < n; i++) {
        ORBB200_CUDA_OK(c, cudaMemcpy2DAsync(p->d_pyr + (size_t)i * g.planeBytes + L0.off, L0.pitch, imgs[i], stride, g.w, g.h,
                                             cudaMemcpyHostToDevice, c.stream));
        if (masks)
            ORBB200_CUDA_OK(c, cudaMemcpy2DAsync(p->d_mask + (size_t)i * g.planeBytes + L0.off, L0.pitch, masks[i], mstride, g.w, g.h,
                                                 cudaMemcpyHostToDevice, c.stream));
    }
    return ORBB200_OK;
}

void enqueue_pyramid(Ctx& c, BirdPlan* p, int n, bool withMask, int nLevels)
{
    const BirdGeom& g = p->g;
    if (n >= 8 && !withMask) {              // batches: one CTA builds an image's whole pyramid
        bird_pyramid_kernel<<<n, 1024, 0, c.stream>>>(p->d_pyr, g.planeBytes, g, p->d_tab, nLevels);
        c.launches++;
        return;
    }
    for (int l = 1; l < nLevels; l++) {
        const BirdLevel& D = g.lv[l];
        dim3 grid((D.w + 127) / 128, D.h, n);
        bird_resize_kernel<<<grid, 128, 0, c.stream>>>(p->d_pyr, g.planeBytes, g.lv[l - 1], D, p->d_tab, 0);
        c.launches++;
        if (withMask) {
            bird_resize_kernel<<<grid, 128, 0, c.stream>>>(p->d_mask, g.planeBytes, g.lv[l - 1], D, p->d_tab, 1);
            c.launches++;
        }
    }
    bird_border_kernel<<<dim3(64, nLevels, n), 256, 0, c.stream>>>(p->d_pyr, g.planeBytes, g);
    c.launches++;
}

// maskMode: 0 none; 1 one mask per image in p->d_mask level 0 (its pyramid is built here); 2 the shared mask pyramid of
// orbb200_bird_set_mask (p->d_maskShared, already built)
void enqueue_blur(Ctx& c, BirdPlan* p, int n, int nLevels, cudaStream_t stream)
{
    const BirdGeom& g = p->g;
    GaussK gk;
    {   // cv::getGaussianKernel(7, 2, CV_32F)
        double sum = 0;
        for (int i = 0; i < 7; i++) { const double x = i - 3.0; gk.k[i] = (float)std::exp(-0.5 / 4.0 * x * x); sum += gk.k[i]; }
        sum = 1. / sum;
        for (int i = 0; i < 7; i++) gk.k[i] = (float)(gk.k[i] * sum);
    }
    int maxW = 1, maxH = 1;
    for (int l = 0; l < nLevels; l++) { maxW = std::max(maxW, g.lv[l].w); maxH = std::max(maxH, g.lv[l].h); }
    bird_blur_kernel<<<dim3((((maxW + 3) / 4) * ((maxH + 15) / 16) + 127) / 128, 1, n * BV_LEVELS), 128, 0, stream>>>(p->d_pyr, p->d_blur, g.planeBytes, g, gk, nLevels);
    c.launches++;
}

// forkBlur: the caller will run enqueue_compute on the same pyramid (detect + cornerSubPix + compute in one call).  For one or two
// images the blurred pyramid -- which only compute's descriptors read and which depends on nothing but the pyramid -- is then
// built on a side stream beside FAST, retainBest and cornerSubPix instead of after them.
int enqueue_detect(Ctx& c, BirdPlan* p, int n, int maskMode, bool* forkBlur = nullptr, float* ptsOut = nullptr, bool maskPyramidCached = false)
{
    const BirdGeom& g = p->g;
    { StageTimer t(c, 9); enqueue_pyramid(c, p, n, maskMode == 1 && !maskPyramidCached, BV_LEVELS); }
    if (forkBlur) *forkBlur = false;
    if (forkBlur && n <= 2 && !c.timing && c.stream4) {
        cudaEventRecord(c.evFork4, c.stream);
        cudaStreamWaitEvent(c.stream4, c.evFork4, 0);
        enqueue_blur(c, p, n, BV_LEVELS, c.stream4);
        cudaEventRecord(c.evJoin4, c.stream4);
        *forkBlur = true;                            // the caller hands this to enqueue_compute, which joins the side stream
    }
    StageTimer t(c, 10);
    cudaMemsetAsync(p->d_candCount, 0, sizeof(int32_t) * MAX_LEVELS * n, c.stream);
    launch_fast_cells(c, p->d_pyr, g.planeBytes, g.candPerImg, BV_FAST_TH, BV_FAST_TH, 1, p->d_cells, p->nCells, p->need, p->d_cand,
                      p->d_candCount, n);
    const size_t smem = (size_t)BV_SORT_CAP * 12;      // keys + responses + two u16 stopper lists
    if (smem > ensure_max_dynamic_smem(c.device, (const void*)bird_select_kernel<SEL_THREADS>, SMEM_BIRD_SELECT) ||
        smem > ensure_max_dynamic_smem(c.device, (const void*)bird_select_kernel<SEL_THREADS_FEW>, SMEM_BIRD_SELECT_FEW)) {
        c.err = "bird_select_kernel: shared memory";
        return ORBB200_ERR_CUDA;
    }
    const uint8_t* mask = maskMode == 1 ? p->d_mask : maskMode == 2 ? p->d_maskShared : nullptr;
    if (n <= 2 && !c.selectTiers) {
        // one or two images: every level at once, one 1024-thread CTA each with the largest carve (an SM per level is there for the
        // taking; the tiers below would run one after the other, and level 0's sort and partitions are 4x shorter per thread)
        bird_select_kernel<SEL_THREADS_FEW><<<dim3(n, BV_LEVELS), SEL_THREADS_FEW, smem, c.stream>>>(
            g, p->d_pyr, mask, maskMode == 1 ? g.planeBytes : 0u, p->d_cand, p->d_candCount, p->d_lvlKp, p->d_lvlCount, c.d_status, BV_SORT_CAP, BV_SORT_CAP, 0, 0);
        c.launches++;
        bird_finish_kernel<<<dim3((g.kpPerImg + 7) / 8, n), 256, 0, c.stream>>>(g, p->d_pyr, p->d_lvlKp, p->d_lvlCount, p->d_kps, p->d_counts, ptsOut);
        c.launches++;
        ORBB200_CUDA_OK(c, cudaGetLastError());
        return ORBB200_OK;
    }
    struct Tier { int key, n; };
    const Tier tiers[3] = {{2048, 2048}, {8192, 5120}, {BV_SORT_CAP, BV_SORT_CAP}};
    for (int t = 0; t < 3; t++) {
        // levels that can reach this tier at all (candCap is the NMS bound w*h/4 of a level; levels shrink with their index)
        int nl = BV_LEVELS;
        if (t > 0) { nl = 0; while (nl < BV_LEVELS && g.lv[nl].candCap > tiers[t - 1].n) nl++; }
        if (nl == 0) break;
        bird_select_kernel<SEL_THREADS><<<dim3(n, nl), SEL_THREADS, (size_t)tiers[t].key * 4 + (size_t)tiers[t].n * 8, c.stream>>>(
            g, p->d_pyr, mask, maskMode == 1 ? g.planeBytes : 0u, p->d_cand, p->d_candCount, p->d_lvlKp, p->d_lvlCount, c.d_status,
            tiers[t].key, tiers[t].n, t ? tiers[t - 1].key : 0, t ? tiers[t - 1].n : 0);
        if (t < 2) c.launches++;
    }
    c.launches++;
    bird_finish_kernel<<<dim3((g.kpPerImg + 7) / 8, n), 256, 0, c.stream>>>(g, p->d_pyr, p->d_lvlKp, p->d_lvlCount, p->d_kps, p->d_counts, ptsOut);
    c.launches++;
    ORBB200_CUDA_OK(c, cudaGetLastError());
    return ORBB200_OK;
}

int ensure_win_mask(Ctx& c, int winW, int winH)
{
    BirdState& S = state(c);
    if (S.winW == winW && S.winH == winH) return ORBB200_OK;
    const int ww = 2 * winW + 1, wh = 2 * winH + 1;
    std::vector<float> m((size_t)ww * wh);
    for (int i = 0; i < wh; i++) {
        const float y = (float)(i - winH) / winH;
        const float vy = std::exp(-y * y);
        for (int j = 0; j < ww; j++) {
            const float x = (float)(j - winW) / winW;
            m[(size_t)i * ww + j] = (float)(vy * std::exp(-x * x));
        }
    }
    if (!S.d_winMask) ORBB200_CUDA_OK(c, cudaMalloc((void**)&S.d_winMask, sizeof(float) * (2 * BV_MAX_WIN + 1) * (2 * BV_MAX_WIN + 1)));
    if (!S.d_winMaskD) ORBB200_CUDA_OK(c, cudaMalloc((void**)&S.d_winMaskD, sizeof(double) * (2 * BV_MAX_WIN + 1) * (2 * BV_MAX_WIN + 1)));
    std::vector<double> md(m.begin(), m.end());
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(S.d_winMask, m.data(), m.size() * sizeof(float), cudaMemcpyHostToDevice, c.stream));
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(S.d_winMaskD, md.data(), md.size() * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
    S.winW = winW; S.winH = winH;
    return ORBB200_OK;
}

int enqueue_subpix(Ctx& c, BirdPlan* p, int n, const int32_t* d_counts, int nFixed, int winW, int winH, int maxCount, double epsilon)
{
    const BirdGeom& g = p->g;
    if (winW <= 0 || winH <= 0 || winW > BV_MAX_WIN || winH > BV_MAX_WIN) { c.err = "cornerSubPix: window half-size must be in 1..7"; return ORBB200_ERR_ARG; }
    if (g.w < winW * 2 + 5 || g.h < winH * 2 + 5) { c.err = "cornerSubPix: image smaller than the window"; return ORBB200_ERR_ARG; }
    const int rc = ensure_win_mask(c, winW, winH);
    if (rc != ORBB200_OK) return rc;
    const int maxIters = std::min(std::max(maxCount, 1), 100);
    double eps = std::max(epsilon, 0.);
    eps *= eps;
    const BirdLevel& L0 = g.lv[0];
    BirdState& S = state(c);
    if (!S.d_work) ORBB200_CUDA_OK(c, cudaMalloc((void**)&S.d_work, 4 * sizeof(int)));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c.device);
    if (winW == S5_WIN && winH == S5_WIN && n >= 8) {    // (fewer images: too few corners to fill the lanes, the warp-per-corner form below is 2.4x faster)
        // the reference's window: streaming form in ceil(maxIters / budget) phases + exact hand-over of the few corners
        // whose patch is not inside the image
        const int phases = (maxIters + S5_BUDGET - 1) / S5_BUDGET;
        const size_t workInts = 2 * (size_t)phases + 8;
        if (S.workInts < workInts) {
            cudaStreamSynchronize(c.stream);
            if (S.d_work5) cudaFree(S.d_work5);
            S.d_work5 = nullptr; S.workInts = 0;
            ORBB200_CUDA_OK(c, cudaMalloc((void**)&S.d_work5, workInts * sizeof(int)));
            S.workInts = workInts;
        }
        cudaMemsetAsync(S.d_work5, 0, workInts * sizeof(int), c.stream);
        // reach of the staged patch: 5 (two CTAs per SM), 3 (three) or 2 (four CTAs per SM at 125 registers: 16 resident warps)
        const int ctas = c.subpixCtasPerSm;
        const size_t smem = sizeof(unsigned) * S5_THREADS * (ctas <= 2 ? S5Geom<5>::STRIDE : ctas == 3 ? S5Geom<3>::STRIDE : S5Geom<2>::STRIDE);
        const void* kfn = ctas <= 2 ? (const void*)bird_subpix5_kernel<5, 2> : ctas == 3 ? (const void*)bird_subpix5_kernel<3, 3> : (const void*)bird_subpix5_kernel<2, 4>;
        if (smem > ensure_max_dynamic_smem(c.device, kfn, ctas <= 2 ? SMEM_BIRD_SUBPIX : ctas == 3 ? SMEM_BIRD_SUBPIX3 : SMEM_BIRD_SUBPIX2)) { c.err = "bird_subpix5_kernel: shared memory"; return ORBB200_ERR_CUDA; }
        const long long slots = (long long)n * g.kpPerImg;
        const int grid = std::max(1, (int)std::min<long long>((long long)sms * c.subpixCtasPerSm, (slots + S5_THREADS - 1) / S5_THREADS));
        int* heads = S.d_work5;                      // [phases]
        int* outCounts = S.d_work5 + phases;         // [phases]
        int* slowCount = S.d_work5 + 2 * phases;     // + head of the generic kernel
        S5Args A{};
        A.imgs = p->d_pyr + L0.off; A.imgStrideBytes = g.planeBytes; A.pitch = L0.pitch; A.cols = g.w; A.rows = g.h;
        A.pts = p->d_pts; A.pts0 = p->d_pts0; A.iters = p->d_iters; A.ptsPerImg = g.kpPerImg; A.nImages = n; A.counts = d_counts; A.nFixed = nFixed;
        A.winMaskD = S.d_winMaskD; A.maxIters = maxIters; A.eps = eps; A.slowList = p->d_slow; A.slowCount = slowCount;
        for (int ph = 0; ph < phases; ph++) {
            A.listIn = ph == 0 ? nullptr : p->d_list[(ph - 1) & 1];
            A.listInCount = ph == 0 ? nullptr : outCounts + (ph - 1);
            A.head = heads + ph;
            A.listOut = p->d_list[ph & 1]; A.listOutCount = outCounts + ph;
            if (ctas <= 2) bird_subpix5_kernel<5, 2><<<grid, S5_THREADS, smem, c.stream>>>(A);
            else if (ctas == 3) bird_subpix5_kernel<3, 3><<<grid, S5_THREADS, smem, c.stream>>>(A);
            else bird_subpix5_kernel<2, 4><<<grid, S5_THREADS, smem, c.stream>>>(A);
            c.launches++;
        }
        bird_subpix_thread_kernel<<<std::min(sms, std::max(n, 1) * 4), SP_THREADS, 0, c.stream>>>(
            p->d_pyr + L0.off, g.planeBytes, L0.pitch, g.w, g.h, p->d_pts, g.kpPerImg, n, d_counts, nFixed, S.d_winMask, winW, winH, maxIters, eps, slowCount + 1,
            p->d_slow, slowCount);
    }
    else if (n >= 8) {    // many corners in flight: persistent one-thread-per-corner form
        cudaMemsetAsync(S.d_work, 0, 4 * sizeof(int), c.stream);
        bird_subpix_thread_kernel<<<sms * 8, SP_THREADS, 0, c.stream>>>(
            p->d_pyr + L0.off, g.planeBytes, L0.pitch, g.w, g.h, p->d_pts, g.kpPerImg, n, d_counts, nFixed, S.d_winMask, winW, winH, maxIters, eps, S.d_work,
            nullptr, nullptr);
    }
    else if (winW == S5_WIN && winH == S5_WIN && !c.subpixGenericWarp)
        bird_subpix_warp5_kernel<<<dim3((g.kpPerImg + SP_WARPS - 1) / SP_WARPS, n), SP_WARPS * 32, 0, c.stream>>>(
            p->d_pyr + L0.off, g.planeBytes, L0.pitch, g.w, g.h, p->d_pts, (size_t)g.kpPerImg, d_counts, nFixed, S.d_winMaskD, maxIters, eps);
    else
        bird_subpix_kernel<<<dim3((g.kpPerImg + SP_WARPS - 1) / SP_WARPS, n), SP_WARPS * 32, 0, c.stream>>>(
            p->d_pyr + L0.off, g.planeBytes, L0.pitch, g.w, g.h, p->d_pts, (size_t)g.kpPerImg, d_counts, nFixed, S.d_winMask, winW, winH, maxIters, eps);
    c.launches++;
    ORBB200_CUDA_OK(c, cudaGetLastError());
    return ORBB200_OK;
}

// d_kps/d_counts -> filtered into d_kps2/d_counts2, descriptors in d_desc
int enqueue_compute(Ctx& c, BirdPlan* p, int n, int nLevels, bool blurForked = false, const float* refinedPts = nullptr)
{
    const BirdGeom& g = p->g;
    if (n <= 2) bird_filter_kernel<1024><<<n, 1024, 0, c.stream>>>(g, p->d_kps, p->d_counts, p->d_kps2, p->d_counts2, refinedPts);
    else bird_filter_kernel<256><<<n, 256, 0, c.stream>>>(g, p->d_kps, p->d_counts, p->d_kps2, p->d_counts2, refinedPts);
    c.launches++;
    if (blurForked) {                               // enqueue_detect started it on the side stream
        cudaStreamWaitEvent(c.stream, c.evJoin4, 0);
    } else {
        enqueue_blur(c, p, n, nLevels, c.stream);
    }
    bird_describe_kernel<<<dim3(n >= 8 ? 12 : (g.kpPerImg + 7) / 8, n), 256, 0, c.stream>>>(g, p->d_blur, p->d_kps2, p->d_counts2, p->d_desc);
    c.launches++;
    ORBB200_CUDA_OK(c, cudaGetLastError());
    return ORBB200_OK;
}

int check_bird_status(Ctx& c)
{
    int32_t st = 0;
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(&st, c.d_status, sizeof(st), cudaMemcpyDeviceToHost, c.stream));
    ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
    if (st != 0) {
        cudaMemsetAsync(c.d_status, 0, sizeof(int32_t), c.stream);
        c.err = st == 3 ? "bird: more FAST corners on a level than the selection kernel holds (16384)"
                        : "bird: more keypoints tie at the retainBest threshold than the output holds";
        return ORBB200_ERR_CAPACITY;
    }
    return ORBB200_OK;
}

int download(Ctx& c, BirdPlan* p, int n, const orbb200_kp_t* d_kps, const int32_t* d_counts, orbb200_kp_t* kps, uint8_t* desc, int cap, int* n_out)
{
    const BirdGeom& g = p->g;
    std::vector<int32_t> cnt(n);
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(cnt.data(), d_counts, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c.stream));
    const int rc = check_bird_status(c);        // synchronises
    if (rc != ORBB200_OK) return rc;
    for (int i = 0; i < n; i++) {
        const int m = std::min(cnt[i], cap);
        if (cnt[i] > cap) { c.err = "bird: output capacity too small"; return ORBB200_ERR_CAPACITY; }
        if (m > 0 && kps) ORBB200_CUDA_OK(c, cudaMemcpyAsync(kps + (size_t)i * cap, d_kps + (size_t)i * g.kpPerImg, sizeof(orbb200_kp_t) * m, cudaMemcpyDeviceToHost, c.stream));
        if (m > 0 && desc) ORBB200_CUDA_OK(c, cudaMemcpyAsync(desc + (size_t)i * cap * 32, p->d_desc + (size_t)i * g.kpPerImg * 32, (size_t)m * 32, cudaMemcpyDeviceToHost, c.stream));
        n_out[i] = m;
    }
    ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
    return ORBB200_OK;
}

}  // namespace

// ---- device-resident pipeline for the batched frame step (api.cu) -------------------------------------------------------
static int ensure_step_buffers(Ctx& c, BirdPlan* p)
{
    if (p->d_qx) return ORBB200_OK;
    c.allocEpoch++;
    const size_t Q = ((size_t)p->batch + 1) * p->g.kpPerImg;
    const size_t K = (size_t)p->g.kpPerImg;
    const bool ok = cudaMalloc((void**)&p->d_qx, Q * 4) == cudaSuccess && cudaMalloc((void**)&p->d_qy, Q * 4) == cudaSuccess &&
                    cudaMalloc((void**)&p->d_qangle, Q * 4) == cudaSuccess && cudaMalloc((void**)&p->d_qlevel, Q * 4) == cudaSuccess &&
                    cudaMalloc((void**)&p->d_qvalid, Q) == cudaSuccess && cudaMalloc((void**)&p->d_carryKps, K * sizeof(orbb200_kp_t)) == cudaSuccess &&
                    cudaMalloc((void**)&p->d_carryDesc, K * 32) == cudaSuccess && cudaMalloc((void**)&p->d_carryCount, 4) == cudaSuccess;
    if (!ok) { cudaGetLastError(); c.err = "bird step: cudaMalloc failed"; return ORBB200_ERR_CUDA; }
    ORBB200_CUDA_OK(c, cudaMemsetAsync(p->d_carryCount, 0, 4, c.stream));
    p->carryValid = false;
    return ORBB200_OK;
}

int bird_set_mask(Ctx& c, int w, int h, int nfeatures, int batch, const uint8_t* mask, size_t stride)
{
    BirdPlan* p = get_plan(c, w, h, nfeatures, std::max(batch, 1));
    if (!p) return ORBB200_ERR_CUDA;
    const BirdGeom& g = p->g;
    c.allocEpoch++;                             // captured frame steps hold the mask pointer (or its absence)
    if (!mask) {
        if (p->d_maskShared) { cudaStreamSynchronize(c.stream); cudaFree(p->d_maskShared); p->d_maskShared = nullptr; }
        return ORBB200_OK;
    }
    if (!p->d_maskShared) {
        ORBB200_CUDA_OK(c, cudaMalloc((void**)&p->d_maskShared, g.planeBytes));
        ORBB200_CUDA_OK(c, cudaMemsetAsync(p->d_maskShared, 0, g.planeBytes, c.stream));      // zero margin: masked outside the image
    }
    const BirdLevel& L0 = g.lv[0];
    ORBB200_CUDA_OK(c, cudaMemcpy2DAsync(p->d_maskShared + L0.off, L0.pitch, mask, stride, g.w, g.h, cudaMemcpyHostToDevice, c.stream));
    for (int l = 1; l < BV_LEVELS; l++) {       // resize(prev mask level, INTER_LINEAR_EXACT) + threshold(254, TOZERO) (orb.cpp)
        const BirdLevel& D = g.lv[l];
        bird_resize_kernel<<<dim3((D.w + 127) / 128, D.h, 1), 128, 0, c.stream>>>(p->d_maskShared, g.planeBytes, g.lv[l - 1], D, p->d_tab, 1);
        c.launches++;
    }
    ORBB200_CUDA_OK(c, cudaGetLastError());
    ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));      // the host mask may be pageable memory
    return ORBB200_OK;
}

// cv::ORB detect(mask) + cornerSubPix(5,5; 40; 1e-3) + compute on n device-resident images, everything enqueued on the
// context's stream; then the query arrays of the frame-to-frame birdview matching.  chain == false forgets the carried frame.
static void fill_step_view(BirdPlan* p, BirdStepView* out)
{
    out->d_kps = p->d_kps2; out->d_desc = p->d_desc; out->d_counts = p->d_counts2; out->kpPerImg = p->g.kpPerImg;
    out->d_qx = p->d_qx; out->d_qy = p->d_qy; out->d_qangle = p->d_qangle; out->d_qlevel = p->d_qlevel; out->d_qvalid = p->d_qvalid;
    out->d_carryDesc = p->d_carryDesc;
    out->plan = p;
}

// the pools a step of n images will use (allocated on first use), without enqueueing anything
int bird_step_view(Ctx& c, int w, int h, int nfeatures, int n, BirdStepView* out)
{
    BirdPlan* p = get_plan(c, w, h, nfeatures, n);
    if (!p) return ORBB200_ERR_CUDA;
    const int rc = ensure_step_buffers(c, p);
    if (rc != ORBB200_OK) return rc;
    fill_step_view(p, out);
    return ORBB200_OK;
}

int bird_step_enqueue(Ctx& c, int w, int h, int nfeatures, int n, const uint8_t* d_imgs, size_t imgBytes, size_t stride, bool chain, BirdStepView* out)
{
    BirdPlan* p = get_plan(c, w, h, nfeatures, n);
    if (!p) return ORBB200_ERR_CUDA;
    int rc = ensure_step_buffers(c, p);
    if (rc != ORBB200_OK) return rc;
    const BirdGeom& g = p->g;
    if (!chain || !p->carryValid) ORBB200_CUDA_OK(c, cudaMemsetAsync(p->d_carryCount, 0, 4, c.stream));
    {
        StageTimer t(c, 9);
        bird_import_kernel<<<dim3(g.h, n), 128, 0, c.stream>>>(d_imgs, imgBytes, stride, p->d_pyr, g.planeBytes, g.lv[0]);
        c.launches++;
    }
    bool blurForked = false;
    // (the detector writes the (x, y) list cornerSubPix refines, and compute's border filter reads the refined list: no copies between)
    rc = enqueue_detect(c, p, n, p->d_maskShared ? 2 : 0, &blurForked, p->d_pts);
    if (rc != ORBB200_OK) return rc;
    {
        StageTimer t(c, 11);
        if (g.w >= 15 && g.h >= 15) {
            rc = enqueue_subpix(c, p, n, p->d_counts, 0, 5, 5, 40, 0.001);
            if (rc != ORBB200_OK) return rc;
        }
    }
    {
        StageTimer t(c, 12);
        rc = enqueue_compute(c, p, n, BV_LEVELS, blurForked, p->d_pts);
        if (rc != ORBB200_OK) return rc;
        bird_queries_kernel<<<dim3((g.kpPerImg + 255) / 256, n + 1), 256, 0, c.stream>>>(p->d_kps2, p->d_counts2, p->d_carryKps, p->d_carryCount, g.kpPerImg,
                                                                                        p->d_qx, p->d_qy, p->d_qangle, p->d_qlevel, p->d_qvalid);
        c.launches++;
    }
    ORBB200_CUDA_OK(c, cudaGetLastError());
    fill_step_view(p, out);
    return ORBB200_OK;
}

// after the matching has been enqueued: the last frame of this step becomes the carried frame of the next one
int bird_step_carry(Ctx& c, const BirdStepView& v, int n)
{
    BirdPlan* p = static_cast<BirdPlan*>(v.plan);
    const size_t K = (size_t)p->g.kpPerImg;
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(p->d_carryKps, p->d_kps2 + (size_t)(n - 1) * K, K * sizeof(orbb200_kp_t), cudaMemcpyDeviceToDevice, c.stream));
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(p->d_carryDesc, p->d_desc + (size_t)(n - 1) * K * 32, K * 32, cudaMemcpyDeviceToDevice, c.stream));
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(p->d_carryCount, p->d_counts2 + (n - 1), 4, cudaMemcpyDeviceToDevice, c.stream));
    p->carryValid = true;
    return ORBB200_OK;
}

int bird_step_status(Ctx& c) { return check_bird_status(c); }

void bird_destroy(Ctx& c)
{
    if (!c.bird) return;
    BirdState* S = static_cast<BirdState*>(c.bird);
    for (auto& kv : S->plans) free_plan(kv.second);
    if (S->d_winMask) cudaFree(S->d_winMask);
    if (S->d_winMaskD) cudaFree(S->d_winMaskD);
    if (S->d_work) cudaFree(S->d_work);
    if (S->d_work5) cudaFree(S->d_work5);
    delete S;
    c.bird = nullptr;
}

}  // namespace orbb200

using namespace orbb200;

#define BIRD_ENTER(ctx)                                             \
    if (!(ctx)) return ORBB200_ERR_ARG;                             \
    Ctx& c = (ctx)->c;                                              \
    ORBB200_CUDA_OK(c, cudaSetDevice(c.device))

extern "C" {

int orbb200_bird_max_keypoints(orbb200_ctx* ctx, int w, int h, int nfeatures)
{
    BIRD_ENTER(ctx);
    if (w <= 0 || h <= 0 || nfeatures <= 0) return 0;
    BirdPlan* p = get_plan(c, w, h, nfeatures, 1);
    return p ? p->g.kpPerImg : 0;
}

int orbb200_bird_detect(orbb200_ctx* ctx, const uint8_t* img, const uint8_t* mask, int w, int h, size_t stride, size_t mask_stride,
                        int nfeatures, orbb200_kp_t* kps, int cap, int* n_out)
{
    BIRD_ENTER(ctx);
    if (!img || !kps || !n_out || w <= 0 || h <= 0 || nfeatures <= 0 || w > 4000 || h > 4000) { c.err = "bird_detect: bad argument"; return ORBB200_ERR_ARG; }
    BirdPlan* p = get_plan(c, w, h, nfeatures, 1);
    if (!p) return ORBB200_ERR_CUDA;
    int rc = upload_images(c, p, &img, mask ? &mask : nullptr, 1, stride, mask_stride);
    if (rc == ORBB200_OK) rc = enqueue_detect(c, p, 1, mask != nullptr ? 1 : 0);
    if (rc != ORBB200_OK) return rc;
    return download(c, p, 1, p->d_kps, p->d_counts, kps, nullptr, cap, n_out);
}

int orbb200_corner_subpix(orbb200_ctx* ctx, const uint8_t* img, int w, int h, size_t stride, float* pts, int n, int win_w, int win_h,
                          int max_iter, double eps)
{
    BIRD_ENTER(ctx);
    if (!img || n < 0 || (n > 0 && !pts) || w <= 0 || h <= 0 || w > 4000 || h > 4000) { c.err = "corner_subpix: bad argument"; return ORBB200_ERR_ARG; }
    if (n == 0) return ORBB200_OK;
    BirdPlan* p = get_plan(c, w, h, 2000, 1);
    if (!p) return ORBB200_ERR_CUDA;
    int rc = upload_images(c, p, &img, nullptr, 1, stride, 0);
    if (rc != ORBB200_OK) return rc;
    const int cap = p->g.kpPerImg;
    for (int o = 0; o < n; o += cap) {
        const int m = std::min(cap, n - o);
        ORBB200_CUDA_OK(c, cudaMemcpyAsync(p->d_pts, pts + 2 * (size_t)o, sizeof(float) * 2 * m, cudaMemcpyHostToDevice, c.stream));
        rc = enqueue_subpix(c, p, 1, nullptr, m, win_w, win_h, max_iter, eps);
        if (rc != ORBB200_OK) return rc;
        ORBB200_CUDA_OK(c, cudaMemcpyAsync(pts + 2 * (size_t)o, p->d_pts, sizeof(float) * 2 * m, cudaMemcpyDeviceToHost, c.stream));
        ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
    }
    return ORBB200_OK;
}

int orbb200_bird_compute(orbb200_ctx* ctx, const uint8_t* img, int w, int h, size_t stride, orbb200_kp_t* kps, int n, uint8_t* desc,
                         int* n_out)
{
    BIRD_ENTER(ctx);
    if (!img || n < 0 || !n_out || (n > 0 && (!kps || !desc)) || w <= 0 || h <= 0 || w > 4000 || h > 4000) { c.err = "bird_compute: bad argument"; return ORBB200_ERR_ARG; }
    *n_out = 0;
    if (n == 0) return ORBB200_OK;
    BirdPlan* p = get_plan(c, w, h, 2000, 1);
    if (!p) return ORBB200_ERR_CUDA;
    if (n > p->g.kpPerImg) { c.err = "bird_compute: too many keypoints for one call"; return ORBB200_ERR_ARG; }
    int nLevels = 0;
    for (int i = 0; i < n; i++) {
        if (kps[i].octave < 0 || kps[i].octave >= BV_LEVELS) { c.err = "bird_compute: keypoint octave out of range"; return ORBB200_ERR_ARG; }
        nLevels = std::max(nLevels, kps[i].octave);
    }
    nLevels++;
    int rc = upload_images(c, p, &img, nullptr, 1, stride, 0);
    if (rc != ORBB200_OK) return rc;
    const int32_t cnt = n;
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(p->d_kps, kps, sizeof(orbb200_kp_t) * n, cudaMemcpyHostToDevice, c.stream));
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(p->d_counts, &cnt, 4, cudaMemcpyHostToDevice, c.stream));
    enqueue_pyramid(c, p, 1, false, nLevels);
    rc = enqueue_compute(c, p, 1, nLevels);
    if (rc != ORBB200_OK) return rc;
    return download(c, p, 1, p->d_kps2, p->d_counts2, kps, desc, n, n_out);
}

int orbb200_bird_extract_batch(orbb200_ctx* ctx, const uint8_t* const* imgs, const uint8_t* const* masks, int n, int w, int h, size_t stride,
                               size_t mask_stride, int nfeatures, orbb200_kp_t* kps, uint8_t* desc, int cap_per_img, int* n_out)
{
    BIRD_ENTER(ctx);
    if (!imgs || n <= 0 || !kps || !desc || !n_out || w <= 0 || h <= 0 || nfeatures <= 0 || w > 4000 || h > 4000) { c.err = "bird_extract: bad argument"; return ORBB200_ERR_ARG; }
    BirdPlan* p = get_plan(c, w, h, nfeatures, n);
    if (!p) return ORBB200_ERR_CUDA;
    const BirdGeom& g = p->g;
    if (!c.hostCopies.empty()) { ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream)); deliver_host_copies(c); }   // staged results of an earlier frame step
    int rc = ORBB200_OK;
    // One image per call (the birdview block of Frame::Frame, src/Frame.cc:328-342): the image goes through the pinned staging block (a
    // host memcpy + a true DMA instead of the copy engine's synchronous bounce of pageable memory), the results come back through it
    // in one synchronisation, and the mask -- the vehicle mask, the same for every frame -- is compared with the copy kept from the last
    // call: when equal its upload and the seven resize + threshold launches of its pyramid are skipped.
    const BirdLevel& L0 = g.lv[0];
    const size_t imgB = (size_t)w * h, outB = 128 + (size_t)g.kpPerImg * 60;
    const bool staged = c.stageUploads && n == 1 && imgB * 2 <= STAGE_D2H_OFF - STAGE_H2D_OFF && outB <= STAGE_LIMIT - STAGE_D2H_OFF && ensure_scratch(c, 0, STAGE_LIMIT);
    bool maskCached = false;
    if (staged) {
        uint8_t* hs = c.h_scratch + STAGE_H2D_OFF;
        for (int y = 0; y < h; y++) memcpy(hs + (size_t)y * w, imgs[0] + (size_t)y * stride, (size_t)w);
        ORBB200_CUDA_OK(c, cudaMemcpy2DAsync(p->d_pyr + L0.off, L0.pitch, hs, (size_t)w, (size_t)w, (size_t)h, cudaMemcpyHostToDevice, c.stream));
        if (masks) {
            maskCached = p->maskCacheValid && p->hostMask.size() == imgB;
            for (int y = 0; maskCached && y < h; y++) maskCached = memcmp(p->hostMask.data() + (size_t)y * w, masks[0] + (size_t)y * mask_stride, (size_t)w) == 0;
            if (!maskCached) {
                p->hostMask.resize(imgB);
                uint8_t* hm = hs + imgB;
                for (int y = 0; y < h; y++) {
                    memcpy(hm + (size_t)y * w, masks[0] + (size_t)y * mask_stride, (size_t)w);
                    memcpy(p->hostMask.data() + (size_t)y * w, masks[0] + (size_t)y * mask_stride, (size_t)w);
                }
                ORBB200_CUDA_OK(c, cudaMemcpy2DAsync(p->d_mask + L0.off, L0.pitch, hm, (size_t)w, (size_t)w, (size_t)h, cudaMemcpyHostToDevice, c.stream));
                p->maskCacheValid = true;
            }
        }
    } else {
        rc = upload_images(c, p, imgs, masks, n, stride, mask_stride);
    }
    bool blurForked = false;
    if (rc == ORBB200_OK) rc = enqueue_detect(c, p, n, masks != nullptr ? 1 : 0, &blurForked, p->d_pts, maskCached);
    if (rc != ORBB200_OK) return rc;
    // cornerSubPix(img, pts, Size(5,5), Size(-1,-1), TermCriteria(EPS + MAX_ITER, 40, 0.001))   (src/Frame.cc:335-336)
    if (g.w >= 15 && g.h >= 15) {
        rc = enqueue_subpix(c, p, n, p->d_counts, 0, 5, 5, 40, 0.001);
        if (rc != ORBB200_OK) return rc;
    }
    rc = enqueue_compute(c, p, n, BV_LEVELS, blurForked, p->d_pts);
    if (rc != ORBB200_OK) return rc;
    if (staged) {
        uint8_t* ho = c.h_scratch + STAGE_D2H_OFF;
        int32_t* hStatus = reinterpret_cast<int32_t*>(ho);
        int32_t* hCnt = reinterpret_cast<int32_t*>(ho + 64);
        uint8_t* hK = ho + 128;
        uint8_t* hD = hK + (size_t)g.kpPerImg * sizeof(orbb200_kp_t);
        ORBB200_CUDA_OK(c, cudaMemcpyAsync(hStatus, c.d_status, sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
        ORBB200_CUDA_OK(c, cudaMemcpyAsync(hCnt, p->d_counts2, sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
        ORBB200_CUDA_OK(c, cudaMemcpyAsync(hK, p->d_kps2, (size_t)g.kpPerImg * sizeof(orbb200_kp_t), cudaMemcpyDeviceToHost, c.stream));
        ORBB200_CUDA_OK(c, cudaMemcpyAsync(hD, p->d_desc, (size_t)g.kpPerImg * 32, cudaMemcpyDeviceToHost, c.stream));
        ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
        if (*hStatus != 0) return check_bird_status(c);       // reads, reports and clears the flag
        const int m = *hCnt;
        if (m > cap_per_img) { c.err = "bird: output capacity too small"; return ORBB200_ERR_CAPACITY; }
        memcpy(kps, hK, (size_t)std::max(m, 0) * sizeof(orbb200_kp_t));
        memcpy(desc, hD, (size_t)std::max(m, 0) * 32);
        n_out[0] = m;
        return ORBB200_OK;
    }
    return download(c, p, n, p->d_kps2, p->d_counts2, kps, desc, cap_per_img, n_out);
}

int orbb200_bird_results_device(orbb200_ctx* ctx, int w, int h, int nfeatures, const orbb200_kp_t** d_kps, const uint8_t** d_desc,
                                const int32_t** d_counts, int* cap_per_img)
{
    BIRD_ENTER(ctx);
    BirdState& S = state(c);
    auto it = S.plans.find(std::make_tuple(w, h, nfeatures));
    if (it == S.plans.end()) { c.err = "bird_results_device: no birdview extraction of this size yet"; return ORBB200_ERR_ARG; }
    BirdPlan* p = it->second;
    if (d_kps) *d_kps = p->d_kps2;
    if (d_desc) *d_desc = p->d_desc;
    if (d_counts) *d_counts = p->d_counts2;
    if (cap_per_img) *cap_per_img = p->g.kpPerImg;
    return ORBB200_OK;
}

int orbb200_bird_extract(orbb200_ctx* ctx, const uint8_t* img, const uint8_t* mask, int w, int h, size_t stride, size_t mask_stride,
                         int nfeatures, orbb200_kp_t* kps, uint8_t* desc, int cap, int* n_out)
{
    return orbb200_bird_extract_batch(ctx, &img, mask ? &mask : nullptr, 1, w, h, stride, mask_stride, nfeatures, kps, desc, cap, n_out);
}

}  // extern "C"
