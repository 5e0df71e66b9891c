/* ORACLE -- TEST INFRASTRUCTURE ONLY (see orb_oracle.h).
 *
 * The birdview front-end of the reference (src/Frame.cc:328-342):
 *     cv::Ptr<cv::ORB> extractorBird = cv::ORB::create(2000);
 *     extractorBird->detect(mBirdviewImg, mvKeysBird, mBirdviewMask);
 *     cv::cornerSubPix(mBirdviewImg, pts, Size(5,5), Size(-1,-1), TermCriteria(EPS+MAX_ITER, 40, 0.001));
 *     extractorBird->compute(mBirdviewImg, mvKeysBird, mDescriptorsBird);
 * cv::ORB and cv::cornerSubPix live in OpenCV, which is not under /root/reference (find_package(OpenCV 3.0), version
 * not pinned).  This file restates OpenCV 4.x's algorithms (modules/features2d/src/orb.cpp, keypoint.cpp,
 * modules/imgproc/src/resize.cpp [INTER_LINEAR_EXACT], cornersubpix.cpp, samplers.cpp) and is pinned against the only
 * runnable OpenCV, cv2 4.13: tests/test_oracle.py compares detect / cornerSubPix / compute with live cv2 calls and with
 * the vectors committed in tests/golden/bird_orb_*.npz (made by tests/golden/make_golden_bird.py), element order
 * included.  KeyPointsFilter::retainBest's order comes from libstdc++'s std::nth_element / std::partition, written out
 * here so that the order is a defined part of the contract.
 */
#include "orb_oracle.h"
#include "../include/orbb200_pattern.inc"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

const signed char BPAT_X[512] = {ORBB200_PATTERN_X_INIT};
const signed char BPAT_Y[512] = {ORBB200_PATTERN_Y_INIT};

inline int cvRoundD(double v) { return (int)std::nearbyint(v); }
inline int cvRoundF(float v) { return (int)std::nearbyintf(v); }
inline int cvFloorF(float v) { int i = (int)v; return i - (i > v); }
inline int cvFloorD(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeilD(double v) { int i = (int)v; return i + (i < v); }

// ---- cv::resize(..., INTER_LINEAR_EXACT) on CV_8UC1 (resize.cpp, resize_bitExact): 8.8 fixed-point coefficients
//      from (double) coordinates, horizontal pass in 8.8, vertical pass rounded from 16.16 ----
struct ExactCoef { int ofs; int c0, c1; };

void exact_coeffs(int ssize, int dsize, std::vector<ExactCoef>& out)
{
    out.resize(dsize);
    const double scale = 1.0 / ((double)dsize / ssize);
    for (int v = 0; v < dsize; v++) {
        const double fval = scale * ((double)v + 0.5) - 0.5;
        const int ival = cvFloorD(fval);
        ExactCoef c;
        if (ival >= 0 && ssize > 1) {
            if (ival < ssize - 1) {
                c.ofs = ival;
                c.c1 = cvRoundD((fval - (double)ival) * 256.0);
                c.c0 = 256 - c.c1;
            } else {
                c.ofs = ssize - 1; c.c0 = 256; c.c1 = 0;
            }
        } else {
            c.ofs = 0; c.c0 = 256; c.c1 = 0;
        }
        out[v] = c;
    }
}

void resize_linear_exact(const uint8_t* src, int sw, int sh, size_t sstep, uint8_t* dst, int dw, int dh, size_t dstep)
{
    std::vector<ExactCoef> cx, cy;
    exact_coeffs(sw, dw, cx);
    exact_coeffs(sh, dh, cy);
    std::vector<uint16_t> H((size_t)sh * dw);
    for (int y = 0; y < sh; y++) {
        const uint8_t* s = src + y * sstep;
        for (int x = 0; x < dw; x++) {
            const int x0 = cx[x].ofs, x1 = std::min(x0 + 1, sw - 1);
            H[(size_t)y * dw + x] = (uint16_t)(s[x0] * cx[x].c0 + s[x1] * cx[x].c1);
        }
    }
    for (int y = 0; y < dh; y++) {
        const int y0 = cy[y].ofs, y1 = std::min(y0 + 1, sh - 1);
        for (int x = 0; x < dw; x++) {
            const uint32_t v = (uint32_t)H[(size_t)y0 * dw + x] * cy[y].c0 + (uint32_t)H[(size_t)y1 * dw + x] * cy[y].c1;
            const uint32_t r = (v + 32768u) >> 16;
            dst[y * dstep + x] = (uint8_t)std::min(r, 255u);
        }
    }
}

inline int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// ---- the packed pyramid of ORB_Impl::detectAndCompute (orb.cpp) ----
struct Layer { int x, y, w, h; };
struct Pyramid {
    int border = 0, bufW = 0, bufH = 0;
    std::vector<Layer> layers;
    std::vector<float> scale;
    std::vector<uint8_t> img, mask;
    bool hasMask = false;
    const uint8_t* at(int level, int x, int y) const { return &img[(size_t)(layers[level].y + y) * bufW + layers[level].x + x]; }
};

const int EDGE_THRESHOLD = 31, PATCH_SIZE = 31, HALF_PATCH = 15, FAST_THRESHOLD = 20, HARRIS_BLOCK_SIZE = 9;
const double SCALE_FACTOR = (double)1.2f;          // ORB::create's float default stored in a double member

float get_scale(int level) { return (float)std::pow(SCALE_FACTOR, (double)level); }

void build_pyramid(Pyramid& P, const uint8_t* image, const uint8_t* mask, int w, int h, size_t step, size_t mstep, int nLevels)
{
    const int descPatchSize = cvCeilD(HALF_PATCH * std::sqrt(2.0));
    const int border = std::max(EDGE_THRESHOLD, std::max(descPatchSize, HARRIS_BLOCK_SIZE / 2)) + 1;
    P.border = border;
    const float level0_inv_scale = 1.0f / get_scale(0);
    const int l0w = cvRoundF(w * level0_inv_scale), l0h = cvRoundF(h * level0_inv_scale);
    P.bufW = (l0w + border * 2 + 15) & ~15;
    int level_dy = l0h + border * 2, ox = 0, oy = 0;
    P.layers.resize(nLevels); P.scale.resize(nLevels);
    for (int level = 0; level < nLevels; level++) {
        const float scale = get_scale(level);
        P.scale[level] = scale;
        const float inv_scale = 1.0f / scale;
        const int sw = cvRoundF(w * inv_scale), sh = cvRoundF(h * inv_scale);
        const int ww = sw + border * 2, wh = sh + border * 2;
        if (ox + ww > P.bufW) { ox = 0; oy += level_dy; level_dy = wh; }
        P.layers[level] = {ox + border, oy + border, sw, sh};
        ox += ww;
    }
    P.bufH = oy + level_dy;
    P.img.assign((size_t)P.bufW * P.bufH, 0);
    P.hasMask = mask != nullptr;
    if (P.hasMask) P.mask.assign((size_t)P.bufW * P.bufH, 0);

    std::vector<uint8_t> prev, prevM;                       // previous level (tight copies)
    int pw = w, ph = h;
    prev.resize((size_t)w * h);
    for (int y = 0; y < h; y++) memcpy(&prev[(size_t)y * w], image + y * step, w);
    if (mask) { prevM.resize((size_t)w * h); for (int y = 0; y < h; y++) memcpy(&prevM[(size_t)y * w], mask + y * mstep, w); }
    const std::vector<uint8_t> image0 = prev, mask0 = prevM;
    for (int level = 0; level < nLevels; level++) {
        const Layer L = P.layers[level];
        std::vector<uint8_t> cur, curM;
        if (level != 0) {
            cur.resize((size_t)L.w * L.h);
            resize_linear_exact(prev.data(), pw, ph, pw, cur.data(), L.w, L.h, L.w);
            if (mask) {
                curM.resize((size_t)L.w * L.h);
                resize_linear_exact(prevM.data(), pw, ph, pw, curM.data(), L.w, L.h, L.w);
                for (auto& m : curM) m = m > 254 ? m : 0;                  // threshold(254, THRESH_TOZERO)
            }
        } else {
            cur = image0; curM = mask0;
        }
        // copyMakeBorder: image REFLECT_101, mask CONSTANT 0
        for (int y = -border; y < L.h + border; y++) {
            uint8_t* d = &P.img[(size_t)(L.y + y) * P.bufW + L.x];
            const uint8_t* s = &cur[(size_t)reflect101(y, L.h) * L.w];
            for (int x = -border; x < L.w + border; x++) d[x] = s[reflect101(x, L.w)];
        }
        if (mask)
            for (int y = 0; y < L.h; y++) memcpy(&P.mask[(size_t)(L.y + y) * P.bufW + L.x], &curM[(size_t)y * L.w], L.w);
        if (level > 0) { prev.swap(cur); prevM.swap(curM); pw = L.w; ph = L.h; }
        // (level 0: prev stays the input image, so level 1 is resized from it)
    }
}

// ---- KeyPointsFilter::retainBest (keypoint.cpp) with libstdc++'s algorithms written out ----
typedef oracle_kp_t KP;
inline bool greaterResp(const KP& a, const KP& b) { return a.response > b.response; }

void move_median_to_first(KP* v, int result, int a, int b, int c)
{
    if (greaterResp(v[a], v[b])) {
        if (greaterResp(v[b], v[c])) std::swap(v[result], v[b]);
        else if (greaterResp(v[a], v[c])) std::swap(v[result], v[c]);
        else std::swap(v[result], v[a]);
    } else if (greaterResp(v[a], v[c])) std::swap(v[result], v[a]);
    else if (greaterResp(v[b], v[c])) std::swap(v[result], v[c]);
    else std::swap(v[result], v[b]);
}

int unguarded_partition(KP* v, int first, int last, int pivot)
{
    while (true) {
        while (greaterResp(v[first], v[pivot])) ++first;
        --last;
        while (greaterResp(v[pivot], v[last])) --last;
        if (!(first < last)) return first;
        std::swap(v[first], v[last]);
        ++first;
    }
}

void insertion_sort(KP* v, int first, int last)
{
    if (first == last) return;
    for (int i = first + 1; i < last; i++) {
        const KP val = v[i];
        if (greaterResp(val, v[first])) {
            for (int j = i; j > first; j--) v[j] = v[j - 1];
            v[first] = val;
        } else {
            int j = i;
            while (greaterResp(val, v[j - 1])) { v[j] = v[j - 1]; j--; }
            v[j] = val;
        }
    }
}

// std::__heap_select + iter_swap: the depth-limit fallback of std::__introselect
void adjust_heap(KP* first, int holeIndex, int len, KP value)
{
    const int topIndex = holeIndex;
    int secondChild = holeIndex;
    while (secondChild < (len - 1) / 2) {
        secondChild = 2 * (secondChild + 1);
        if (greaterResp(first[secondChild], first[secondChild - 1])) secondChild--;
        first[holeIndex] = first[secondChild];
        holeIndex = secondChild;
    }
    if ((len & 1) == 0 && secondChild == (len - 2) / 2) {
        secondChild = 2 * (secondChild + 1);
        first[holeIndex] = first[secondChild - 1];
        holeIndex = secondChild - 1;
    }
    int parent = (holeIndex - 1) / 2;                                   // std::__push_heap
    while (holeIndex > topIndex && greaterResp(first[parent], value)) {
        first[holeIndex] = first[parent];
        holeIndex = parent;
        parent = (holeIndex - 1) / 2;
    }
    first[holeIndex] = value;
}

void heap_select(KP* v, int first, int middle, int last)
{
    KP* f = v + first;
    const int len = middle - first;
    if (len >= 2)                                                        // std::__make_heap
        for (int parent = (len - 2) / 2;; parent--) {
            const KP value = f[parent];
            adjust_heap(f, parent, len, value);
            if (parent == 0) break;
        }
    for (int i = middle; i < last; i++)
        if (greaterResp(v[i], v[first])) {                               // std::__pop_heap(first, middle, i)
            const KP value = v[i];
            v[i] = v[first];
            adjust_heap(f, 0, len, value);
        }
}

void nth_element_resp(KP* v, int n, int nth)
{
    int first = 0, last = n;
    if (first == last || nth == last) return;
    int depth = 0;
    for (int k = n; k > 1; k >>= 1) depth++;
    depth *= 2;
    while (last - first > 3) {
        if (depth == 0) {
            heap_select(v, first, nth + 1, last);
            std::swap(v[first], v[nth]);
            return;
        }
        --depth;
        const int mid = first + (last - first) / 2;
        move_median_to_first(v, first, first + 1, mid, last - 1);
        const int cut = unguarded_partition(v, first + 1, last, first);
        if (cut <= nth) first = cut; else last = cut;
    }
    insertion_sort(v, first, last);
}

int partition_ge(KP* v, int first, int last, float thr)
{
    while (true) {
        while (true) {
            if (first == last) return first;
            if (v[first].response >= thr) ++first; else break;
        }
        --last;
        while (true) {
            if (first == last) return first;
            if (!(v[last].response >= thr)) --last; else break;
        }
        std::swap(v[first], v[last]);
        ++first;
    }
}

void retain_best(std::vector<KP>& k, int n_points)
{
    if (n_points >= 0 && (int)k.size() > n_points) {
        if (n_points == 0) { k.clear(); return; }
        nth_element_resp(k.data(), (int)k.size(), n_points - 1);
        const float ambiguous = k[n_points - 1].response;
        const int end = partition_ge(k.data(), n_points, (int)k.size(), ambiguous);
        k.resize(end);
    }
}

void features_per_level(int nfeatures, int nlevels, std::vector<int>& out)
{
    out.resize(nlevels);
    const float factor = (float)(1.0 / SCALE_FACTOR);
    float nd = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        out[l] = cvRoundF(nd);
        sum += out[l];
        nd *= factor;
    }
    out[nlevels - 1] = std::max(nfeatures - sum, 0);
}

void umax_table(int (&umax)[HALF_PATCH + 2])
{
    int v, v0;
    const int vmax = cvFloorF(HALF_PATCH * std::sqrt(2.f) / 2 + 1);
    const int vmin = cvCeilD(HALF_PATCH * std::sqrt(2.f) / 2);
    for (v = 0; v <= HALF_PATCH + 1; v++) umax[v] = 0;
    for (v = 0; v <= vmax; ++v) umax[v] = cvRoundD(std::sqrt((double)HALF_PATCH * HALF_PATCH - v * v));
    for (v = HALF_PATCH, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
    }
}

void harris_responses(const Pyramid& P, std::vector<KP>& pts)
{
    const int blockSize = 7, r = blockSize / 2, step = P.bufW;
    const float harris_k = 0.04f;
    const float scale = 1.f / ((1 << 2) * blockSize * 255.f);
    const float scale_sq_sq = scale * scale * scale * scale;
    for (auto& kp : pts) {
        const int x0 = cvRoundF(kp.x), y0 = cvRoundF(kp.y), z = kp.octave;
        const uint8_t* ptr0 = &P.img[(size_t)(y0 - r + P.layers[z].y) * step + (x0 - r + P.layers[z].x)];
        int a = 0, b = 0, c = 0;
        for (int i = 0; i < blockSize; i++)
            for (int j = 0; j < blockSize; j++) {
                const uint8_t* ptr = ptr0 + i * step + j;
                const int Ix = (ptr[1] - ptr[-1]) * 2 + (ptr[-step + 1] - ptr[-step - 1]) + (ptr[step + 1] - ptr[step - 1]);
                const int Iy = (ptr[step] - ptr[-step]) * 2 + (ptr[step - 1] - ptr[-step - 1]) + (ptr[step + 1] - ptr[-step + 1]);
                a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
            }
        kp.response = ((float)a * b - (float)c * c - harris_k * ((float)a + b) * ((float)a + b)) * scale_sq_sq;
    }
}

void ic_angles(const Pyramid& P, std::vector<KP>& pts, const int* u_max)
{
    const int step = P.bufW;
    for (auto& kp : pts) {
        const Layer& L = P.layers[kp.octave];
        const uint8_t* center = &P.img[(size_t)(cvRoundF(kp.y) + L.y) * step + cvRoundF(kp.x) + L.x];
        int m_01 = 0, m_10 = 0;
        for (int u = -HALF_PATCH; u <= HALF_PATCH; ++u) m_10 += u * center[u];
        for (int v = 1; v <= HALF_PATCH; ++v) {
            int v_sum = 0;
            const int d = u_max[v];
            for (int u = -d; u <= d; ++u) {
                const int val_plus = center[u + v * step], val_minus = center[u - v * step];
                v_sum += (val_plus - val_minus);
                m_10 += u * (val_plus + val_minus);
            }
            m_01 += v * v_sum;
        }
        kp.angle = oracle_fast_atan2((float)m_01, (float)m_10);
    }
}

// KeyPointsFilter::runByImageBorder: Rect(Point(b,b), Point(w-b,h-b)).contains(Point(kp.pt)) with Point2f -> Point rounding
void run_by_image_border(std::vector<KP>& k, int w, int h, int borderSize)
{
    if (borderSize <= 0) return;
    if (h <= borderSize * 2 || w <= borderSize * 2) { k.clear(); return; }
    size_t o = 0;
    for (size_t i = 0; i < k.size(); i++) {
        const int x = cvRoundF(k[i].x), y = cvRoundF(k[i].y);
        if (x >= borderSize && x < w - borderSize && y >= borderSize && y < h - borderSize) k[o++] = k[i];
    }
    k.resize(o);
}

void detect(const uint8_t* image, const uint8_t* mask, int w, int h, size_t step, size_t mstep, int nfeatures, int nlevels,
            std::vector<KP>& all)
{
    Pyramid P;
    build_pyramid(P, image, mask, w, h, step, mstep, nlevels);
    std::vector<int> quota;
    features_per_level(nfeatures, nlevels, quota);
    int umax[HALF_PATCH + 2];
    umax_table(umax);
    all.clear();
    std::vector<int> counters(nlevels);
    std::vector<int32_t> xyr;
    for (int level = 0; level < nlevels; level++) {
        const Layer L = P.layers[level];
        const uint8_t* img = P.at(level, 0, 0);
        xyr.resize((size_t)3 * std::max(L.w * L.h, 1));
        const int n = oracle_fast9(img, L.w, L.h, P.bufW, FAST_THRESHOLD, 1, xyr.data(), L.w * L.h);
        std::vector<KP> kps;
        kps.reserve(n);
        for (int i = 0; i < n; i++) {
            KP kp{(float)xyr[3 * i], (float)xyr[3 * i + 1], 7.f, -1.f, (float)xyr[3 * i + 2], 0, -1};
            if (P.hasMask) {                                             // KeyPointsFilter::runByPixelsMask
                const int my = (int)(kp.y + 0.5f), mx = (int)(kp.x + 0.5f);
                if (P.mask[(size_t)(L.y + my) * P.bufW + L.x + mx] == 0) continue;
            }
            kps.push_back(kp);
        }
        run_by_image_border(kps, L.w, L.h, EDGE_THRESHOLD);
        retain_best(kps, 2 * quota[level]);
        const float sf = P.scale[level];
        for (auto& kp : kps) { kp.octave = level; kp.size = PATCH_SIZE * sf; }
        counters[level] = (int)kps.size();
        all.insert(all.end(), kps.begin(), kps.end());
    }
    if (all.empty()) return;
    harris_responses(P, all);
    std::vector<KP> out;
    size_t off = 0;
    for (int level = 0; level < nlevels; level++) {
        std::vector<KP> kps(all.begin() + off, all.begin() + off + counters[level]);
        off += counters[level];
        retain_best(kps, quota[level]);
        out.insert(out.end(), kps.begin(), kps.end());
    }
    all.swap(out);
    ic_angles(P, all, umax);
    for (auto& kp : all) {
        const float scale = P.scale[kp.octave];
        kp.x *= scale; kp.y *= scale;
    }
}

// ---- cv::getRectSubPix CV_8U -> CV_32F (samplers.cpp: getRectSubPix_Cn_ with adjustRect).  cv2 4.13 evaluates the
//      four-tap sum as (top pair) + (bottom pair) -- established against cv2.getRectSubPix, tests/test_oracle.py ----
struct IRect { int x, y, width, height; };

const uint8_t* adjust_rect(const uint8_t* src, ptrdiff_t src_step, int src_w, int src_h, int win_w, int win_h, int ipx, int ipy, IRect* pRect)
{
    IRect rect;
    if (ipx >= 0) { src += ipx; rect.x = 0; }
    else { rect.x = -ipx; if (rect.x > win_w) rect.x = win_w; }
    if (ipx < src_w - win_w) rect.width = win_w;
    else {
        rect.width = src_w - ipx - 1;
        if (rect.width < 0) { src += rect.width; rect.width = 0; }
    }
    if (ipy >= 0) { src += ipy * src_step; rect.y = 0; }
    else rect.y = -ipy;
    if (ipy < src_h - win_h) rect.height = win_h;
    else {
        rect.height = src_h - ipy - 1;
        if (rect.height < 0) { src += rect.height * src_step; rect.height = 0; }
    }
    *pRect = rect;
    return src - rect.x;
}

void get_rect_sub_pix_8u32f(const uint8_t* src, ptrdiff_t src_step, int src_w, int src_h, float* dst, int dst_step, int win_w, int win_h,
                            float cx, float cy)
{
    float centerx = cx, centery = cy;
    centerx -= (win_w - 1) * 0.5f;
    centery -= (win_h - 1) * 0.5f;
    const int ipx = cvFloorF(centerx), ipy = cvFloorF(centery);
    if (0 <= ipx && ipx + win_w < src_w && 0 <= ipy && ipy + win_h < src_h && win_w > 0 && win_h > 0) {
        // getRectSubPix_8u32f: window inside the image
        float a = centerx - ipx;
        const float b = centery - ipy;
        a = std::max(a, 0.0001f);
        const float a12 = a * (1.f - b);
        const float a22 = a * b;
        const float b1 = 1.f - b;
        const float b2 = b;
        const double s = (1. - a) / a;
        src += ipy * src_step + ipx;
        for (int i = 0; i < win_h; i++, src += src_step, dst += dst_step) {
            float prev = (1 - a) * (b1 * src[0] + b2 * src[src_step]);
            for (int j = 0; j < win_w; j++) {
                const float t = a12 * src[j + 1] + a22 * src[j + 1 + src_step];
                dst[j] = prev + t;
                prev = (float)(t * s);
            }
        }
        return;
    }
    // getRectSubPix_Cn_<uchar, float, float, nop, nop>
    const float a = centerx - ipx, b = centery - ipy;
    const float a11 = (1.f - a) * (1.f - b), a12 = a * (1.f - b), a21 = (1.f - a) * b, a22 = a * b;
    const float b1 = 1.f - b, b2 = b;
    if (0 <= ipx && ipx < src_w - win_w && 0 <= ipy && ipy < src_h - win_h) {
        src += ipy * src_step + ipx;
        for (int i = 0; i < win_h; i++, src += src_step, dst += dst_step)
            for (int j = 0; j < win_w; j++)
                dst[j] = src[j] * a11 + src[j + 1] * a12 + src[j + src_step] * a21 + src[j + src_step + 1] * a22;
        return;
    }
    IRect r;
    src = adjust_rect(src, src_step, src_w, src_h, win_w, win_h, ipx, ipy, &r);
    for (int i = 0; i < win_h; i++, dst += dst_step) {
        const uint8_t* src2 = src + src_step;
        if (i < r.y || i >= r.height) src2 -= src_step;
        float s0 = src[r.x] * b1 + src2[r.x] * b2;
        for (int j = 0; j < r.x; j++) dst[j] = s0;
        s0 = src[r.width] * b1 + src2[r.width] * b2;
        for (int j = r.width; j < win_w; j++) dst[j] = s0;
        for (int j = r.x; j < r.width; j++)
            dst[j] = src[j] * a11 + src[j + 1] * a12 + src2[j] * a21 + src2[j + 1] * a22;
        if (i < r.height) src = src2;
    }
}

// ---- cv::cornerSubPix (cornersubpix.cpp), zeroZone (-1,-1) ----
void corner_subpix(const uint8_t* src, int cols, int rows, size_t step, float* corners, int count, int winW, int winH, int maxCount,
                   double epsilon)
{
    const int MAX_ITERS = 100;
    const int win_w = winW * 2 + 1, win_h = winH * 2 + 1;
    const int max_iters = std::min(std::max(maxCount, 1), MAX_ITERS);
    double eps = std::max(epsilon, 0.);
    eps *= eps;
    std::vector<float> mask((size_t)win_w * win_h), subpix_buf((size_t)(win_w + 2) * (win_h + 2));
    for (int i = 0; i < win_h; i++) {
        const float y = (float)(i - winH) / winH;
        const float vy = std::exp(-y * y);
        for (int j = 0; j < win_w; j++) {
            const float x = (float)(j - winW) / winW;
            mask[i * win_w + j] = (float)(vy * std::exp(-x * x));
        }
    }
    for (int pt_i = 0; pt_i < count; pt_i++) {
        const float cTx = corners[2 * pt_i], cTy = corners[2 * pt_i + 1];
        float cIx = cTx, cIy = cTy;
        int iter = 0;
        double err = 0;
        do {
            double a = 0, b = 0, c = 0, bb1 = 0, bb2 = 0;
            get_rect_sub_pix_8u32f(src, (ptrdiff_t)step, cols, rows, subpix_buf.data(), win_w + 2, win_w + 2, win_h + 2, cIx, cIy);
            const float* subpix = &subpix_buf[(win_w + 2) + 1];
            for (int i = 0, k = 0; i < win_h; i++, subpix += win_w + 2) {
                const double py = i - winH;
                for (int j = 0; j < win_w; j++, k++) {
                    const double m = mask[k];
                    const double tgx = subpix[j + 1] - subpix[j - 1];
                    const double tgy = subpix[j + win_w + 2] - subpix[j - win_w - 2];
                    const double gxx = tgx * tgx * m;
                    const double gxy = tgx * tgy * m;
                    const double gyy = tgy * tgy * m;
                    const double px = j - winW;
                    a += gxx; b += gxy; c += gyy;
                    bb1 += gxx * px + gxy * py;
                    bb2 += gxy * px + gyy * py;
                }
            }
            const double det = a * c - b * b;
            if (std::fabs(det) <= DBL_EPSILON * DBL_EPSILON) break;
            const double scale = 1.0 / det;
            const float nx = (float)(cIx + c * scale * bb1 - b * scale * bb2);
            const float ny = (float)(cIy - b * scale * bb1 + a * scale * bb2);
            err = (nx - cIx) * (nx - cIx) + (ny - cIy) * (ny - cIy);
            cIx = nx; cIy = ny;
            if (cIx < 0 || cIx >= cols || cIy < 0 || cIy >= rows) break;
        } while (++iter < max_iters && err > eps);
        if (std::fabs(cIx - cTx) > winW || std::fabs(cIy - cTy) > winH) { cIx = cTx; cIy = cTy; }
        corners[2 * pt_i] = cIx; corners[2 * pt_i + 1] = cIy;
    }
}

// cv::getGaussianKernel(7, 2.0, CV_32F): exp(-x^2 / (2 sigma^2)) evaluated in double, stored as float, normalised by
// the double sum of the stored floats (smooth.dispatch.cpp, getGaussianKernelBitExact is only used for the fixed-point path)
void gaussian_kernel7(float (&k)[7])
{
    const int n = 7;
    const double sigmaX = 2.0, scale2X = -0.5 / (sigmaX * sigmaX);
    double sum = 0;
    for (int i = 0; i < n; i++) {
        const double x = i - (n - 1) * 0.5;
        const double t = std::exp(scale2X * x * x);
        k[i] = (float)t;
        sum += k[i];
    }
    sum = 1. / sum;
    for (int i = 0; i < n; i++) k[i] = (float)(k[i] * sum);
}

// sepFilter2D(8U -> 8U) with the float 7-tap Gaussian on a w x h region whose 3-pixel surroundings are addressable
// (in place allowed: rows are filtered horizontally into a float buffer first)
void sep_gauss7_f32_u8(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep)
{
    float gk[7];
    gaussian_kernel7(gk);
    std::vector<float> H((size_t)(h + 6) * w);
    for (int y = -3; y < h + 3; y++) {
        const uint8_t* s = src + (ptrdiff_t)y * (ptrdiff_t)sstep;
        float* hr = &H[(size_t)(y + 3) * w];
        for (int x = 0; x < w; x++) {
            float acc = gk[0] * s[x - 3];
            for (int k = 1; k < 7; k++) acc += gk[k] * s[x - 3 + k];
            hr[x] = acc;
        }
    }
    for (int y = 0; y < h; y++) {
        uint8_t* d = dst + (size_t)y * dstep;
        for (int x = 0; x < w; x++) {
            float acc = gk[3] * H[(size_t)(y + 3) * w + x];
            for (int j = 1; j <= 3; j++) acc += gk[3 + j] * (H[(size_t)(y + 3 + j) * w + x] + H[(size_t)(y + 3 - j) * w + x]);
            const int v = cvRoundF(acc);
            d[x] = (uint8_t)std::min(std::max(v, 0), 255);
        }
    }
}

// ---- ORB::compute on provided keypoints (detectAndCompute, useProvidedKeypoints = true) ----
int compute(const uint8_t* image, int w, int h, size_t step, std::vector<KP>& kps, std::vector<uint8_t>& desc)
{
    int nLevels = 0;
    bool sortedByLevel = true;
    for (size_t i = 0; i < kps.size(); i++) {
        const int level = kps[i].octave;
        if (i > 0 && level < kps[i - 1].octave) sortedByLevel = false;
        nLevels = std::max(nLevels, level);
    }
    nLevels++;
    Pyramid P;
    build_pyramid(P, image, nullptr, w, h, step, 0, nLevels);
    run_by_image_border(kps, w, h, EDGE_THRESHOLD);
    if (!sortedByLevel) {
        std::vector<std::vector<KP>> byLevel(nLevels);
        for (auto& kp : kps) byLevel[kp.octave].push_back(kp);
        kps.clear();
        for (auto& v : byLevel) kps.insert(kps.end(), v.begin(), v.end());
    }
    desc.assign(kps.size() * 32, 0);
    if (kps.empty()) return 0;
    // GaussianBlur(workingMat, workingMat, Size(7,7), 2, 2, BORDER_REFLECT_101) on a SUB-MATRIX without BORDER_ISOLATED:
    // OpenCV's bit-exact 8-bit Gaussian is skipped for sub-matrices and the call falls through to
    // sepFilter2D(8U -> 8U) with the float kernel getGaussianKernel(7, 2, CV_32F) (smooth.dispatch.cpp).  Its arithmetic,
    // established against cv2.sepFilter2D: rows = float sum of k[i]*p[i] left to right; columns = k[3]*H[y] then
    // += k[3+j]*(H[y+j] + H[y-j]) for j = 1..3; result cvRound-ed and saturated.  The pixels next to the level come from
    // the buffer = the reflect-101 margin.
    for (int level = 0; level < nLevels; level++) {
        const Layer L = P.layers[level];
        uint8_t* base = &P.img[(size_t)L.y * P.bufW + L.x];
        sep_gauss7_f32_u8(base, L.w, L.h, P.bufW, base, P.bufW);
    }
    const int stepP = P.bufW;
    for (size_t j = 0; j < kps.size(); j++) {
        const KP& kpt = kps[j];
        const Layer& L = P.layers[kpt.octave];
        const float scale = 1.f / P.scale[kpt.octave];
        float angle = kpt.angle;
        angle *= (float)(3.14159265358979323846 / 180.f);
        const float a = std::cos(angle), b = std::sin(angle);          // cos(float) -> cosf
        const uint8_t* center = &P.img[(size_t)(cvRoundF(kpt.y * scale) + L.y) * stepP + cvRoundF(kpt.x * scale) + L.x];
        uint8_t* d = &desc[j * 32];
        for (int i = 0; i < 32; i++) {
            int val = 0;
            for (int k = 0; k < 8; k++) {
                int t[2];
                for (int e = 0; e < 2; e++) {
                    const int idx = 16 * i + 2 * k + e;
                    const float x = BPAT_X[idx] * a - BPAT_Y[idx] * b;
                    const float y = BPAT_X[idx] * b + BPAT_Y[idx] * a;
                    t[e] = center[cvRoundF(y) * stepP + cvRoundF(x)];
                }
                val |= (t[0] < t[1]) << k;
            }
            d[i] = (uint8_t)val;
        }
    }
    return (int)kps.size();
}

}  // namespace

extern "C" {

/* sepFilter2D(src, CV_8U, g, g, BORDER_REFLECT_101) with g = getGaussianKernel(7, 2, CV_32F) on a whole image */
void oracle_sep_gauss7_f32_u8(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep)
{
    std::vector<uint8_t> pad((size_t)(w + 6) * (h + 6));
    for (int y = -3; y < h + 3; y++)
        for (int x = -3; x < w + 3; x++) pad[(size_t)(y + 3) * (w + 6) + x + 3] = src[(size_t)reflect101(y, h) * sstep + reflect101(x, w)];
    sep_gauss7_f32_u8(&pad[(size_t)3 * (w + 6) + 3], w, h, w + 6, dst, dstep);
}

void oracle_resize_linear_exact_u8(const uint8_t* src, int sw, int sh, size_t sstep, uint8_t* dst, int dw, int dh, size_t dstep)
{
    resize_linear_exact(src, sw, sh, sstep, dst, dw, dh, dstep);
}

/* cv::ORB::create(nfeatures)->detect(image, keypoints, mask); mask may be NULL.  Returns the count (<= cap written). */
int oracle_bird_detect(const uint8_t* image, const uint8_t* mask, int w, int h, size_t step, size_t mstep, int nfeatures,
                       oracle_kp_t* out, int cap)
{
    std::vector<KP> all;
    detect(image, mask, w, h, step, mstep, nfeatures, 8, all);
    for (int i = 0; i < (int)all.size() && i < cap; i++) out[i] = all[i];
    return (int)all.size();
}

void oracle_get_rect_sub_pix_8u32f(const uint8_t* src, int w, int h, size_t step, float cx, float cy, int win_w, int win_h, float* dst)
{
    get_rect_sub_pix_8u32f(src, (ptrdiff_t)step, w, h, dst, win_w, win_w, win_h, cx, cy);
}

void oracle_corner_subpix(const uint8_t* image, int w, int h, size_t step, float* pts, int n, int win_w, int win_h, int max_iter,
                          double eps)
{
    corner_subpix(image, w, h, step, pts, n, win_w, win_h, max_iter, eps);
}

/* cv::ORB::compute(image, keypoints, descriptors): kps is filtered in place (runByImageBorder) -> returns the new count. */
int oracle_bird_compute(const uint8_t* image, int w, int h, size_t step, oracle_kp_t* kps, int n, uint8_t* desc)
{
    std::vector<KP> v(kps, kps + n);
    std::vector<uint8_t> d;
    const int m = compute(image, w, h, step, v, d);
    for (int i = 0; i < m; i++) kps[i] = v[i];
    if (m > 0) memcpy(desc, d.data(), (size_t)m * 32);
    return m;
}

/* The whole birdview front-end of src/Frame.cc:328-342. */
int oracle_bird_extract(const uint8_t* image, const uint8_t* mask, int w, int h, size_t step, size_t mstep, int nfeatures,
                        oracle_kp_t* kps, uint8_t* desc, int cap)
{
    std::vector<KP> all;
    detect(image, mask, w, h, step, mstep, nfeatures, 8, all);
    std::vector<float> pts(2 * all.size());
    for (size_t i = 0; i < all.size(); i++) { pts[2 * i] = all[i].x; pts[2 * i + 1] = all[i].y; }
    if (!all.empty()) corner_subpix(image, w, h, step, pts.data(), (int)all.size(), 5, 5, 40, 0.001);
    for (size_t i = 0; i < all.size(); i++) { all[i].x = pts[2 * i]; all[i].y = pts[2 * i + 1]; }
    std::vector<uint8_t> d;
    const int m = compute(image, w, h, step, all, d);
    for (int i = 0; i < m && i < cap; i++) { kps[i] = all[i]; memcpy(desc + (size_t)i * 32, &d[(size_t)i * 32], 32); }
    return m;
}

}  // extern "C"
