/* ORACLE -- TEST INFRASTRUCTURE ONLY (see orb_oracle.h).  Extractor half.
 *
 * Follows /root/reference/src/ORBextractor.cc statement by statement where the reference has its own
 * code, and restates the OpenCV calls it makes (resize, GaussianBlur, FAST, fastAtan2, cvRound).
 * Compile with -ffp-contract=off.
 */
#include "orb_oracle.h"
#include "../include/orbb200_pattern.inc"

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <list>
#include <utility>
#include <chrono>
#include <vector>

namespace {

typedef unsigned char uchar;

inline int cvRoundF(float v) { return (int)lrintf(v); }   /* round-half-even (SSE cvtss2si) */
inline int cvRoundD(double v) { return (int)lrint(v); }
inline int cvFloorF(float v) { int i = (int)v; return i - (i > v); }
inline short sat_short(int v) { return (short)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }

/* ------------------------------------------------------------------------------------------------
 * cv::resize INTER_LINEAR, CV_8UC1 (OpenCV imgproc/resize.cpp: resizeGeneric_ + HResizeLinear +
 * VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>), INTER_RESIZE_COEF_BITS = 11.
 * Called by the reference at src/ORBextractor.cc:1120.
 * ---------------------------------------------------------------------------------------------- */
void resize_u8(const uchar* src, int sw, int sh, size_t sstep, uchar* dst, int dw, int dh, size_t dstep)
{
    const double scale_x = (double)sw / dw, scale_y = (double)sh / dh;
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> ialpha(dw * 2), ibeta(dh * 2);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cvFloorF(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        ialpha[dx * 2] = sat_short(cvRoundF((1.f - fx) * 2048));
        ialpha[dx * 2 + 1] = sat_short(cvRoundF(fx * 2048));
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cvFloorF(fy);
        fy -= sy;
        yofs[dy] = sy;
        ibeta[dy * 2] = sat_short(cvRoundF((1.f - fy) * 2048));
        ibeta[dy * 2 + 1] = sat_short(cvRoundF(fy * 2048));
    }
    std::vector<int> row0(dw), row1(dw);
    for (int dy = 0; dy < dh; dy++) {
        int sy0 = std::min(std::max(yofs[dy], 0), sh - 1);
        int sy1 = std::min(std::max(yofs[dy] + 1, 0), sh - 1);
        const uchar* S0 = src + (size_t)sy0 * sstep;
        const uchar* S1 = src + (size_t)sy1 * sstep;
        for (int dx = 0; dx < dw; dx++) {
            int sx = xofs[dx], sx1 = std::min(sx + 1, sw - 1);
            int a0 = ialpha[dx * 2], a1 = ialpha[dx * 2 + 1];
            row0[dx] = S0[sx] * a0 + S0[sx1] * a1;
            row1[dx] = S1[sx] * a0 + S1[sx1] * a1;
        }
        int b0 = ibeta[dy * 2], b1 = ibeta[dy * 2 + 1];
        uchar* D = dst + (size_t)dy * dstep;
        for (int dx = 0; dx < dw; dx++)
            D[dx] = (uchar)((((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2);
    }
}

inline int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

/* ------------------------------------------------------------------------------------------------
 * cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) on CV_8UC1 (OpenCV 4.x fixed-point path:
 * ufixedpoint16 kernel {18,34,48,56,48,34,18}/256, result (v + 2^15) >> 16).
 * Called by the reference at src/ORBextractor.cc:1086.
 * ---------------------------------------------------------------------------------------------- */
void gauss7_u8(const uchar* src, int w, int h, size_t sstep, uchar* dst, size_t dstep)
{
    /* horizontal pass into u16 rows (max 255*256 = 65280), vertical pass in u32; borders by index
     * reflection exactly as BORDER_REFLECT_101 */
    std::vector<uint16_t> hbuf((size_t)w * h);
    std::vector<uchar> pad((size_t)w + 6);
    for (int y = 0; y < h; y++) {
        const uchar* S = src + (size_t)y * sstep;
        for (int i = 0; i < 3; i++) { pad[i] = S[reflect101(i - 3, w)]; pad[w + 3 + i] = S[reflect101(w + i, w)]; }
        memcpy(pad.data() + 3, S, (size_t)w);
        uint16_t* H = &hbuf[(size_t)y * w];
        const uchar* P = pad.data();
        for (int x = 0; x < w; x++)
            H[x] = (uint16_t)(18 * (P[x] + P[x + 6]) + 34 * (P[x + 1] + P[x + 5]) + 48 * (P[x + 2] + P[x + 4]) + 56 * P[x + 3]);
    }
    for (int y = 0; y < h; y++) {
        const uint16_t* R[7];
        for (int k = 0; k < 7; k++) R[k] = &hbuf[(size_t)reflect101(y + k - 3, h) * w];
        uchar* D = dst + (size_t)y * dstep;
        for (int x = 0; x < w; x++) {
            uint32_t s = 18u * ((uint32_t)R[0][x] + R[6][x]) + 34u * ((uint32_t)R[1][x] + R[5][x]) +
                         48u * ((uint32_t)R[2][x] + R[4][x]) + 56u * (uint32_t)R[3][x];
            D[x] = (uchar)((s + 32768u) >> 16);
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * cv::FAST TYPE_9_16 (OpenCV features2d/fast.cpp FAST_t<16> + fast_score.cpp cornerScore<16>).
 * Called by the reference at src/ORBextractor.cc:809,814 on a per-cell sub-image.
 * ---------------------------------------------------------------------------------------------- */
const int RING_DX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int RING_DY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

/* M-1 where M = max over the 16 arcs of 9 contiguous ring pixels of min(+-(v - p_k)). */
inline int fast_score_px(const uchar* p, size_t step)
{
    int d[25];
    const int v = p[0];
    for (int k = 0; k < 16; k++) d[k] = v - p[(ptrdiff_t)RING_DY[k] * (ptrdiff_t)step + RING_DX[k]];
    for (int k = 16; k < 25; k++) d[k] = d[k - 16];
    int best = -256;
    for (int k = 0; k < 16; k++) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; j++) { mn = std::min(mn, d[k + j]); mx = std::max(mx, d[k + j]); }
        best = std::max(best, std::max(mn, -mx));
    }
    return best - 1;
}

void fast_score_map(const uchar* img, int w, int h, size_t step, int32_t* score)
{
    std::fill(score, score + (size_t)w * h, 0);
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int s = fast_score_px(img + (size_t)y * step + x, step);
            score[(size_t)y * w + x] = s > 0 ? s : 0;
        }
}

struct XYR { int x, y, r; };

void fast9(const uchar* img, int w, int h, size_t step, int threshold, bool nms, std::vector<XYR>& out)
{
    out.clear();
    if (w < 7 || h < 7) return;
    /* corner score where the pixel is a corner at this threshold, else 0 (OpenCV's zeroed row buffers).
     * Quick rejection as in OpenCV's scalar path: a 9-arc contains one pixel of every opposite pair. */
    std::vector<int> sc((size_t)w * h, 0);
    std::vector<char> isc((size_t)w * h, 0);
    ptrdiff_t ofs[16];
    for (int k = 0; k < 16; k++) ofs[k] = (ptrdiff_t)RING_DY[k] * (ptrdiff_t)step + RING_DX[k];
    for (int y = 3; y < h - 3; y++) {
        const uchar* row = img + (size_t)y * step;
        for (int x = 3; x < w - 3; x++) {
            const uchar* p = row + x;
            const int v = p[0], lo = v - threshold, hi = v + threshold;
            int br = 1, dk = 1;   /* still possible: brighter arc / darker arc */
            for (int k = 0; k < 8 && (br | dk); k += 2) {
                const int a = p[ofs[k]], b = p[ofs[k + 8]];
                br &= (a > hi) | (b > hi);
                dk &= (a < lo) | (b < lo);
            }
            if (!(br | dk)) continue;
            int s = fast_score_px(p, step);
            if (s >= threshold) { sc[(size_t)y * w + x] = s; isc[(size_t)y * w + x] = 1; }
        }
    }
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            if (!isc[(size_t)y * w + x]) continue;
            int s = sc[(size_t)y * w + x];
            if (nms) {
                const int* c = &sc[(size_t)y * w + x];
                if (!(s > c[-1] && s > c[1] && s > c[-w - 1] && s > c[-w] && s > c[-w + 1] &&
                      s > c[w - 1] && s > c[w] && s > c[w + 1]))
                    continue;
            }
            out.push_back({x, y, s});
        }
}

float fast_atan2(float y, float x)
{
    static const float p1 = 0.9997878412794807f * (float)(180 / M_PI);
    static const float p3 = -0.3258083974640975f * (float)(180 / M_PI);
    static const float p5 = 0.1555786518463281f * (float)(180 / M_PI);
    static const float p7 = -0.04432655554792128f * (float)(180 / M_PI);
    float ax = std::abs(x), ay = std::abs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* ------------------------------------------------------------------------------------------------
 * ORBextractor (reference src/ORBextractor.cc)
 * ---------------------------------------------------------------------------------------------- */
const int PATCH_SIZE = 31;
const int HALF_PATCH_SIZE = 15;
const int EDGE_THRESHOLD = 19;

const signed char PAT_X[512] = {ORBB200_PATTERN_X_INIT};
const signed char PAT_Y[512] = {ORBB200_PATTERN_Y_INIT};

struct Img {
    std::vector<uchar> buf;
    int cols = 0, rows = 0;
    size_t step = 0;
    uchar* ptr(int y, int x) { return buf.data() + (size_t)y * step + x; }
    const uchar* ptr(int y, int x) const { return buf.data() + (size_t)y * step + x; }
    void create(int w, int h) { cols = w; rows = h; step = (size_t)w; buf.assign((size_t)w * h, 0); }
};

/* IC_Angle, src/ORBextractor.cc:77-104 */
float IC_Angle(const Img& image, float ptx, float pty, const std::vector<int>& u_max)
{
    int m_01 = 0, m_10 = 0;
    const uchar* center = image.ptr(cvRoundF(pty), cvRoundF(ptx));
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
    int step = (int)image.step;
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int v_sum = 0;
        int d = u_max[v];
        for (int u = -d; u <= d; ++u) {
            int val_plus = center[u + v * step], val_minus = center[u - v * step];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return fast_atan2((float)m_01, (float)m_10);
}

const float factorPI = (float)(M_PI / 180.f);

/* computeOrbDescriptor, src/ORBextractor.cc:108-147 */
void computeOrbDescriptor(const oracle_kp_t& kpt, const Img& img, uchar* desc)
{
    float angle = (float)kpt.angle * factorPI;
    float a = (float)cosf(angle), b = (float)sinf(angle);
    const uchar* center = img.ptr(cvRoundF(kpt.y), cvRoundF(kpt.x));
    const int step = (int)img.step;
    int p = 0;
    for (int i = 0; i < 32; ++i, p += 16) {
        int val = 0;
        for (int k = 0; k < 8; k++) {
            int i0 = p + 2 * k, i1 = i0 + 1;
            int t0 = center[cvRoundF(PAT_X[i0] * b + PAT_Y[i0] * a) * step + cvRoundF(PAT_X[i0] * a - PAT_Y[i0] * b)];
            int t1 = center[cvRoundF(PAT_X[i1] * b + PAT_Y[i1] * a) * step + cvRoundF(PAT_X[i1] * a - PAT_Y[i1] * b)];
            val |= (t0 < t1) << k;
        }
        desc[i] = (uchar)val;
    }
}

struct KP { float x, y, response; };   /* pt + response: all DistributeOctTree looks at */

struct ExtractorNode {
    std::vector<KP> vKeys;
    int ULx = 0, ULy = 0, URx = 0, URy = 0, BLx = 0, BLy = 0, BRx = 0, BRy = 0;
    std::list<ExtractorNode>::iterator lit;
    bool bNoMore = false;
    long seq = 0;   /* creation order: stands in for the heap address (see header) */

    /* ExtractorNode::DivideNode, src/ORBextractor.cc:481-537 */
    void DivideNode(ExtractorNode& n1, ExtractorNode& n2, ExtractorNode& n3, ExtractorNode& n4)
    {
        const int halfX = (int)ceil(static_cast<float>(URx - ULx) / 2);
        const int halfY = (int)ceil(static_cast<float>(BRy - ULy) / 2);
        n1.ULx = ULx; n1.ULy = ULy;
        n1.URx = ULx + halfX; n1.URy = ULy;
        n1.BLx = ULx; n1.BLy = ULy + halfY;
        n1.BRx = ULx + halfX; n1.BRy = ULy + halfY;
        n2.ULx = n1.URx; n2.ULy = n1.URy;
        n2.URx = URx; n2.URy = URy;
        n2.BLx = n1.BRx; n2.BLy = n1.BRy;
        n2.BRx = URx; n2.BRy = ULy + halfY;
        n3.ULx = n1.BLx; n3.ULy = n1.BLy;
        n3.URx = n1.BRx; n3.URy = n1.BRy;
        n3.BLx = BLx; n3.BLy = BLy;
        n3.BRx = n1.BRx; n3.BRy = BLy;
        n4.ULx = n3.URx; n4.ULy = n3.URy;
        n4.URx = n2.BRx; n4.URy = n2.BRy;
        n4.BLx = n3.BRx; n4.BLy = n3.BRy;
        n4.BRx = BRx; n4.BRy = BRy;
        for (size_t i = 0; i < vKeys.size(); i++) {
            const KP& kp = vKeys[i];
            if (kp.x < n1.URx) {
                if (kp.y < n1.BRy) n1.vKeys.push_back(kp);
                else n3.vKeys.push_back(kp);
            } else if (kp.y < n1.BRy) n2.vKeys.push_back(kp);
            else n4.vKeys.push_back(kp);
        }
        if (n1.vKeys.size() == 1) n1.bNoMore = true;
        if (n2.vKeys.size() == 1) n2.bNoMore = true;
        if (n3.vKeys.size() == 1) n3.bNoMore = true;
        if (n4.vKeys.size() == 1) n4.bNoMore = true;
    }
};

typedef std::pair<int, ExtractorNode*> SizeNode;
struct SizeSeqLess {   /* pair<int,ExtractorNode*> operator< with seq standing in for the address */
    bool operator()(const SizeNode& a, const SizeNode& b) const
    {
        if (a.first != b.first) return a.first < b.first;
        return a.second->seq < b.second->seq;
    }
};

/* ORBextractor::DistributeOctTree, src/ORBextractor.cc:539-763 */
std::vector<KP> DistributeOctTree(const std::vector<KP>& vToDistributeKeys, const int minX, const int maxX,
                                  const int minY, const int maxY, const int N)
{
    std::vector<KP> vResultKeys;
    if (maxX - minX <= 0 || maxY - minY <= 0) return vResultKeys;
    const int nIni = (int)round(static_cast<float>(maxX - minX) / (maxY - minY));
    if (nIni <= 0) return vResultKeys;   /* reference divides by zero here; defined as "no keypoints" */
    const float hX = static_cast<float>(maxX - minX) / nIni;

    std::list<ExtractorNode> lNodes;
    long seq = 0;
    std::vector<ExtractorNode*> vpIniNodes(nIni);
    for (int i = 0; i < nIni; i++) {
        ExtractorNode ni;
        ni.ULx = (int)(hX * static_cast<float>(i)); ni.ULy = 0;
        ni.URx = (int)(hX * static_cast<float>(i + 1)); ni.URy = 0;
        ni.BLx = ni.ULx; ni.BLy = maxY - minY;
        ni.BRx = ni.URx; ni.BRy = maxY - minY;
        ni.seq = seq++;
        lNodes.push_back(ni);
        vpIniNodes[i] = &lNodes.back();
    }
    for (size_t i = 0; i < vToDistributeKeys.size(); i++) {
        const KP& kp = vToDistributeKeys[i];
        vpIniNodes[std::min((int)(kp.x / hX), nIni - 1)]->vKeys.push_back(kp);   /* min(): UB guard only */
    }
    std::list<ExtractorNode>::iterator lit = lNodes.begin();
    while (lit != lNodes.end()) {
        if (lit->vKeys.size() == 1) { lit->bNoMore = true; lit++; }
        else if (lit->vKeys.empty()) lit = lNodes.erase(lit);
        else lit++;
    }

    bool bFinish = false;
    std::vector<SizeNode> vSizeAndPointerToNode;

#define ORACLE_ADD_CHILD(n, COUNT)                                                      \
    if (n.vKeys.size() > 0) {                                                           \
        n.seq = seq++;                                                                  \
        lNodes.push_front(n);                                                           \
        if (n.vKeys.size() > 1) {                                                       \
            COUNT;                                                                      \
            vSizeAndPointerToNode.push_back(std::make_pair((int)n.vKeys.size(), &lNodes.front())); \
            lNodes.front().lit = lNodes.begin();                                        \
        }                                                                               \
    }

    while (!bFinish) {
        int prevSize = (int)lNodes.size();
        lit = lNodes.begin();
        int nToExpand = 0;
        vSizeAndPointerToNode.clear();
        while (lit != lNodes.end()) {
            if (lit->bNoMore) { lit++; continue; }
            ExtractorNode n1, n2, n3, n4;
            lit->DivideNode(n1, n2, n3, n4);
            ORACLE_ADD_CHILD(n1, nToExpand++)
            ORACLE_ADD_CHILD(n2, nToExpand++)
            ORACLE_ADD_CHILD(n3, nToExpand++)
            ORACLE_ADD_CHILD(n4, nToExpand++)
            lit = lNodes.erase(lit);
        }
        if ((int)lNodes.size() >= N || (int)lNodes.size() == prevSize) {
            bFinish = true;
        } else if (((int)lNodes.size() + nToExpand * 3) > N) {
            while (!bFinish) {
                prevSize = (int)lNodes.size();
                std::vector<SizeNode> vPrevSizeAndPointerToNode = vSizeAndPointerToNode;
                vSizeAndPointerToNode.clear();
                std::sort(vPrevSizeAndPointerToNode.begin(), vPrevSizeAndPointerToNode.end(), SizeSeqLess());
                for (int j = (int)vPrevSizeAndPointerToNode.size() - 1; j >= 0; j--) {
                    ExtractorNode n1, n2, n3, n4;
                    vPrevSizeAndPointerToNode[j].second->DivideNode(n1, n2, n3, n4);
                    ORACLE_ADD_CHILD(n1, (void)0)
                    ORACLE_ADD_CHILD(n2, (void)0)
                    ORACLE_ADD_CHILD(n3, (void)0)
                    ORACLE_ADD_CHILD(n4, (void)0)
                    lNodes.erase(vPrevSizeAndPointerToNode[j].second->lit);
                    if ((int)lNodes.size() >= N) break;
                }
                if ((int)lNodes.size() >= N || (int)lNodes.size() == prevSize) bFinish = true;
            }
        }
    }
#undef ORACLE_ADD_CHILD

    vResultKeys.reserve(lNodes.size());
    for (lit = lNodes.begin(); lit != lNodes.end(); lit++) {
        std::vector<KP>& vNodeKeys = lit->vKeys;
        KP* pKP = &vNodeKeys[0];
        float maxResponse = pKP->response;
        for (size_t k = 1; k < vNodeKeys.size(); k++) {
            if (vNodeKeys[k].response > maxResponse) { pKP = &vNodeKeys[k]; maxResponse = vNodeKeys[k].response; }
        }
        vResultKeys.push_back(*pKP);
    }
    return vResultKeys;
}

}  // namespace

struct oracle_extractor {
    int nfeatures; double scaleFactor; int nlevels, iniThFAST, minThFAST;
    std::vector<int> mnFeaturesPerLevel, umax;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<Img> mvImagePyramid, blurred;
    std::vector<std::vector<XYR> > candidates;          /* debug: per level, region coords */
    std::vector<std::vector<oracle_kp_t> > levelKeys;   /* debug: per level, level coords */
    /* accumulated wall time per stage (bench.py's per-stage CPU denominators): pyramid, FAST cells, octree, orientation, blur, descriptors */
    double stageNs[6] = {0, 0, 0, 0, 0, 0};
    static double now_ns() { return (double)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

    /* ORBextractor::ORBextractor, src/ORBextractor.cc:410-470 */
    oracle_extractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
        : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST)
    {
        mvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels);
        mvScaleFactor[0] = 1.0f; mvLevelSigma2[0] = 1.0f;
        for (int i = 1; i < nlevels; i++) {
            mvScaleFactor[i] = (float)(mvScaleFactor[i - 1] * scaleFactor);
            mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
        }
        mvInvScaleFactor.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
        for (int i = 0; i < nlevels; i++) {
            mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i];
            mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i];
        }
        mvImagePyramid.resize(nlevels); blurred.resize(nlevels);
        mnFeaturesPerLevel.resize(nlevels);
        float factor = (float)(1.0f / scaleFactor);
        float nDesiredFeaturesPerScale = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
        int sumFeatures = 0;
        for (int level = 0; level < nlevels - 1; level++) {
            mnFeaturesPerLevel[level] = cvRoundF(nDesiredFeaturesPerScale);
            sumFeatures += mnFeaturesPerLevel[level];
            nDesiredFeaturesPerScale *= factor;
        }
        mnFeaturesPerLevel[nlevels - 1] = std::max(nfeatures - sumFeatures, 0);

        umax.resize(HALF_PATCH_SIZE + 1);
        int v, v0, vmax = cvFloorF(HALF_PATCH_SIZE * sqrtf(2.f) / 2 + 1);
        int vmin = (int)ceilf(HALF_PATCH_SIZE * sqrtf(2.f) / 2);
        const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
        for (v = 0; v <= vmax; ++v) umax[v] = cvRoundD(sqrt(hp2 - v * v));
        for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
    }

    /* ORBextractor::ComputePyramid, src/ORBextractor.cc:1107-1132.  The 19-px reflect-101 border the
     * reference adds is never read by the path (keypoints are >=19 px inside), so it is not stored. */
    void ComputePyramid(const uchar* img, int w, int h, size_t step)
    {
        for (int level = 0; level < nlevels; ++level) {
            float scale = mvInvScaleFactor[level];
            int sw = cvRoundF((float)w * scale), sh = cvRoundF((float)h * scale);
            mvImagePyramid[level].create(std::max(sw, 0), std::max(sh, 0));
            if (sw <= 0 || sh <= 0) continue;
            if (level != 0) {
                const Img& p = mvImagePyramid[level - 1];
                resize_u8(p.buf.data(), p.cols, p.rows, p.step, mvImagePyramid[level].buf.data(), sw, sh, mvImagePyramid[level].step);
            } else {
                for (int y = 0; y < h; y++) memcpy(mvImagePyramid[0].ptr(y, 0), img + (size_t)y * step, (size_t)w);
            }
        }
    }

    /* ORBextractor::ComputeKeyPointsOctTree, src/ORBextractor.cc:765-853 */
    void ComputeKeyPointsOctTree(std::vector<std::vector<oracle_kp_t> >& allKeypoints)
    {
        allKeypoints.assign(nlevels, std::vector<oracle_kp_t>());
        candidates.assign(nlevels, std::vector<XYR>());
        const float W = 30;
        std::vector<XYR> vKeysCell;
        for (int level = 0; level < nlevels; ++level) {
            const Img& im = mvImagePyramid[level];
            const int minBorderX = EDGE_THRESHOLD - 3;
            const int minBorderY = minBorderX;
            const int maxBorderX = im.cols - EDGE_THRESHOLD + 3;
            const int maxBorderY = im.rows - EDGE_THRESHOLD + 3;
            std::vector<KP> vToDistributeKeys;
            const float width = (float)(maxBorderX - minBorderX);
            const float height = (float)(maxBorderY - minBorderY);
            const int nCols = (int)(width / W);
            const int nRows = (int)(height / W);
            /* levels too small for a single 30-px cell: the reference's arithmetic is undefined
             * (division by zero); defined here as "no keypoints on this level" */
            if (nCols <= 0 || nRows <= 0) continue;
            const int wCell = (int)ceil(width / nCols);
            const int hCell = (int)ceil(height / nRows);
            const double tFast0 = now_ns();
            for (int i = 0; i < nRows; i++) {
                const float iniY = (float)(minBorderY + i * hCell);
                float maxY = iniY + hCell + 6;
                if (iniY >= maxBorderY - 3) continue;
                if (maxY > maxBorderY) maxY = (float)maxBorderY;
                for (int j = 0; j < nCols; j++) {
                    const float iniX = (float)(minBorderX + j * wCell);
                    float maxX = iniX + wCell + 6;
                    if (iniX >= maxBorderX - 6) continue;
                    if (maxX > maxBorderX) maxX = (float)maxBorderX;
                    const int y0 = (int)iniY, y1 = (int)maxY, x0 = (int)iniX, x1 = (int)maxX;
                    fast9(im.ptr(y0, x0), x1 - x0, y1 - y0, im.step, iniThFAST, true, vKeysCell);
                    if (vKeysCell.empty())
                        fast9(im.ptr(y0, x0), x1 - x0, y1 - y0, im.step, minThFAST, true, vKeysCell);
                    for (size_t k = 0; k < vKeysCell.size(); k++) {
                        KP kp;
                        kp.x = (float)vKeysCell[k].x + j * wCell;
                        kp.y = (float)vKeysCell[k].y + i * hCell;
                        kp.response = (float)vKeysCell[k].r;
                        vToDistributeKeys.push_back(kp);
                        candidates[level].push_back({(int)kp.x, (int)kp.y, vKeysCell[k].r});
                    }
                }
            }
            const double tOct0 = now_ns();
            stageNs[1] += tOct0 - tFast0;
            std::vector<KP> keys = DistributeOctTree(vToDistributeKeys, minBorderX, maxBorderX, minBorderY, maxBorderY,
                                                     mnFeaturesPerLevel[level]);
            stageNs[2] += now_ns() - tOct0;
            const int scaledPatchSize = (int)(PATCH_SIZE * mvScaleFactor[level]);
            std::vector<oracle_kp_t>& keypoints = allKeypoints[level];
            keypoints.resize(keys.size());
            for (size_t i = 0; i < keys.size(); i++) {
                oracle_kp_t& k = keypoints[i];
                k.x = keys[i].x + minBorderX;
                k.y = keys[i].y + minBorderY;
                k.size = (float)scaledPatchSize;
                k.angle = -1.f;
                k.response = keys[i].response;
                k.octave = level;
                k.class_id = -1;
            }
        }
        const double tOri0 = now_ns();
        for (int level = 0; level < nlevels; ++level)
            for (size_t i = 0; i < allKeypoints[level].size(); i++)
                allKeypoints[level][i].angle = IC_Angle(mvImagePyramid[level], allKeypoints[level][i].x, allKeypoints[level][i].y, umax);
        stageNs[3] += now_ns() - tOri0;
    }

    /* ORBextractor::operator(), src/ORBextractor.cc:1043-1105 */
    int extract(const uchar* img, int w, int h, size_t step, oracle_kp_t* kps, uchar* desc, int cap)
    {
        if (!img || w <= 0 || h <= 0) return 0;
        const double tPyr0 = now_ns();
        ComputePyramid(img, w, h, step);
        stageNs[0] += now_ns() - tPyr0;
        ComputeKeyPointsOctTree(levelKeys);
        int nkeypoints = 0;
        for (int level = 0; level < nlevels; ++level) nkeypoints += (int)levelKeys[level].size();
        if (nkeypoints > cap) return -1;
        int offset = 0;
        for (int level = 0; level < nlevels; ++level) {
            std::vector<oracle_kp_t>& keypoints = levelKeys[level];
            int nkeypointsLevel = (int)keypoints.size();
            const Img& src = mvImagePyramid[level];
            const double tBlur0 = now_ns();
            blurred[level].create(src.cols, src.rows);
            if (src.cols > 0 && src.rows > 0)
                gauss7_u8(src.buf.data(), src.cols, src.rows, src.step, blurred[level].buf.data(), blurred[level].step);
            const double tDesc0 = now_ns();
            stageNs[4] += tDesc0 - tBlur0;
            if (nkeypointsLevel == 0) continue;
            for (int i = 0; i < nkeypointsLevel; i++)
                computeOrbDescriptor(keypoints[i], blurred[level], desc + (size_t)(offset + i) * 32);
            stageNs[5] += now_ns() - tDesc0;
            float scale = mvScaleFactor[level];
            for (int i = 0; i < nkeypointsLevel; i++) {
                oracle_kp_t k = keypoints[i];
                if (level != 0) { k.x *= scale; k.y *= scale; }
                kps[offset + i] = k;
            }
            offset += nkeypointsLevel;
        }
        return nkeypoints;
    }
};

/* Frame::ComputeStereoMatches, src/Frame.cc:662-836.  L and R are the two extractors after operator() on the left
 * and right image (the reference reads their public mvImagePyramid, :669,759,771,776). */
static int ComputeStereoMatches(const oracle_extractor* EL, const oracle_extractor* ER,
                                const oracle_kp_t* mvKeys, const uint8_t* mDescriptors, int N,
                                const oracle_kp_t* mvKeysRight, const uint8_t* mDescriptorsRight, int Nr,
                                float mb, float mbf, float* mvuRight, float* mvDepth)
{
    const int TH_HIGH = 100, TH_LOW = 50;
    for (int i = 0; i < N; i++) { mvuRight[i] = -1.0f; mvDepth[i] = -1.0f; }
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
    const int nRows = EL->mvImagePyramid[0].rows;
    const std::vector<float>& mvScaleFactors = EL->mvScaleFactor;
    const std::vector<float>& mvInvScaleFactors = EL->mvInvScaleFactor;
    std::vector<std::vector<size_t> > vRowIndices(nRows, std::vector<size_t>());
    for (int iR = 0; iR < Nr; iR++) {
        const oracle_kp_t& kp = mvKeysRight[iR];
        const float kpY = kp.y;
        const float r = 2.0f * mvScaleFactors[kp.octave];
        const int maxr = (int)ceil(kpY + r);
        const int minr = (int)floor(kpY - r);
        for (int yi = minr; yi <= maxr; yi++)
            if (yi >= 0 && yi < nRows) vRowIndices[yi].push_back(iR);   /* bounds: UB guard only */
    }
    const float minZ = mb;
    const float minD = 0;
    const float maxD = mbf / minZ;
    std::vector<std::pair<int, int> > vDistIdx;
    vDistIdx.reserve(N);
    auto desc_dist = [](const uint8_t* a, const uint8_t* b) {
        int d = 0;
        for (int i = 0; i < 8; i++) { uint32_t wa, wb; memcpy(&wa, a + 4 * i, 4); memcpy(&wb, b + 4 * i, 4); d += __builtin_popcount(wa ^ wb); }
        return d;
    };
    for (int iL = 0; iL < N; iL++) {
        const oracle_kp_t& kpL = mvKeys[iL];
        const int levelL = kpL.octave;
        const float vL = kpL.y;
        const float uL = kpL.x;
        if ((size_t)vL >= (size_t)nRows) continue;                     /* UB guard only */
        const std::vector<size_t>& vCandidates = vRowIndices[(size_t)vL];
        if (vCandidates.empty()) continue;
        const float minU = uL - maxD;
        const float maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH;
        size_t bestIdxR = 0;
        const uint8_t* dL = mDescriptors + (size_t)iL * 32;
        for (size_t iC = 0; iC < vCandidates.size(); iC++) {
            const size_t iR = vCandidates[iC];
            const oracle_kp_t& kpR = mvKeysRight[iR];
            if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
            const float uR = kpR.x;
            if (uR >= minU && uR <= maxU) {
                const int dist = desc_dist(dL, mDescriptorsRight + iR * 32);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < thOrbDist) {
            const float uR0 = mvKeysRight[bestIdxR].x;
            const float scaleFactor = mvInvScaleFactors[kpL.octave];
            const float scaleduL = roundf(kpL.x * scaleFactor);
            const float scaledvL = roundf(kpL.y * scaleFactor);
            const float scaleduR0 = roundf(uR0 * scaleFactor);
            const int w = 5;
            const Img& PL = EL->mvImagePyramid[kpL.octave];
            const Img& PR = ER->mvImagePyramid[kpL.octave];
            /* IL = patch - centre (CV_32F); values are small integers, so float sums below are exact */
            const int yL0 = (int)(scaledvL - w), xL0 = (int)(scaleduL - w);
            const float cL = (float)*PL.ptr(yL0 + w, xL0 + w);
            int bestDistS = INT32_MAX;
            int bestincR = 0;
            const int L = 5;
            std::vector<float> vDists(2 * L + 1);
            const float iniu = scaleduR0 + L - w;
            const float endu = scaleduR0 + L + w + 1;
            if (iniu < 0 || endu >= PR.cols) continue;
            for (int incR = -L; incR <= +L; incR++) {
                const int xR0 = (int)(scaleduR0 + incR - w);
                const float cR = (float)*PR.ptr(yL0 + w, xR0 + w);
                float dist = 0;   /* cv::norm(IL,IR,NORM_L1) */
                for (int dy = 0; dy < 2 * w + 1; dy++)
                    for (int dx = 0; dx < 2 * w + 1; dx++) {
                        const float a = (float)*PL.ptr(yL0 + dy, xL0 + dx) - cL;
                        const int xr = std::min(std::max(xR0 + dx, 0), PR.cols - 1);   /* the reference reads its 19-px border here; never reached for real keypoints */
                        const float b = (float)*PR.ptr(yL0 + dy, xr) - cR;
                        dist += fabsf(a - b);
                    }
                if (dist < bestDistS) { bestDistS = (int)dist; bestincR = incR; }
                vDists[L + incR] = dist;
            }
            if (bestincR == -L || bestincR == L) continue;
            const float dist1 = vDists[L + bestincR - 1];
            const float dist2 = vDists[L + bestincR];
            const float dist3 = vDists[L + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = mvScaleFactors[kpL.octave] * ((float)scaleduR0 + (float)bestincR + deltaR);
            float disparity = (uL - bestuR);
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) { disparity = 0.01; bestuR = uL - 0.01; }
                mvDepth[iL] = mbf / disparity;
                mvuRight[iL] = bestuR;
                vDistIdx.push_back(std::pair<int, int>(bestDistS, iL));
            }
        }
    }
    if (vDistIdx.empty()) return 0;                                    /* reference indexes an empty vector here */
    std::sort(vDistIdx.begin(), vDistIdx.end());
    const float median = (float)vDistIdx[vDistIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    int kept = (int)vDistIdx.size();
    for (int i = (int)vDistIdx.size() - 1; i >= 0; i--) {
        if (vDistIdx[i].first < thDist) break;
        mvuRight[vDistIdx[i].second] = -1;
        mvDepth[vDistIdx[i].second] = -1;
        kept--;
    }
    return kept;
}

extern "C" {

int oracle_compute_stereo_matches(oracle_extractor* EL, oracle_extractor* ER,
                                  const oracle_kp_t* kpsL, const uint8_t* descL, int nL,
                                  const oracle_kp_t* kpsR, const uint8_t* descR, int nR,
                                  float mb, float mbf, float* uRight, float* depth)
{ return ComputeStereoMatches(EL, ER, kpsL, descL, nL, kpsR, descR, nR, mb, mbf, uRight, depth); }

void oracle_resize_u8(const uint8_t* src, int sw, int sh, size_t sstep, uint8_t* dst, int dw, int dh, size_t dstep)
{ resize_u8(src, sw, sh, sstep, dst, dw, dh, dstep); }

void oracle_gauss7_u8(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep)
{ gauss7_u8(src, w, h, sstep, dst, dstep); }

int oracle_fast9(const uint8_t* img, int w, int h, size_t step, int threshold, int nms, int32_t* xyr, int cap)
{
    std::vector<XYR> out;
    fast9(img, w, h, step, threshold, nms != 0, out);
    for (int i = 0; i < (int)out.size() && i < cap; i++) { xyr[3 * i] = out[i].x; xyr[3 * i + 1] = out[i].y; xyr[3 * i + 2] = out[i].r; }
    return (int)out.size();
}

void oracle_fast_score_map(const uint8_t* img, int w, int h, size_t step, int32_t* score) { fast_score_map(img, w, h, step, score); }
float oracle_fast_atan2(float y, float x) { return fast_atan2(y, x); }
int oracle_cv_round(float v) { return cvRoundF(v); }

oracle_extractor* oracle_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST)
{ return new oracle_extractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST); }
void oracle_extractor_destroy(oracle_extractor* e) { delete e; }

int oracle_extract(oracle_extractor* e, const uint8_t* img, int w, int h, size_t step, oracle_kp_t* kps, uint8_t* desc, int cap)
{ return e->extract(img, w, h, step, kps, desc, cap); }

/* accumulated milliseconds per stage since creation / the last reset: pyramid, FAST cells, octree, orientation, blur, descriptors */
void oracle_extractor_stage_ms(oracle_extractor* e, double* out6, int reset)
{
    for (int i = 0; i < 6; i++) { if (out6) out6[i] = e->stageNs[i] * 1e-6; if (reset) e->stageNs[i] = 0; }
}

int oracle_extractor_features_per_level(oracle_extractor* e, int32_t* out)
{ for (int i = 0; i < e->nlevels; i++) out[i] = e->mnFeaturesPerLevel[i]; return e->nlevels; }
int oracle_extractor_scale_factors(oracle_extractor* e, float* out)
{ for (int i = 0; i < e->nlevels; i++) out[i] = e->mvScaleFactor[i]; return e->nlevels; }
int oracle_extractor_umax(oracle_extractor* e, int32_t* out)
{ for (int i = 0; i <= HALF_PATCH_SIZE; i++) out[i] = e->umax[i]; return HALF_PATCH_SIZE + 1; }

int oracle_extractor_level_size(oracle_extractor* e, int level, int* w, int* h)
{ *w = e->mvImagePyramid[level].cols; *h = e->mvImagePyramid[level].rows; return 0; }

int oracle_extractor_level_image(oracle_extractor* e, int level, int blurred, uint8_t* dst, size_t dstep)
{
    const Img& im = blurred ? e->blurred[level] : e->mvImagePyramid[level];
    for (int y = 0; y < im.rows; y++) memcpy(dst + (size_t)y * dstep, im.ptr(y, 0), (size_t)im.cols);
    return 0;
}

int oracle_extractor_level_candidates(oracle_extractor* e, int level, int32_t* xyr, int cap)
{
    const std::vector<XYR>& c = e->candidates[level];
    for (int i = 0; i < (int)c.size() && i < cap; i++) { xyr[3 * i] = c[i].x; xyr[3 * i + 1] = c[i].y; xyr[3 * i + 2] = c[i].r; }
    return (int)c.size();
}

int oracle_extractor_level_keypoints(oracle_extractor* e, int level, oracle_kp_t* kps, int cap)
{
    const std::vector<oracle_kp_t>& k = e->levelKeys[level];
    for (int i = 0; i < (int)k.size() && i < cap; i++) kps[i] = k[i];
    return (int)k.size();
}

int oracle_distribute_octree(const int32_t* xyr, int n, int minX, int maxX, int minY, int maxY, int N, int32_t* out_xyr, int cap)
{
    std::vector<KP> in(n);
    for (int i = 0; i < n; i++) { in[i].x = (float)xyr[3 * i]; in[i].y = (float)xyr[3 * i + 1]; in[i].response = (float)xyr[3 * i + 2]; }
    std::vector<KP> out = DistributeOctTree(in, minX, maxX, minY, maxY, N);
    for (int i = 0; i < (int)out.size() && i < cap; i++) {
        out_xyr[3 * i] = (int)out[i].x; out_xyr[3 * i + 1] = (int)out[i].y; out_xyr[3 * i + 2] = (int)out[i].response;
    }
    return (int)out.size();
}

}  // extern "C"
