// Bodies of all ORB_SLAM2::ORBmatcher methods over the C ABI of liborbb200.so (include/orbb200.h).
//
// How a maintainer uses this file: add it to the ORB_SLAM2 library sources (CMakeLists.txt:49-75) in place of
// src/ORBmatcher.cc -- the class declaration in include/ORBmatcher.h, and therefore every
// call site in Tracking.cc / LocalMapping.cc / LoopClosing.cc, stays untouched.  Each method
//   1. does the per-query geometry on the host with the same cv::Mat / float expressions the reference evaluates (projection,
//      frustum and distance gates, PredictScale) -- a few dozen flops per query, and their rounding is part of the contract;
//   2. flattens the object graph into arrays and makes ONE ABI call for the data-parallel part (window scans + Hamming
//      distances + the loop-carried "already matched" state + the rotation histogram);
//   3. applies the result to the members the reference loop mutates, in the reference's order (for Fuse: a sequential replay of
//      the map mutations with the live isBad() / IsInKeyFrame() / GetMapPoint() tests).
// Here (no reference tree, no OpenCV SDK) it is compiled against compat/*.h, which declare the same members.  matcher_suite.cpp
// drives every method through Frame / KeyFrame / MapPoint objects; tests/test_matcher_suite.py compares what it returns and
// leaves in the objects with the reference's own, unmodified src/ORBmatcher.cc compiled against the same object model.
//
//   method                                                           reference lines   ABI entry
//   DescriptorDistance                                               :1647-1663        (host: one pair is not GPU work)
//   SearchByProjection(Frame&, vector<MapPoint*>&, th)               :45-129           orbb200_search_by_projection
//   SearchByProjection(Frame& Cur, const Frame& Last, th, bMono)     :1328-1470        orbb200_search_by_projection_frame
//   SearchByProjection(Frame& Cur, KeyFrame*, set, th, ORBdist)      :1472-1599        orbb200_search_window_best (BLOCK | ORI)
//   SearchByProjection(KeyFrame*, Scw, points, vpMatched, th)        :290-403          orbb200_search_window_best (BLOCK)
//   SearchByBoW(KeyFrame*, Frame&, ...) / (KeyFrame*, KeyFrame*, ..) :159-288 :522-655 orbb200_search_by_bow
//   SearchForInitialization                                          :405-520          orbb200_search_for_initialization
//   SearchForTriangulation                                           :657-823          orbb200_search_for_triangulation
//   SearchBySim3                                                     :1102-1326        orbb200_search_window_best x2 + mutual check
//   Fuse(KeyFrame*, points, th)                                      :825-975          orbb200_search_window_best (CHI2) + replay
//   Fuse(KeyFrame*, Scw, points, th, vpReplacePoint)                 :977-1100         orbb200_search_window_best + replay
//   BirdviewMatch(F1, F2, vnMatches12, vPrevMatched, windowSize)     :1667-1786        orbb200_birdview_match (prev_xy)
//   BirdviewMatch(const F1, const F2, vnMatches12, windowSize)       :1788-1899        orbb200_birdview_match
//   SearchByMatchBird(Frame& Cur, const Frame& Last, windowSize)     :1901-1921        the same + the reference's copy loop
//   SearchByProjectionBird(Frame&, vector<MapPointBird*>&, r)        :1923-1998        orbb200_search_by_projection_bird
//   SearchByMatchBird(KeyFrame*, Frame&, vpMatches, r)               :2000-2114        orbb200_search_by_match_bird_kf
#include "ORBmatcher.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "orbb200_host.h"

using namespace std;

namespace ORB_SLAM2
{

namespace
{
typedef vector<uint8_t> Bytes;
using orbb200_host::check;

// n descriptor rows of a CV_8U matrix as one contiguous block (rows of a cv::Mat may be strided)
const uint8_t* rows32(const cv::Mat& m, int n, Bytes& tmp)
{
    if (n <= 0) return nullptr;
    if (n == 1 || (size_t)m.step == 32) return m.ptr(0);
    tmp.resize((size_t)n * 32);
    for (int i = 0; i < n; i++) memcpy(&tmp[32 * (size_t)i], m.ptr(i), 32);
    return tmp.data();
}

// Device copies (keypoints + descriptors + uRight + lookup grid), cached per thread by content: Tracking calls three to five
// matchers on one Frame (src/Tracking.cc:1227-1245, 1659-1672).
orbb200_frame* front(orbb200_ctx* c, const Frame& F)
{
    Bytes tmp;
    return orbb200_host::frames().get(c, F.mvKeysUn.data(), rows32(F.mDescriptors, F.N, tmp), (int)F.mvuRight.size() == F.N ? F.mvuRight.data() : nullptr,
                                      F.N, Frame::mnMinX, Frame::mnMinY, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv);
}
orbb200_frame* bird(orbb200_ctx* c, const Frame& F)
{
    Bytes tmp;
    const int n = (int)F.mvKeysBird.size();
    return orbb200_host::frames().get(c, F.mvKeysBird.data(), rows32(F.mDescriptorsBird, n, tmp), nullptr, n, 0.f, 0.f,
                                      Frame::mfGridElementWidthInvBirdview, Frame::mfGridElementHeightInvBirdview);
}
orbb200_frame* front(orbb200_ctx* c, KeyFrame* pKF)
{
    Bytes tmp;
    return orbb200_host::frames().get(c, pKF->mvKeysUn.data(), rows32(pKF->mDescriptors, pKF->N, tmp), (int)pKF->mvuRight.size() == pKF->N ? pKF->mvuRight.data() : nullptr,
                                      pKF->N, (float)pKF->mnMinX, (float)pKF->mnMinY, pKF->mfGridElementWidthInv, pKF->mfGridElementHeightInv);
}

// DBoW2::FeatureVector is a std::map: its iteration order (ascending node id) is the order of the reference's merge loops
struct FeatCSR
{
    vector<int32_t> node, ptr, idx;
    explicit FeatCSR(const DBoW2::FeatureVector& fv)
    {
        ptr.push_back(0);
        for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it)
        {
            node.push_back((int32_t)it->first);
            for (size_t k = 0; k < it->second.size(); k++) idx.push_back((int32_t)it->second[k]);
            ptr.push_back((int32_t)idx.size());
        }
    }
    int n() const { return (int)node.size(); }
};

// query arrays of orbb200_search_window_best
struct Windows
{
    int n;
    Bytes valid, desc;
    vector<float> x, y, r, aux, angle;
    vector<int32_t> lo, hi, bestIdx, bestDist;
    explicit Windows(int n_) : n(n_), valid(n_, 0), desc((size_t)n_ * 32), x(n_), y(n_), r(n_), aux(n_), angle(n_), lo(n_), hi(n_), bestIdx(n_, -1), bestDist(n_, 256) {}
    void set(int i, float u, float v, float radius, int minLevel, int maxLevel, const cv::Mat& d)
    {
        valid[i] = 1; x[i] = u; y[i] = v; r[i] = radius; lo[i] = minLevel; hi[i] = maxLevel;
        memcpy(&desc[32 * (size_t)i], d.ptr(0), 32);
    }
    // returns nmatches; queryOfKp [nKp]
    int run(orbb200_ctx* c, orbb200_frame* f, int nKp, const uint8_t* kpBlocked, const float* invSigma2, int accTh, int flags, vector<int32_t>& queryOfKp)
    {
        queryOfKp.assign(max(nKp, 1), -1);
        int nmatches = 0;
        if (n == 0 || nKp == 0) return 0;
        check(orbb200_search_window_best(c, f, n, valid.data(), x.data(), y.data(), r.data(), lo.data(), hi.data(), desc.data(), aux.data(), angle.data(),
                                         nullptr, kpBlocked, invSigma2, accTh, flags, bestIdx.data(), bestDist.data(), queryOfKp.data(), &nmatches),
              "orbb200_search_window_best");
        return nmatches;
    }
};

// [sR | t] of a Sim3 / SE3 matrix split the way src/ORBmatcher.cc:299-303 and :986-990 do
struct Sim3Pose
{
    cv::Mat Rcw, tcw, Ow;
    explicit Sim3Pose(const cv::Mat& Scw)
    {
        cv::Mat sRcw = Scw.rowRange(0,3).colRange(0,3);
        const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
        Rcw = sRcw/scw;
        tcw = Scw.rowRange(0,3).col(3)/scw;
        Ow = -Rcw.t()*tcw;
    }
};

// the gates every KeyFrame projection search applies after the projection (:342-357, :872-887, :1030-1046): distance inside the
// scale-invariance region, viewing angle below 60 degrees; returns the predicted level or -1
int gate_and_level(MapPoint* pMP, const cv::Mat& p3Dw, const cv::Mat& Ow, KeyFrame* pKF, float& dist3D)
{
    const float maxDistance = pMP->GetMaxDistanceInvariance();
    const float minDistance = pMP->GetMinDistanceInvariance();
    cv::Mat PO = p3Dw-Ow;
    dist3D = cv::norm(PO);
    if(dist3D<minDistance || dist3D>maxDistance)
        return -1;
    cv::Mat Pn = pMP->GetNormal();
    if(PO.dot(Pn)<0.5*dist3D)
        return -1;
    return pMP->PredictScale(dist3D,pKF);
}

// Pinhole projection of a camera-frame point the way the KeyFrame searches write it -- x = X * invz, u = fx * x + cx -- behind their
// "depth must be positive" test.  The reciprocal is a float division in SearchByProjection(KeyFrame*, Scw, ...) and
// Fuse(KeyFrame*, points, th) (:331, :859) and a double division rounded to float in Fuse(KeyFrame*, Scw, ...) and SearchBySim3
// (:1019, :1166, :1246): both are kept, the last bit of invz decides which cell a window starts in.
struct Pixel { float u, v, invz; };
bool pinhole(const cv::Mat& p3Dc, float fx, float fy, float cx, float cy, bool doubleReciprocal, Pixel& px)
{
    const float z = p3Dc.at<float>(2);
    if(z<0.0f)
        return false;
    px.invz = doubleReciprocal ? (float)(1.0/z) : 1/z;
    const float x = p3Dc.at<float>(0)*px.invz;
    const float y = p3Dc.at<float>(1)*px.invz;
    px.u = fx*x+cx;
    px.v = fy*y+cy;
    return true;
}
}  // namespace

// the constants and the constructor of src/ORBmatcher.cc:34-43 (this file takes that file's place)
const int ORBmatcher::TH_HIGH = 100;
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;

ORBmatcher::ORBmatcher(float nnratio, bool checkOri): mfNNratio(nnratio), mbCheckOrientation(checkOri)
{
}

int ORBmatcher::DescriptorDistance(const cv::Mat &a, const cv::Mat &b)
{
    const uint32_t* pa = reinterpret_cast<const uint32_t*>(a.ptr(0));
    const uint32_t* pb = reinterpret_cast<const uint32_t*>(b.ptr(0));
    int dist = 0;
    for (int i = 0; i < 8; i++) dist += __builtin_popcount(pa[i] ^ pb[i]);
    return dist;
}

int ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th)
{
    orbb200_ctx* c = orbb200_host::ThreadContext(F.mvScaleFactors);
    const int nq = (int)vpMapPoints.size();
    if (nq == 0 || F.N == 0) return 0;
    Bytes valid(nq), obs(nq), desc((size_t)nq * 32), blocked(F.N);
    vector<float> u(nq), v(nq), uR(nq), vc(nq);
    vector<int32_t> lvl(nq);
    for (int i = 0; i < nq; i++)
    {
        MapPoint* pMP = vpMapPoints[i];
        valid[i] = pMP->mbTrackInView && !pMP->isBad();                              // :53-58
        u[i] = pMP->mTrackProjX; v[i] = pMP->mTrackProjY; uR[i] = pMP->mTrackProjXR;
        lvl[i] = pMP->mnTrackScaleLevel; vc[i] = pMP->mTrackViewCos;
        if (valid[i] && (lvl[i] < 0 || lvl[i] >= F.mnScaleLevels)) { fprintf(stderr, "orbb200: mnTrackScaleLevel %d outside the frame's pyramid\n", lvl[i]); abort(); }
        obs[i] = pMP->Observations() > 0;                                            // what a later :87-89 test will see
        const cv::Mat d = pMP->GetDescriptor();
        memcpy(&desc[32 * (size_t)i], d.ptr(0), 32);
    }
    for (int k = 0; k < F.N; k++)
        blocked[k] = F.mvpMapPoints[k] && F.mvpMapPoints[k]->Observations() > 0;     // :87-89
    vector<int32_t> bi(nq), bd(nq), qk(F.N);
    int nmatches = 0;
    check(orbb200_search_by_projection(c, front(c, F), nq, valid.data(), u.data(), v.data(), uR.data(), lvl.data(), vc.data(), desc.data(), obs.data(),
                                       blocked.data(), th, mfNNratio, bi.data(), bd.data(), qk.data(), &nmatches), "orbb200_search_by_projection");
    for (int i = 0; i < nq; i++)
        if (bi[i] >= 0) F.mvpMapPoints[bi[i]] = vpMapPoints[i];                      // in query order == the loop's writes (:123)
    return nmatches;
}

int ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono)
{
    orbb200_ctx* c = orbb200_host::ThreadContext(CurrentFrame.mvScaleFactors);
    const int nq = LastFrame.N, n = CurrentFrame.N;
    if (nq == 0 || n == 0) return 0;

    const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0,3).colRange(0,3);
    const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0,3).col(3);
    const cv::Mat twc = -Rcw.t()*tcw;
    const cv::Mat Rlw = LastFrame.mTcw.rowRange(0,3).colRange(0,3);
    const cv::Mat tlw = LastFrame.mTcw.rowRange(0,3).col(3);
    const cv::Mat tlc = Rlw*twc+tlw;
    const bool bForward = tlc.at<float>(2)>CurrentFrame.mb && !bMono;
    const bool bBackward = -tlc.at<float>(2)>CurrentFrame.mb && !bMono;

    Bytes valid(nq, 0), obs(nq, 0), desc((size_t)nq * 32), blocked(n);
    vector<float> qu(nq), qv(nq), qinvz(nq), qangle(nq);
    vector<int32_t> qoct(nq);
    for (int i = 0; i < nq; i++)
    {
        MapPoint* pMP = LastFrame.mvpMapPoints[i];
        if (!pMP || LastFrame.mvbOutlier[i]) continue;
        cv::Mat x3Dw = pMP->GetWorldPos();
        cv::Mat x3Dc = Rcw*x3Dw+tcw;
        const float xc = x3Dc.at<float>(0);
        const float yc = x3Dc.at<float>(1);
        const float invzc = 1.0/x3Dc.at<float>(2);
        if(invzc<0)
            continue;
        float u = CurrentFrame.fx*xc*invzc+CurrentFrame.cx;
        float v = CurrentFrame.fy*yc*invzc+CurrentFrame.cy;
        if(u<CurrentFrame.mnMinX || u>CurrentFrame.mnMaxX)
            continue;
        if(v<CurrentFrame.mnMinY || v>CurrentFrame.mnMaxY)
            continue;
        valid[i] = 1; qu[i] = u; qv[i] = v; qinvz[i] = invzc;
        qoct[i] = LastFrame.mvKeys[i].octave;
        qangle[i] = LastFrame.mvKeysUn[i].angle;
        obs[i] = pMP->Observations() > 0;
        const cv::Mat d = pMP->GetDescriptor();
        memcpy(&desc[32 * (size_t)i], d.ptr(0), 32);
    }
    for (int k = 0; k < n; k++)
        blocked[k] = CurrentFrame.mvpMapPoints[k] && CurrentFrame.mvpMapPoints[k]->Observations() > 0;   // :1403-1405
    vector<int32_t> qk(n, -1);
    int nmatches = 0;
    check(orbb200_search_by_projection_frame(c, front(c, CurrentFrame), nq, valid.data(), qu.data(), qv.data(), qinvz.data(), qoct.data(), qangle.data(),
                                             desc.data(), obs.data(), blocked.data(), th, CurrentFrame.mbf, bForward ? 1 : (bBackward ? 2 : 0),
                                             mbCheckOrientation ? 1 : 0, qk.data(), &nmatches), "orbb200_search_by_projection_frame");
    // The tracker clears CurrentFrame.mvpMapPoints before this call (src/Tracking.cc:1219), so "left NULL by the rotation filter"
    // and "never touched" are the same state.
    for (int k = 0; k < n; k++)
        if (qk[k] >= 0) CurrentFrame.mvpMapPoints[k] = LastFrame.mvpMapPoints[qk[k]];
    return nmatches;
}

int ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, const float th , const int ORBdist)
{
    orbb200_ctx* c = orbb200_host::ThreadContext(CurrentFrame.mvScaleFactors);
    const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0,3).colRange(0,3);
    const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0,3).col(3);
    const cv::Mat Ow = -Rcw.t()*tcw;

    const vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();
    const int nq = (int)vpMPs.size(), n = CurrentFrame.N;
    if (nq == 0 || n == 0) return 0;
    Windows W(nq);
    for (int i = 0; i < nq; i++)
    {
        MapPoint* pMP = vpMPs[i];
        if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;
        cv::Mat x3Dw = pMP->GetWorldPos();
        cv::Mat x3Dc = Rcw*x3Dw+tcw;
        const float xc = x3Dc.at<float>(0);
        const float yc = x3Dc.at<float>(1);
        const float invzc = 1.0/x3Dc.at<float>(2);
        const float u = CurrentFrame.fx*xc*invzc+CurrentFrame.cx;
        const float v = CurrentFrame.fy*yc*invzc+CurrentFrame.cy;
        if(u<CurrentFrame.mnMinX || u>CurrentFrame.mnMaxX)
            continue;
        if(v<CurrentFrame.mnMinY || v>CurrentFrame.mnMaxY)
            continue;
        cv::Mat PO = x3Dw-Ow;
        float dist3D = cv::norm(PO);
        const float maxDistance = pMP->GetMaxDistanceInvariance();
        const float minDistance = pMP->GetMinDistanceInvariance();
        if(dist3D<minDistance || dist3D>maxDistance)
            continue;
        int nPredictedLevel = pMP->PredictScale(dist3D,&CurrentFrame);
        const float radius = th*CurrentFrame.mvScaleFactors[nPredictedLevel];
        W.set(i, u, v, radius, nPredictedLevel-1, nPredictedLevel+1, pMP->GetDescriptor());
        W.angle[i] = pKF->mvKeysUn[i].angle;
    }
    Bytes blocked(n);
    for (int k = 0; k < n; k++) blocked[k] = CurrentFrame.mvpMapPoints[k] != NULL;                      // :1541-1542
    vector<int32_t> qk;
    const int nmatches = W.run(c, front(c, CurrentFrame), n, blocked.data(), nullptr, ORBdist,
                               ORBB200_WB_BLOCK | (mbCheckOrientation ? ORBB200_WB_ORI : 0), qk);
    for (int k = 0; k < n; k++)
        if (qk[k] >= 0) CurrentFrame.mvpMapPoints[k] = vpMPs[qk[k]];                                    // only NULL slots can be taken
    return nmatches;
}

int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*> &vpPoints, vector<MapPoint*> &vpMatched, int th)
{
    orbb200_ctx* c = orbb200_host::ThreadContext(pKF->mvScaleFactors);
    const float &fx = pKF->fx;
    const float &fy = pKF->fy;
    const float &cx = pKF->cx;
    const float &cy = pKF->cy;
    const Sim3Pose S(Scw);

    set<MapPoint*> spAlreadyFound(vpMatched.begin(), vpMatched.end());
    spAlreadyFound.erase(static_cast<MapPoint*>(NULL));

    const int nq = (int)vpPoints.size(), n = pKF->N;
    if (nq == 0 || n == 0) return 0;
    Windows W(nq);
    for (int iMP = 0; iMP < nq; iMP++)
    {
        MapPoint* pMP = vpPoints[iMP];
        if(pMP->isBad() || spAlreadyFound.count(pMP))
            continue;
        cv::Mat p3Dw = pMP->GetWorldPos();
        Pixel px;
        if (!pinhole(S.Rcw*p3Dw+S.tcw, fx, fy, cx, cy, false, px) || !pKF->IsInImage(px.u,px.v)) continue;
        const float u = px.u, v = px.v;
        float dist;
        const int nPredictedLevel = gate_and_level(pMP, p3Dw, S.Ow, pKF, dist);
        if (nPredictedLevel < 0) continue;
        const float radius = th*pKF->mvScaleFactors[nPredictedLevel];
        W.set(iMP, u, v, radius, nPredictedLevel-1, nPredictedLevel, pMP->GetDescriptor());
    }
    Bytes blocked(n);
    for (int k = 0; k < n; k++) blocked[k] = vpMatched[k] != NULL;                                       // :375-376
    vector<int32_t> qk;
    const int nmatches = W.run(c, front(c, pKF), n, blocked.data(), nullptr, TH_LOW, ORBB200_WB_BLOCK, qk);
    for (int k = 0; k < n; k++)
        if (qk[k] >= 0) vpMatched[k] = vpPoints[qk[k]];
    return nmatches;
}

int ORBmatcher::SearchByBoW(KeyFrame* pKF,Frame &F, vector<MapPoint*> &vpMapPointMatches)
{
    orbb200_ctx* c = orbb200_host::ThreadContext(F.mvScaleFactors);
    const vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
    vpMapPointMatches = vector<MapPoint*>(F.N,static_cast<MapPoint*>(NULL));
    const int n1 = pKF->N;
    if (n1 == 0 || F.N == 0) return 0;
    Bytes valid1(n1), tmp;
    vector<float> angle1(n1);
    for (int i = 0; i < n1; i++)
    {
        valid1[i] = vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad();                  // :191-197
        angle1[i] = pKF->mvKeysUn[i].angle;
    }
    const FeatCSR f1(pKF->mFeatVec), f2(F.mFeatVec);
    vector<int32_t> out(F.N, -1);
    int nmatches = 0;
    check(orbb200_search_by_bow(c, rows32(pKF->mDescriptors, n1, tmp), angle1.data(), valid1.data(), n1, front(c, F), nullptr,
                                f1.node.data(), f1.ptr.data(), f1.idx.data(), f1.n(), f2.node.data(), f2.ptr.data(), f2.idx.data(), f2.n(),
                                mfNNratio, mbCheckOrientation ? 1 : 0, 0, out.data(), &nmatches), "orbb200_search_by_bow");
    for (int i = 0; i < F.N; i++)
        if (out[i] >= 0) vpMapPointMatches[i] = vpMapPointsKF[out[i]];
    return nmatches;
}

int ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint *> &vpMatches12)
{
    orbb200_ctx* c = orbb200_host::ThreadContext(pKF1->mvScaleFactors);
    const vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
    const vector<MapPoint*> vpMapPoints2 = pKF2->GetMapPointMatches();
    vpMatches12 = vector<MapPoint*>(vpMapPoints1.size(),static_cast<MapPoint*>(NULL));
    const int n1 = pKF1->N, n2 = pKF2->N;
    if (n1 == 0 || n2 == 0) return 0;
    Bytes valid1(n1), valid2(n2), tmp;
    vector<float> angle1(n1);
    for (int i = 0; i < n1; i++)
    {
        valid1[i] = vpMapPoints1[i] && !vpMapPoints1[i]->isBad();                    // :558-562
        angle1[i] = pKF1->mvKeysUn[i].angle;
    }
    for (int i = 0; i < n2; i++) valid2[i] = vpMapPoints2[i] && !vpMapPoints2[i]->isBad();              // :574-580
    const FeatCSR f1(pKF1->mFeatVec), f2(pKF2->mFeatVec);
    vector<int32_t> out(n1, -1);
    int nmatches = 0;
    check(orbb200_search_by_bow(c, rows32(pKF1->mDescriptors, n1, tmp), angle1.data(), valid1.data(), n1, front(c, pKF2), valid2.data(),
                                f1.node.data(), f1.ptr.data(), f1.idx.data(), f1.n(), f2.node.data(), f2.ptr.data(), f2.idx.data(), f2.n(),
                                mfNNratio, mbCheckOrientation ? 1 : 0, 1, out.data(), &nmatches), "orbb200_search_by_bow");
    for (int i = 0; i < n1; i++)
        if (out[i] >= 0) vpMatches12[i] = vpMapPoints2[out[i]];
    return nmatches;
}

int ORBmatcher::SearchForInitialization(Frame &F1, Frame &F2, vector<cv::Point2f> &vbPrevMatched, vector<int> &vnMatches12, int windowSize)
{
    orbb200_ctx* c = orbb200_host::ThreadContext(F2.mvScaleFactors);
    const int n1 = (int)F1.mvKeysUn.size();
    vnMatches12 = vector<int>(n1,-1);
    if (n1 == 0 || F2.N == 0) return 0;
    Bytes tmp;
    static_assert(sizeof(cv::Point2f) == 8, "cv::Point2f is two packed floats");
    vector<int32_t> m12(n1, -1);
    int nmatches = 0;
    check(orbb200_search_for_initialization(c, (const orbb200_kp_t*)F1.mvKeysUn.data(), rows32(F1.mDescriptors, n1, tmp), n1, front(c, F2),
                                            reinterpret_cast<float*>(vbPrevMatched.data()), windowSize, mfNNratio, mbCheckOrientation ? 1 : 0,
                                            m12.data(), &nmatches), "orbb200_search_for_initialization");
    for (int i = 0; i < n1; i++) vnMatches12[i] = m12[i];
    return nmatches;
}

int ORBmatcher::SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, cv::Mat F12,
                                       vector<pair<size_t, size_t> > &vMatchedPairs, const bool bOnlyStereo)
{
    orbb200_ctx* c = orbb200_host::ThreadContext(pKF2->mvScaleFactors);
    //Compute epipole in second image (:663-670)
    cv::Mat Cw = pKF1->GetCameraCenter();
    cv::Mat R2w = pKF2->GetRotation();
    cv::Mat t2w = pKF2->GetTranslation();
    cv::Mat C2 = R2w*Cw+t2w;
    const float invz = 1.0f/C2.at<float>(2);
    const float ex =pKF2->fx*C2.at<float>(0)*invz+pKF2->cx;
    const float ey =pKF2->fy*C2.at<float>(1)*invz+pKF2->cy;

    vMatchedPairs.clear();
    const int n1 = pKF1->N, n2 = pKF2->N;
    if (n1 == 0 || n2 == 0) return 0;
    Bytes has1(n1), has2(n2), tmp1, tmp2;
    for (int i = 0; i < n1; i++) has1[i] = pKF1->GetMapPoint(i) != NULL;             // "If there is already a MapPoint skip" (:699-701)
    for (int i = 0; i < n2; i++) has2[i] = pKF2->GetMapPoint(i) != NULL;             // (:727-729)
    const FeatCSR f1(pKF1->mFeatVec), f2(pKF2->mFeatVec);
    float F[9];
    for (int r = 0; r < 3; r++)
        for (int k = 0; k < 3; k++) F[3 * r + k] = F12.at<float>(r, k);
    vector<int32_t> pairs(2 * (size_t)n1);
    int npairs = 0;
    check(orbb200_search_for_triangulation(c, (const orbb200_kp_t*)pKF1->mvKeysUn.data(), rows32(pKF1->mDescriptors, n1, tmp1), pKF1->mvuRight.data(), has1.data(), n1,
                                           (const orbb200_kp_t*)pKF2->mvKeysUn.data(), rows32(pKF2->mDescriptors, n2, tmp2), pKF2->mvuRight.data(), has2.data(), n2,
                                           f1.node.data(), f1.ptr.data(), f1.idx.data(), f1.n(), f2.node.data(), f2.ptr.data(), f2.idx.data(), f2.n(),
                                           F, ex, ey, pKF2->mvScaleFactors.data(), pKF2->mvLevelSigma2.data(),
                                           bOnlyStereo ? 1 : 0, mbCheckOrientation ? 1 : 0, pairs.data(), &npairs), "orbb200_search_for_triangulation");
    vMatchedPairs.reserve(npairs);
    for (int i = 0; i < npairs; i++)
        vMatchedPairs.push_back(make_pair((size_t)pairs[2 * i], (size_t)pairs[2 * i + 1]));    // ascending idx1, as :814-819
    return npairs;
}

int ORBmatcher::SearchBySim3(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*> &vpMatches12,
                             const float &s12, const cv::Mat &R12, const cv::Mat &t12, const float th)
{
    orbb200_ctx* c = orbb200_host::ThreadContext(pKF1->mvScaleFactors);
    const float &fx = pKF1->fx;
    const float &fy = pKF1->fy;
    const float &cx = pKF1->cx;
    const float &cy = pKF1->cy;

    cv::Mat R1w = pKF1->GetRotation();
    cv::Mat t1w = pKF1->GetTranslation();
    cv::Mat R2w = pKF2->GetRotation();
    cv::Mat t2w = pKF2->GetTranslation();
    cv::Mat sR12 = s12*R12;
    cv::Mat sR21 = (1.0/s12)*R12.t();
    cv::Mat t21 = -sR21*t12;

    const vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
    const int N1 = vpMapPoints1.size();
    const vector<MapPoint*> vpMapPoints2 = pKF2->GetMapPointMatches();
    const int N2 = vpMapPoints2.size();

    vector<bool> vbAlreadyMatched1(N1,false);
    vector<bool> vbAlreadyMatched2(N2,false);
    for(int i=0; i<N1; i++)
    {
        MapPoint* pMP = vpMatches12[i];
        if(pMP)
        {
            vbAlreadyMatched1[i]=true;
            int idx2 = pMP->GetIndexInKeyFrame(pKF2);
            if(idx2>=0 && idx2<N2)
                vbAlreadyMatched2[idx2]=true;
        }
    }

    // one direction: the map points of keyframe A, moved into camera B by [sR | t] after A's own pose, searched in B's grid
    struct Direction
    {
        static void queries(Windows& W, const vector<MapPoint*>& pts, const vector<bool>& already, const cv::Mat& RAw, const cv::Mat& tAw,
                            const cv::Mat& sRBA, const cv::Mat& tBA, KeyFrame* pKFB, float fx, float fy, float cx, float cy, float th)
        {
            for (int i = 0; i < (int)pts.size(); i++)
            {
                MapPoint* pMP = pts[i];
                if(!pMP || already[i])
                    continue;
                if(pMP->isBad())
                    continue;
                cv::Mat p3Dw = pMP->GetWorldPos();
                cv::Mat p3DcA = RAw*p3Dw + tAw;
                cv::Mat p3DcB = sRBA*p3DcA + tBA;
                Pixel px;
                if (!pinhole(p3DcB, fx, fy, cx, cy, true, px) || !pKFB->IsInImage(px.u,px.v)) continue;
                const float u = px.u, v = px.v;
                const float maxDistance = pMP->GetMaxDistanceInvariance();
                const float minDistance = pMP->GetMinDistanceInvariance();
                const float dist3D = cv::norm(p3DcB);
                if(dist3D<minDistance || dist3D>maxDistance )
                    continue;
                const int nPredictedLevel = pMP->PredictScale(dist3D,pKFB);
                const float radius = th*pKFB->mvScaleFactors[nPredictedLevel];
                W.set(i, u, v, radius, nPredictedLevel-1, nPredictedLevel, pMP->GetDescriptor());
            }
        }
    };
    Windows W1(N1), W2(N2);
    Direction::queries(W1, vpMapPoints1, vbAlreadyMatched1, R1w, t1w, sR21, t21, pKF2, fx, fy, cx, cy, th);     // :1148-1225
    Direction::queries(W2, vpMapPoints2, vbAlreadyMatched2, R2w, t2w, sR12, t12, pKF1, fx, fy, cx, cy, th);     // :1228-1305
    vector<int32_t> unused;
    W1.run(c, front(c, pKF2), pKF2->N, nullptr, nullptr, TH_HIGH, 0, unused);
    W2.run(c, front(c, pKF1), pKF1->N, nullptr, nullptr, TH_HIGH, 0, unused);

    // Check agreement (:1307-1323)
    int nFound = 0;
    for(int i1=0; i1<N1; i1++)
    {
        const int idx2 = W1.bestIdx[i1];
        if(idx2>=0 && W2.bestIdx[idx2]==i1)
        {
            vpMatches12[i1] = vpMapPoints2[idx2];
            nFound++;
        }
    }
    return nFound;
}

int ORBmatcher::Fuse(KeyFrame *pKF, const vector<MapPoint *> &vpMapPoints, const float th)
{
    orbb200_ctx* c = orbb200_host::ThreadContext(pKF->mvScaleFactors);
    cv::Mat Rcw = pKF->GetRotation();
    cv::Mat tcw = pKF->GetTranslation();
    const float &fx = pKF->fx;
    const float &fy = pKF->fy;
    const float &cx = pKF->cx;
    const float &cy = pKF->cy;
    const float &bf = pKF->mbf;
    cv::Mat Ow = pKF->GetCameraCenter();

    const int nMPs = vpMapPoints.size();
    if (nMPs == 0 || pKF->N == 0) return 0;
    // the scan of a point depends on nothing the loop changes (positions, descriptors, the keyframe's keypoints): all scans in one call
    Windows W(nMPs);
    for(int i=0; i<nMPs; i++)
    {
        MapPoint* pMP = vpMapPoints[i];
        if(!pMP)
            continue;
        cv::Mat p3Dw = pMP->GetWorldPos();
        Pixel px;
        if (!pinhole(Rcw*p3Dw + tcw, fx, fy, cx, cy, false, px) || !pKF->IsInImage(px.u,px.v)) continue;
        const float u = px.u, v = px.v;
        const float ur = u-bf*px.invz;
        float dist3D;
        const int nPredictedLevel = gate_and_level(pMP, p3Dw, Ow, pKF, dist3D);
        if (nPredictedLevel < 0) continue;
        const float radius = th*pKF->mvScaleFactors[nPredictedLevel];
        W.set(i, u, v, radius, nPredictedLevel-1, nPredictedLevel, pMP->GetDescriptor());
        W.aux[i] = ur;
    }
    vector<int32_t> unused;
    W.run(c, front(c, pKF), pKF->N, nullptr, pKF->mvInvLevelSigma2.data(), TH_LOW, ORBB200_WB_CHI2, unused);

    // the map mutations, replayed in list order against the live state (:842-850, :951-971): an earlier Replace / AddObservation
    // changes isBad() / IsInKeyFrame() / GetMapPoint() for later entries
    int nFused=0;
    for(int i=0; i<nMPs; i++)
    {
        MapPoint* pMP = vpMapPoints[i];
        if(!pMP)
            continue;
        if(pMP->isBad() || pMP->IsInKeyFrame(pKF))
            continue;
        const int bestIdx = W.bestIdx[i];
        if (bestIdx < 0) continue;
        MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx);
        if(pMPinKF)
        {
            if(!pMPinKF->isBad())
            {
                if(pMPinKF->Observations()>pMP->Observations())
                    pMP->Replace(pMPinKF);
                else
                    pMPinKF->Replace(pMP);
            }
        }
        else
        {
            pMP->AddObservation(pKF,bestIdx);
            pKF->AddMapPoint(pMP,bestIdx);
        }
        nFused++;
    }
    return nFused;
}

int ORBmatcher::Fuse(KeyFrame *pKF, cv::Mat Scw, const vector<MapPoint *> &vpPoints, float th, vector<MapPoint *> &vpReplacePoint)
{
    orbb200_ctx* c = orbb200_host::ThreadContext(pKF->mvScaleFactors);
    const float &fx = pKF->fx;
    const float &fy = pKF->fy;
    const float &cx = pKF->cx;
    const float &cy = pKF->cy;
    const Sim3Pose S(Scw);

    const set<MapPoint*> spAlreadyFound = pKF->GetMapPoints();

    const int nPoints = vpPoints.size();
    if (nPoints == 0 || pKF->N == 0) return 0;
    Windows W(nPoints);
    for(int iMP=0; iMP<nPoints; iMP++)
    {
        MapPoint* pMP = vpPoints[iMP];
        if(spAlreadyFound.count(pMP))
            continue;
        cv::Mat p3Dw = pMP->GetWorldPos();
        Pixel px;
        if (!pinhole(S.Rcw*p3Dw+S.tcw, fx, fy, cx, cy, true, px) || !pKF->IsInImage(px.u,px.v)) continue;
        const float u = px.u, v = px.v;
        float dist3D;
        const int nPredictedLevel = gate_and_level(pMP, p3Dw, S.Ow, pKF, dist3D);
        if (nPredictedLevel < 0) continue;
        const float radius = th*pKF->mvScaleFactors[nPredictedLevel];
        W.set(iMP, u, v, radius, nPredictedLevel-1, nPredictedLevel, pMP->GetDescriptor());
    }
    vector<int32_t> unused;
    W.run(c, front(c, pKF), pKF->N, nullptr, nullptr, TH_LOW, 0, unused);

    int nFused=0;
    for(int iMP=0; iMP<nPoints; iMP++)                                               // replay, :1004-1006 and :1081-1096
    {
        MapPoint* pMP = vpPoints[iMP];
        if(pMP->isBad() || spAlreadyFound.count(pMP))
            continue;
        const int bestIdx = W.bestIdx[iMP];
        if (bestIdx < 0) continue;
        MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx);
        if(pMPinKF)
        {
            if(!pMPinKF->isBad())
                vpReplacePoint[iMP] = pMPinKF;
        }
        else
        {
            pMP->AddObservation(pKF,bestIdx);
            pKF->AddMapPoint(pMP,bestIdx);
        }
        nFused++;
    }
    return nFused;
}

int ORBmatcher::BirdviewMatch(Frame &F1, Frame &F2, vector<int> &vnMatches12, vector<cv::Point2f> &vPrevMatched, int windowSize)
{
    orbb200_ctx* c = orbb200_host::ThreadContext(F2.mvScaleFactors);
    const int n1 = (int)F1.mvKeysBird.size();
    vnMatches12 = vector<int>(n1,-1);
    if (n1 == 0 || F2.mvKeysBird.empty()) return 0;
    Bytes tmp;
    vector<int32_t> m12(n1, -1);
    int nmatches = 0;
    check(orbb200_birdview_match(c, (const orbb200_kp_t*)F1.mvKeysBird.data(), rows32(F1.mDescriptorsBird, n1, tmp), n1, bird(c, F2),
                                 reinterpret_cast<float*>(vPrevMatched.data()), windowSize, mfNNratio, mbCheckOrientation ? 1 : 0, m12.data(), &nmatches),
          "orbb200_birdview_match");
    for (int i = 0; i < n1; i++) vnMatches12[i] = m12[i];
    return nmatches;
}

int ORBmatcher::BirdviewMatch(const Frame &F1, const Frame &F2, vector<int> &vnMatches12, int windowSize)
{
    orbb200_ctx* c = orbb200_host::ThreadContext(F2.mvScaleFactors);
    const int n1 = (int)F1.mvKeysBird.size();
    vnMatches12 = vector<int>(n1, -1);
    if (n1 == 0 || F2.mvKeysBird.empty()) return 0;
    Bytes tmp;
    vector<int32_t> m12(n1, -1);
    int nmatches = 0;
    check(orbb200_birdview_match(c, (const orbb200_kp_t*)F1.mvKeysBird.data(), rows32(F1.mDescriptorsBird, n1, tmp), n1, bird(c, F2),
                                 nullptr, windowSize, mfNNratio, mbCheckOrientation ? 1 : 0, m12.data(), &nmatches), "orbb200_birdview_match");
    for (int i = 0; i < n1; i++) vnMatches12[i] = m12[i];
    return nmatches;
}

int ORBmatcher::SearchByMatchBird(Frame &CurrentFrame, const Frame &LastFrame, const int windowSize)
{
    int nmatches = 0;
    std::vector<int> vnMatches12;
    BirdviewMatch(LastFrame,CurrentFrame,vnMatches12,windowSize);

    for(int k=0;k<(int)LastFrame.mvKeysBird.size();k++)                              // the copy loop of :1907-1918
    {
        int idx2 = vnMatches12[k];
        if(idx2<0)
            continue;
        MapPointBird *pMPBird = LastFrame.mvpMapPointsBird[k];
        if(pMPBird)
        {
            CurrentFrame.mvpMapPointsBird[idx2] = pMPBird;
            nmatches++;
        }
    }

    return nmatches;
}

int ORBmatcher::SearchByProjectionBird(Frame &F, const vector<MapPointBird*> &vpMapPointsBird, const float r)
{
    orbb200_ctx* c = orbb200_host::ThreadContext(F.mvScaleFactors);
    const int nq = (int)vpMapPointsBird.size(), n = (int)F.mvKeysBird.size();
    if (nq == 0 || n == 0) return 0;
    cv::Mat Tbw = Frame::Tbc*F.mTcw;
    Bytes valid(nq, 0), obs(nq, 0), desc((size_t)nq * 32), blocked(n);
    vector<float> qx(nq), qy(nq);
    for (int iMP = 0; iMP < nq; iMP++)
    {
        MapPointBird* pMPBird = vpMapPointsBird[iMP];
        if(pMPBird->mnLastFrameSeen==F.mnId)
            continue;
        cv::Mat worldPos = pMPBird->GetWorldPos();
        cv::Mat localPos = Tbw.rowRange(0,3).colRange(0,3)*worldPos+Tbw.rowRange(0,3).col(3);
        if(fabs(localPos.at<float>(2))>0.2)
            continue;
        cv::Point2f pt = Frame::ProjectXYZ2Birdview(cv::Point3f(localPos.at<float>(0),localPos.at<float>(1),localPos.at<float>(2)));
        if(pt.x<0||pt.x>=Frame::birdviewCols||pt.y<0||pt.y>=Frame::birdviewRows)
            continue;
        valid[iMP] = 1; qx[iMP] = pt.x; qy[iMP] = pt.y;
        obs[iMP] = pMPBird->Observations() > 0;
        const cv::Mat d = pMPBird->GetDescriptor();
        memcpy(&desc[32 * (size_t)iMP], d.ptr(0), 32);
    }
    for (int k = 0; k < n; k++)
        blocked[k] = F.mvpMapPointsBird[k] && F.mvpMapPointsBird[k]->Observations() > 0;                // :1963-1965
    vector<int32_t> qk(n, -1);
    int nmatches = 0;
    check(orbb200_search_by_projection_bird(c, bird(c, F), nq, valid.data(), qx.data(), qy.data(), desc.data(), obs.data(), blocked.data(), r, mfNNratio,
                                            qk.data(), &nmatches), "orbb200_search_by_projection_bird");
    for (int k = 0; k < n; k++)
        if (qk[k] >= 0) F.mvpMapPointsBird[k] = vpMapPointsBird[qk[k]];
    return nmatches;
}

int ORBmatcher::SearchByMatchBird(KeyFrame *pKF, Frame &F, std::vector<MapPointBird*> &vpMapPointMatchesBird, const float r)
{
    orbb200_ctx* c = orbb200_host::ThreadContext(F.mvScaleFactors);
    const vector<MapPointBird*> vpMapPointsBirdKF = pKF->GetMapPointMatchesBird();
    const int nk = (int)vpMapPointsBirdKF.size(), n = (int)F.mvKeysBird.size();
    vpMapPointMatchesBird = vector<MapPointBird*>(n,static_cast<MapPointBird*>(NULL));
    if (nk == 0 || n == 0) return 0;
    Bytes has(nk, 0), desc((size_t)nk * 32, 0);
    for (int k = 0; k < nk; k++)
    {
        MapPointBird *pMPBird = vpMapPointsBirdKF[k];
        if(!pMPBird)
            continue;
        has[k] = 1;
        const cv::Mat d1 = pMPBird->GetDescriptor();
        memcpy(&desc[32 * (size_t)k], d1.ptr(0), 32);
    }
    vector<int32_t> mpOfKp(n, -1);
    int nmatches = 0;
    check(orbb200_search_by_match_bird_kf(c, (const orbb200_kp_t*)pKF->mvKeysBird.data(), has.data(), desc.data(), nk, bird(c, F), r, mfNNratio,
                                          mbCheckOrientation ? 1 : 0, mpOfKp.data(), &nmatches), "orbb200_search_by_match_bird_kf");
    for (int i = 0; i < n; i++)
        if (mpOfKp[i] >= 0) vpMapPointMatchesBird[i] = vpMapPointsBirdKF[mpOfKp[i]];
    return nmatches;
}

} // namespace ORB_SLAM2
