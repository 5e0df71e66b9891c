// Bodies of ORB_SLAM2::ORBmatcher methods over the C ABI of liborbb200.so (include/orbb200.h).
//
// How a maintainer uses this file: add it to the ORB_SLAM2 library sources (CMakeLists.txt:49-75) and delete the bodies
// of the same methods from src/ORBmatcher.cc -- the class declaration in include/ORBmatcher.h, and therefore every call
// site in Tracking.cc / LocalMapping.cc / LoopClosing.cc, stays untouched.  Each method flattens the object graph into
// arrays, makes one ABI call and applies the result to the members the reference loop mutates, in the same order.
// Here (no reference tree, no OpenCV SDK) it is compiled against compat/*.h, which declare the same members, and
// driven by matcher_driver.cpp; tests/test_gpu_parity.py::test_cpp_matcher_shim checks it against the oracle.
//
//   ORBmatcher::DescriptorDistance                         src/ORBmatcher.cc:1647-1663   (host: one pair is not GPU work)
//   ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th)      :45-129    -> orbb200_search_by_projection
//   ORBmatcher::SearchForTriangulation(KF1, KF2, F12, vMatchedPairs, bOnlyStereo)    :657-823   -> orbb200_search_for_triangulation
//   ORBmatcher::BirdviewMatch(const Frame&, const Frame&, vnMatches12, windowSize)  :1788-1899 -> orbb200_birdview_match
//   ORBmatcher::SearchByMatchBird(Frame&, const Frame&, windowSize)     :1901-1921 -> the same + the reference's copy loop
//
// The remaining methods follow the same pattern; their adapters are listed in INTEGRATION.md section 3.
#include "ORBmatcher.h"

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "orbb200_host.h"

using namespace std;

namespace ORB_SLAM2
{

// Frames are uploaded per call here.  In the reference tree, give Frame two members (orbb200_frame* mDevFrame,
// mDevFrameBird) filled at the end of Frame::Frame and pass those instead.
namespace
{
struct DevFrame
{
    orbb200_frame* h = nullptr;
    DevFrame(const std::vector<cv::KeyPoint>& keys, const cv::Mat& desc, const std::vector<float>* uRight,
             float minX, float minY, float invW, float invH)
    {
        std::vector<uint8_t> d(keys.size() * 32);
        for (size_t i = 0; i < keys.size(); i++) memcpy(&d[32 * i], desc.ptr((int)i), 32);      // rows may be strided
        orbb200_host::check(orbb200_frame_upload(orbb200_host::ThreadContext(), &h, (const orbb200_kp_t*)keys.data(), d.data(),
                                                 uRight && !uRight->empty() ? uRight->data() : nullptr, (int)keys.size(),
                                                 minX, minY, invW, invH), "orbb200_frame_upload");
    }
    ~DevFrame() { if (h) orbb200_frame_free(h); }
};
}  // namespace

#ifndef ORBB200_HAVE_ORBSLAM          // inside the reference tree src/ORBmatcher.cc keeps these definitions
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
#endif

int ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th)
{
    const int nq = (int)vpMapPoints.size();
    vector<uint8_t> valid(nq), obs(nq), desc((size_t)nq * 32), blocked(F.N);
    vector<float> u(nq), v(nq), uR(nq), vc(nq);
    vector<int32_t> lvl(nq);
    for (int i = 0; i < nq; i++)
    {
        MapPoint* pMP = vpMapPoints[i];
        valid[i] = pMP->mbTrackInView && !pMP->isBad();                              // :53-58
        u[i] = pMP->mTrackProjX; v[i] = pMP->mTrackProjY; uR[i] = pMP->mTrackProjXR;
        lvl[i] = pMP->mnTrackScaleLevel; vc[i] = pMP->mTrackViewCos;
        obs[i] = pMP->Observations() > 0;                                            // what a later :87-89 test will see
        const cv::Mat d = pMP->GetDescriptor();
        memcpy(&desc[32 * (size_t)i], d.ptr(0), 32);
    }
    for (int k = 0; k < F.N; k++)
        blocked[k] = F.mvpMapPoints[k] && F.mvpMapPoints[k]->Observations() > 0;     // :87-89
    DevFrame dF(F.mvKeysUn, F.mDescriptors, &F.mvuRight, Frame::mnMinX, Frame::mnMinY, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv);
    vector<int32_t> bi(nq), bd(nq), qk(F.N);
    int nmatches = 0;
    // the scale factors the window radius uses (:72) are the context's own tables (same float chain as mvScaleFactors)
    orbb200_host::check(orbb200_search_by_projection(orbb200_host::ThreadContext(), dF.h, nq, valid.data(), u.data(), v.data(), uR.data(),
                                                     lvl.data(), vc.data(), desc.data(), obs.data(), blocked.data(), th, mfNNratio,
                                                     bi.data(), bd.data(), qk.data(), &nmatches), "orbb200_search_by_projection");
    for (int i = 0; i < nq; i++)
        if (bi[i] >= 0) F.mvpMapPoints[bi[i]] = vpMapPoints[i];                      // in query order == the loop's writes (:123)
    return nmatches;
}

int ORBmatcher::SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, cv::Mat F12,
                                       vector<pair<size_t, size_t> > &vMatchedPairs, const bool bOnlyStereo)
{
    //Compute epipole in second image (the reference's own lines, :663-670)
    cv::Mat Cw = pKF1->GetCameraCenter();
    cv::Mat R2w = pKF2->GetRotation();
    cv::Mat t2w = pKF2->GetTranslation();
    cv::Mat C2 = R2w*Cw+t2w;
    const float invz = 1.0f/C2.at<float>(2);
    const float ex =pKF2->fx*C2.at<float>(0)*invz+pKF2->cx;
    const float ey =pKF2->fy*C2.at<float>(1)*invz+pKF2->cy;

    // flatten both keyframes; DBoW2::FeatureVector is a std::map, so its iteration order (ascending node id) is the
    // order of the reference's merge loop (:687-780)
    struct Flat
    {
        vector<uint8_t> desc, hasMP;
        vector<int32_t> node, ptr, idx;
        Flat(KeyFrame* pKF)
        {
            const int n = pKF->N;
            desc.resize((size_t)n * 32); hasMP.resize(n);
            for (int i = 0; i < n; i++)
            {
                memcpy(&desc[32 * (size_t)i], pKF->mDescriptors.ptr(i), 32);
                hasMP[i] = pKF->GetMapPoint(i) != NULL;                                  // "If there is already a MapPoint skip" (:699-701, :727-729)
            }
            ptr.push_back(0);
            for (DBoW2::FeatureVector::const_iterator it = pKF->mFeatVec.begin(); it != pKF->mFeatVec.end(); it++)
            {
                node.push_back((int32_t)it->first);
                for (size_t k = 0; k < it->second.size(); k++) idx.push_back((int32_t)it->second[k]);
                ptr.push_back((int32_t)idx.size());
            }
        }
    };
    Flat f1(pKF1), f2(pKF2);
    float F[9];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) F[3 * r + c] = F12.at<float>(r, c);
    vector<int32_t> pairs(2 * (size_t)std::max(pKF1->N, 1));
    int npairs = 0;
    orbb200_host::check(orbb200_search_for_triangulation(orbb200_host::ThreadContext(),
                            (const orbb200_kp_t*)pKF1->mvKeysUn.data(), f1.desc.data(), pKF1->mvuRight.data(), f1.hasMP.data(), pKF1->N,
                            (const orbb200_kp_t*)pKF2->mvKeysUn.data(), f2.desc.data(), pKF2->mvuRight.data(), f2.hasMP.data(), pKF2->N,
                            f1.node.data(), f1.ptr.data(), f1.idx.data(), (int)f1.node.size(),
                            f2.node.data(), f2.ptr.data(), f2.idx.data(), (int)f2.node.size(),
                            F, ex, ey, pKF2->mvScaleFactors.data(), pKF2->mvLevelSigma2.data(),
                            bOnlyStereo ? 1 : 0, mbCheckOrientation ? 1 : 0, pairs.data(), &npairs), "orbb200_search_for_triangulation");
    vMatchedPairs.clear();
    vMatchedPairs.reserve(npairs);
    for (int i = 0; i < npairs; i++)
        vMatchedPairs.push_back(make_pair((size_t)pairs[2 * i], (size_t)pairs[2 * i + 1]));    // ascending idx1, as :814-819
    return npairs;
}

int ORBmatcher::BirdviewMatch(const Frame &F1, const Frame &F2, vector<int> &vnMatches12, int windowSize)
{
    const int n1 = (int)F1.mvKeysBird.size();
    vnMatches12 = vector<int>(n1, -1);
    if (n1 == 0) return 0;
    vector<uint8_t> d1((size_t)n1 * 32);
    for (int i = 0; i < n1; i++) memcpy(&d1[32 * (size_t)i], F1.mDescriptorsBird.ptr(i), 32);
    DevFrame dF2(F2.mvKeysBird, F2.mDescriptorsBird, nullptr, 0.f, 0.f, Frame::mfGridElementWidthInvBirdview, Frame::mfGridElementHeightInvBirdview);
    vector<int32_t> m12(n1, -1);
    int nmatches = 0;
    orbb200_host::check(orbb200_birdview_match(orbb200_host::ThreadContext(), (const orbb200_kp_t*)F1.mvKeysBird.data(), d1.data(), n1, dF2.h,
                                               nullptr, windowSize, mfNNratio, mbCheckOrientation ? 1 : 0, m12.data(), &nmatches),
                        "orbb200_birdview_match");
    for (int i = 0; i < n1; i++) vnMatches12[i] = m12[i];
    return nmatches;
}

int ORBmatcher::SearchByMatchBird(Frame &CurrentFrame, const Frame &LastFrame, const int windowSize)
{
    int nmatches = 0;
    std::vector<int> vnMatches12;
    BirdviewMatch(LastFrame,CurrentFrame,vnMatches12,windowSize);

    for(int k=0;k<(int)LastFrame.mvKeysBird.size();k++)                              // the reference's own loop, :1907-1918
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

} // namespace ORB_SLAM2
