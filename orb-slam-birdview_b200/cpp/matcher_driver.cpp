// Test driver for ORBmatcher_b200.cc: builds Frame / MapPoint / MapPointBird objects from a flat binary case file, calls
// the ORB_SLAM2::ORBmatcher methods exactly like Tracking does (src/Tracking.cc:1652-1672 SearchByProjection on the
// local map points, :1241 SearchByMatchBird, and BirdviewMatch directly) and dumps what they left in the objects.
//   matcher_driver <case.bin> <out.bin>
// case.bin: int32 {nF, nq, n1, n2, window}; float32 {th, nnratio, nnratioBird, minX, minY, invW, invH, invWbird, invHbird};
//   front frame: kps[nF] (28 B), desc[nF][32], uRight[nF] f32, kp_obs[nF] i32 (-1: no MapPoint on the keypoint, else its Observations())
//   map points:  inview[nq] u8, bad[nq] u8, u,v,uR,viewcos [nq] f32, level[nq] i32, obs[nq] i32, desc[nq][32]
//   birdview:    kps1[n1], desc1[n1][32], hasmp1[n1] u8, kps2[n2], desc2[n2][32]
// out.bin: int32 nmatches; int32 mp_of_kp[nF] (index of the map point now on the keypoint, -2: the pre-existing one, -1: none);
//          int32 nmBird; int32 vnMatches12[n1]; int32 nmSearchByMatchBird; int32 bird_mp_of_kp2[n2] (LastFrame keypoint index or -1);
//          int32 DescriptorDistance(desc1[0], desc2[0]); int32 nTri (-1: no keyframe section); int32 pairs[nTri][2]
// optional keyframe section of case.bin: see the reader below
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ORBmatcher.h"

using namespace ORB_SLAM2;


static std::vector<unsigned char> buf;
static size_t off = 0;
template <class T> static const T* take(size_t n) { const T* p = reinterpret_cast<const T*>(buf.data() + off); off += n * sizeof(T); if (off > buf.size()) { fprintf(stderr, "case file too short\n"); exit(2); } return p; }

static cv::Mat rows32(const unsigned char* d, int n)
{
    cv::Mat m(n > 0 ? n : 1, 32, CV_8U);
    if (n > 0) memcpy(m.data, d, (size_t)n * 32);
    return m;
}

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: matcher_driver case.bin out.bin\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    fseek(f, 0, SEEK_END); buf.resize(ftell(f)); fseek(f, 0, SEEK_SET);
    if (fread(buf.data(), 1, buf.size(), f) != buf.size()) return 2;
    fclose(f);
    const int* hi = take<int>(5);
    const int nF = hi[0], nq = hi[1], n1 = hi[2], n2 = hi[3], window = hi[4];
    const float* hf = take<float>(9);
    const float th = hf[0], nnratio = hf[1], nnratioBird = hf[2];
    Frame::mnMinX = hf[3]; Frame::mnMinY = hf[4]; Frame::mfGridElementWidthInv = hf[5]; Frame::mfGridElementHeightInv = hf[6];
    Frame::mfGridElementWidthInvBirdview = hf[7]; Frame::mfGridElementHeightInvBirdview = hf[8];

    // ---- front frame + local map points: Tracking::SearchLocalPoints -> matcher.SearchByProjection(mCurrentFrame, mvpLocalMapPoints, th)
    Frame F;
    F.N = nF;
    const cv::KeyPoint* kps = take<cv::KeyPoint>(nF);
    F.mvKeysUn.assign(kps, kps + nF);
    F.mDescriptors = rows32(take<unsigned char>((size_t)nF * 32), nF);
    const float* ur = take<float>(nF);
    F.mvuRight.assign(ur, ur + nF);
    const int* kpObs = take<int>(nF);
    std::vector<MapPoint> existing(nF);
    F.mvpMapPoints.assign(nF, static_cast<MapPoint*>(NULL));
    for (int i = 0; i < nF; i++)
        if (kpObs[i] >= 0) { existing[i].nObs = kpObs[i]; F.mvpMapPoints[i] = &existing[i]; }
    F.mvScaleFactors.resize(8);
    F.mvScaleFactors[0] = 1.0f;
    for (int i = 1; i < 8; i++) F.mvScaleFactors[i] = F.mvScaleFactors[i - 1] * 1.2f;

    const unsigned char* inview = take<unsigned char>(nq);
    const unsigned char* bad = take<unsigned char>(nq);
    const float* qu = take<float>(nq); const float* qv = take<float>(nq); const float* quR = take<float>(nq); const float* qvc = take<float>(nq);
    const int* qlvl = take<int>(nq); const int* qobs = take<int>(nq);
    const unsigned char* qdesc = take<unsigned char>((size_t)nq * 32);
    std::vector<MapPoint> mps(nq);
    std::vector<MapPoint*> vpLocalMapPoints(nq);
    for (int i = 0; i < nq; i++) {
        MapPoint& p = mps[i];
        p.mbTrackInView = inview[i] != 0; p.mbBad = bad[i] != 0;
        p.mTrackProjX = qu[i]; p.mTrackProjY = qv[i]; p.mTrackProjXR = quR[i]; p.mTrackViewCos = qvc[i];
        p.mnTrackScaleLevel = qlvl[i]; p.nObs = qobs[i];
        p.mDescriptor = rows32(qdesc + (size_t)i * 32, 1);
        vpLocalMapPoints[i] = &p;
    }
    ORBmatcher matcher(nnratio);
    const int nmatches = matcher.SearchByProjection(F, vpLocalMapPoints, th);

    // ---- birdview: BirdviewMatch(Last, Cur) and SearchByMatchBird(Cur, Last, window)
    Frame Last, Cur;
    const cv::KeyPoint* k1 = take<cv::KeyPoint>(n1);
    Last.mvKeysBird.assign(k1, k1 + n1);
    const unsigned char* d1 = take<unsigned char>((size_t)n1 * 32);
    Last.mDescriptorsBird = rows32(d1, n1);
    const unsigned char* hasmp1 = take<unsigned char>(n1);
    std::vector<MapPointBird> birds(n1);
    Last.mvpMapPointsBird.assign(n1, static_cast<MapPointBird*>(NULL));
    for (int i = 0; i < n1; i++) { birds[i].mnId = i; if (hasmp1[i]) Last.mvpMapPointsBird[i] = &birds[i]; }
    const cv::KeyPoint* k2 = take<cv::KeyPoint>(n2);
    Cur.mvKeysBird.assign(k2, k2 + n2);
    const unsigned char* d2 = take<unsigned char>((size_t)n2 * 32);
    Cur.mDescriptorsBird = rows32(d2, n2);
    Cur.mvpMapPointsBird.assign(n2, static_cast<MapPointBird*>(NULL));
    Last.mvScaleFactors = F.mvScaleFactors; Cur.mvScaleFactors = F.mvScaleFactors;
    ORBmatcher matcherBird(nnratioBird, true);                  // Tracking.cc:325-326: new ORBmatcher(0.99,true)
    std::vector<int> vnMatches12;
    const int nmBird = matcherBird.BirdviewMatch(Last, Cur, vnMatches12, window);
    const int nmSBM = matcherBird.SearchByMatchBird(Cur, Last, window);

    // ORBB200_MATCHER_TIME=N: wall time per call of the two per-frame searches through the adapters and the objects (the frame's device
    // copy is cached by content after the first call, as Tracking's three to five searches on one Frame find it)
    if (const char* e = getenv("ORBB200_MATCHER_TIME")) {
        const int reps = atoi(e) > 0 ? atoi(e) : 100;
        const std::vector<MapPoint*> initial = [&] { std::vector<MapPoint*> v(nF, static_cast<MapPoint*>(NULL)); for (int i = 0; i < nF; i++) if (kpObs[i] >= 0) v[i] = &existing[i]; return v; }();
        auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        double t0 = now();
        for (int r = 0; r < reps; r++) { F.mvpMapPoints = initial; matcher.SearchByProjection(F, vpLocalMapPoints, th); }
        const double msProj = (now() - t0) / reps;
        t0 = now();
        for (int r = 0; r < reps; r++) matcherBird.BirdviewMatch(Last, Cur, vnMatches12, window);
        const double msBird = (now() - t0) / reps;
        printf("TIMING {\"keypoints\": %d, \"map_points\": %d, \"SearchByProjection_ms\": %.4f, \"bird_keypoints\": %d, \"BirdviewMatch_ms\": %.4f}\n",
               nF, nq, msProj, n2, msBird);
    }

    // ---- two keyframes: LocalMapping::CreateNewMapPoints -> matcher.SearchForTriangulation(mpCurrentKeyFrame, pKF2, F12, vMatchedIndices, false)
    std::vector<std::pair<size_t, size_t> > vMatchedIndices;
    int nTri = -1;
    if (off < buf.size()) {
        const int* ti = take<int>(5);
        const int m1 = ti[0], m2 = ti[1], nn1 = ti[2], nn2 = ti[3], onlyStereo = ti[4];
        const float* tf = take<float>(4 + 3 + 3 + 9 + 9);     // fx fy cx cy | Ow(KF1) | tcw(KF2) | Rcw(KF2) | F12
        KeyFrame KF1, KF2;
        std::vector<MapPoint> dummy(1);
        KeyFrame* kfs[2] = {&KF1, &KF2};
        const int ms[2] = {m1, m2};
        for (int s = 0; s < 2; s++) {
            KeyFrame& K = *kfs[s];
            K.N = ms[s];
            const cv::KeyPoint* kk = take<cv::KeyPoint>(K.N);
            K.mvKeysUn.assign(kk, kk + K.N);
            K.mDescriptors = rows32(take<unsigned char>((size_t)K.N * 32), K.N);
            const float* u = take<float>(K.N);
            K.mvuRight.assign(u, u + K.N);
            const unsigned char* has = take<unsigned char>(K.N);
            K.mvpMapPoints.assign(K.N, static_cast<MapPoint*>(NULL));
            for (int i = 0; i < K.N; i++) if (has[i]) K.mvpMapPoints[i] = &dummy[0];
            K.fx = tf[0]; K.fy = tf[1]; K.cx = tf[2]; K.cy = tf[3];
        }
        const int nns[2] = {nn1, nn2};
        for (int s = 0; s < 2; s++) {
            const int* node = take<int>(nns[s]);
            const int* ptr = take<int>(nns[s] + 1);
            const int* idx = take<int>(ptr[nns[s]]);
            for (int k = 0; k < nns[s]; k++)
                for (int j = ptr[k]; j < ptr[k + 1]; j++) kfs[s]->mFeatVec[(unsigned)node[k]].push_back((unsigned)idx[j]);
        }
        const float* sf2 = take<float>(8);
        const float* ls2 = take<float>(8);
        KF2.mvScaleFactors.assign(sf2, sf2 + 8);
        KF2.mvLevelSigma2.assign(ls2, ls2 + 8);
        KF1.mvScaleFactors = KF2.mvScaleFactors;
        KF1.Ow = cv::Mat(3, 1, CV_32F); KF1.Tcw = cv::Mat::eye(4, 4, CV_32F); KF2.Tcw = cv::Mat::eye(4, 4, CV_32F);
        cv::Mat F12(3, 3, CV_32F);
        for (int i = 0; i < 3; i++) { KF1.Ow.at<float>(i) = tf[4 + i]; KF2.Tcw.at<float>(i, 3) = tf[7 + i]; }
        for (int i = 0; i < 9; i++) { KF2.Tcw.at<float>(i / 3, i % 3) = tf[10 + i]; F12.at<float>(i / 3, i % 3) = tf[19 + i]; }
        ORBmatcher matcherTri(0.6, true);                       // LocalMapping.cc:225: ORBmatcher matcher(0.6,false) uses checkOri false; true exercises the histogram
        nTri = matcherTri.SearchForTriangulation(&KF1, &KF2, F12, vMatchedIndices, onlyStereo != 0);
    }

    FILE* o = fopen(argv[2], "wb");
    fwrite(&nmatches, 4, 1, o);
    for (int i = 0; i < nF; i++) {
        int v = -1;
        if (F.mvpMapPoints[i]) v = F.mvpMapPoints[i] == &existing[i] ? -2 : (int)(F.mvpMapPoints[i] - &mps[0]);
        fwrite(&v, 4, 1, o);
    }
    fwrite(&nmBird, 4, 1, o);
    fwrite(vnMatches12.data(), 4, vnMatches12.size(), o);
    fwrite(&nmSBM, 4, 1, o);
    for (int i = 0; i < n2; i++) { const int v = Cur.mvpMapPointsBird[i] ? (int)Cur.mvpMapPointsBird[i]->mnId : -1; fwrite(&v, 4, 1, o); }
    const int dd = (n1 > 0 && n2 > 0) ? ORBmatcher::DescriptorDistance(Last.mDescriptorsBird, Cur.mDescriptorsBird) : -1;
    fwrite(&dd, 4, 1, o);
    fwrite(&nTri, 4, 1, o);
    for (size_t i = 0; i < vMatchedIndices.size(); i++) { const int p[2] = {(int)vMatchedIndices[i].first, (int)vMatchedIndices[i].second}; fwrite(p, 4, 2, o); }
    fclose(o);
    printf("%d projection matches, %d birdview matches, %d birdview landmarks carried over\n", nmatches, nmBird, nmSBM);
    return 0;
}
