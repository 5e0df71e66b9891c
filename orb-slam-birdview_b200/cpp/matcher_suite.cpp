// Test driver: every public ORB_SLAM2::ORBmatcher method, called through Frame / KeyFrame / MapPoint / MapPointBird objects the
// way Tracking, LocalMapping and LoopClosing call them, on a scene read from a flat binary file; what each call returned and
// left in the objects is written to an output file.  The same source is linked twice:
//   matcher_suite        + ORBmatcher_b200.cc (the adapters over liborbb200.so)              -> the product path (needs a GPU)
//   matcher_suite_ref    + the reference's own, UNMODIFIED src/ORBmatcher.cc (oracle/_ref)   -> the expected output (CPU)
// Both against the compat object model (compat/*.h); tests/test_matcher_suite.py writes the scene and compares the outputs.
//
//   matcher_suite <scene.bin> <out.bin>
// Scene layout: see read_scene() below and tests/matcher_scene.py (the writer).  Every call runs on a fresh copy of the scene.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <vector>

#include "ORBmatcher.h"

using namespace ORB_SLAM2;
using std::vector;

static std::vector<unsigned char> buf;
static size_t off = 0;
template <class T> static const T* take(size_t n)
{
    const T* p = reinterpret_cast<const T*>(buf.data() + off);
    off += n * sizeof(T);
    if (off > buf.size()) { fprintf(stderr, "scene file too short\n"); exit(2); }
    return p;
}
static int geti() { return *take<int>(1); }
static float getf() { return *take<float>(1); }

static cv::Mat rows32(const unsigned char* d, int n)
{
    cv::Mat m(n > 0 ? n : 1, 32, CV_8U);
    if (n > 0) memcpy(m.data, d, (size_t)n * 32);
    return m;
}
static cv::Mat matf(const float* v, int r, int c)
{
    cv::Mat m(r, c, CV_32F);
    for (int i = 0; i < r; i++) for (int j = 0; j < c; j++) m.at<float>(i, j) = v[i * c + j];
    return m;
}
static DBoW2::FeatureVector featvec()
{
    DBoW2::FeatureVector fv;
    const int nn = geti();
    const int* node = take<int>(nn);
    const int* ptr = take<int>(nn + 1);
    const int* idx = take<int>(ptr[nn]);
    for (int k = 0; k < nn; k++)
        for (int j = ptr[k]; j < ptr[k + 1]; j++) fv[(unsigned)node[k]].push_back((unsigned)idx[j]);
    return fv;
}

struct Scene {
    vector<MapPoint> mps;
    vector<MapPointBird> birds;
    vector<Frame> frames;          // [0] = last, [1] = current
    vector<KeyFrame> kfs;          // [0], [1]
    // call parameters
    float th_local, nnratio_local, th_frame, th_kf, th_scw, th_fuse, th_sim3, r_bird_proj, r_bird_kf, nnratio_bow, nnratio_init, nnratio_bird, nnratio_tri, s12;
    int orb_dist, win_init, win_bird, only_stereo;
    cv::Mat Scw, R12, t12, F12;
    vector<int> already_found, scw_points, scw_matched, fuse_points, fuse_scw_points, sim3_matches12, bird_proj_points;
    vector<cv::Point2f> init_prev, bird_prev;
};

static void read_views(Scene& S, bool keyframes, int count, float scaleFactor, int nLevels)
{
    for (int v = 0; v < count; v++) {
        const int id = geti(), N = geti(), Nb = geti();
        const float* intr = take<float>(6);                    // fx fy cx cy bf b
        const float* Tcw = take<float>(16);
        const cv::KeyPoint* kps = take<cv::KeyPoint>(N);
        const unsigned char* desc = take<unsigned char>((size_t)N * 32);
        const float* ur = take<float>(N);
        const int* mp = take<int>(N);
        const unsigned char* outl = take<unsigned char>(N);
        DBoW2::FeatureVector fv = featvec();
        const cv::KeyPoint* bk = take<cv::KeyPoint>(Nb);
        const unsigned char* bdesc = take<unsigned char>((size_t)Nb * 32);
        const int* bmp = take<int>(Nb);
        vector<float> sf(nLevels), isf(nLevels), s2(nLevels), is2(nLevels);
        sf[0] = 1.f; s2[0] = 1.f;
        for (int i = 1; i < nLevels; i++) { sf[i] = sf[i - 1] * scaleFactor; s2[i] = sf[i] * sf[i]; }
        for (int i = 0; i < nLevels; i++) { isf[i] = 1.0f / sf[i]; is2[i] = 1.0f / s2[i]; }
        cv::Mat T = matf(Tcw, 4, 4);
        if (!keyframes) {
            S.frames.push_back(Frame());
            Frame& F = S.frames.back();
            F.mnId = id; F.N = N;
            F.fx = intr[0]; F.fy = intr[1]; F.cx = intr[2]; F.cy = intr[3]; F.mbf = intr[4]; F.mb = intr[5];
            F.mvKeys.assign(kps, kps + N); F.mvKeysUn = F.mvKeys;
            F.mDescriptors = rows32(desc, N);
            F.mvuRight.assign(ur, ur + N);
            F.mvDepth.assign(N, -1.f);
            F.mvpMapPoints.assign(N, static_cast<MapPoint*>(NULL));
            F.mvbOutlier.assign(N, false);
            for (int i = 0; i < N; i++) { if (mp[i] >= 0) F.mvpMapPoints[i] = &S.mps[mp[i]]; F.mvbOutlier[i] = outl[i] != 0; }
            F.mFeatVec = fv;
            F.mTcw = T;
            F.mnScaleLevels = nLevels; F.mfScaleFactor = scaleFactor; F.mfLogScaleFactor = log(scaleFactor);
            F.mvScaleFactors = sf; F.mvInvScaleFactors = isf; F.mvLevelSigma2 = s2; F.mvInvLevelSigma2 = is2;
            F.mvKeysBird.assign(bk, bk + Nb);
            F.mDescriptorsBird = rows32(bdesc, Nb);
            F.mvpMapPointsBird.assign(Nb, static_cast<MapPointBird*>(NULL));
            for (int i = 0; i < Nb; i++) if (bmp[i] >= 0) F.mvpMapPointsBird[i] = &S.birds[bmp[i]];
            F.AssignFeaturesToGrid();
        } else {
            S.kfs.push_back(KeyFrame());
            KeyFrame& K = S.kfs.back();
            K.mnId = id; K.N = N;
            K.fx = intr[0]; K.fy = intr[1]; K.cx = intr[2]; K.cy = intr[3]; K.mbf = intr[4]; K.mb = intr[5];
            K.invfx = 1.0f / K.fx; K.invfy = 1.0f / K.fy;
            K.mvKeys.assign(kps, kps + N); K.mvKeysUn = K.mvKeys;
            K.mDescriptors = rows32(desc, N);
            K.mvuRight.assign(ur, ur + N);
            K.mvpMapPoints.assign(N, static_cast<MapPoint*>(NULL));
            for (int i = 0; i < N; i++) if (mp[i] >= 0) K.mvpMapPoints[i] = &S.mps[mp[i]];
            K.mFeatVec = fv;
            K.Tcw = T;
            cv::Mat Rwc = T.rowRange(0, 3).colRange(0, 3).t();
            K.Ow = -Rwc * T.rowRange(0, 3).col(3);             // src/KeyFrame.cc:71-73
            K.mnScaleLevels = nLevels; K.mfScaleFactor = scaleFactor; K.mfLogScaleFactor = log(scaleFactor);
            K.mvScaleFactors = sf; K.mvLevelSigma2 = s2; K.mvInvLevelSigma2 = is2;
            K.mnMinX = (int)Frame::mnMinX; K.mnMinY = (int)Frame::mnMinY; K.mnMaxX = (int)Frame::mnMaxX; K.mnMaxY = (int)Frame::mnMaxY;
            K.mfGridElementWidthInv = Frame::mfGridElementWidthInv; K.mfGridElementHeightInv = Frame::mfGridElementHeightInv;
            K.mvKeysBird.assign(bk, bk + Nb);
            K.mvpMapPointsBird.assign(Nb, static_cast<MapPointBird*>(NULL));
            for (int i = 0; i < Nb; i++) if (bmp[i] >= 0) K.mvpMapPointsBird[i] = &S.birds[bmp[i]];
            K.AssignFeaturesToGrid();
        }
    }
}

static vector<int> geti_list() { const int n = geti(); const int* p = take<int>(n); return vector<int>(p, p + n); }

static void read_scene(Scene& S)
{
    off = 0;
    MapPoint::mutationLog.clear();
    const int nLevels = geti();
    const float scaleFactor = getf();
    const float* g = take<float>(8);          // mnMinX mnMaxX mnMinY mnMaxY invW invH invWbird invHbird
    Frame::mnMinX = g[0]; Frame::mnMaxX = g[1]; Frame::mnMinY = g[2]; Frame::mnMaxY = g[3];
    Frame::mfGridElementWidthInv = g[4]; Frame::mfGridElementHeightInv = g[5];
    Frame::mfGridElementWidthInvBirdview = g[6]; Frame::mfGridElementHeightInvBirdview = g[7];
    Frame::birdviewCols = geti(); Frame::birdviewRows = geti();
    Frame::Tbc = matf(take<float>(16), 4, 4);
    const int NM = geti();
    S.mps.assign(NM, MapPoint());
    for (int i = 0; i < NM; i++) {
        MapPoint& p = S.mps[i];
        p.mnId = i;
        p.mWorldPos = matf(take<float>(3), 3, 1);
        p.mNormalVector = matf(take<float>(3), 3, 1);
        p.mfMinDistance = getf(); p.mfMaxDistance = getf();
        p.mDescriptor = rows32(take<unsigned char>(32), 1);
        p.nObs = geti(); p.mbBad = geti() != 0;
        p.mbTrackInView = geti() != 0;
        p.mTrackProjX = getf(); p.mTrackProjY = getf(); p.mTrackProjXR = getf(); p.mTrackViewCos = getf();
        p.mnTrackScaleLevel = geti();
        p.mnLastFrameSeen = (unsigned long)geti();
    }
    const int NB = geti();
    S.birds.assign(NB, MapPointBird());
    for (int i = 0; i < NB; i++) {
        MapPointBird& b = S.birds[i];
        b.mnId = i;
        b.mWorldPos = matf(take<float>(3), 3, 1);
        b.mDescriptor = rows32(take<unsigned char>(32), 1);
        b.nObs = geti();
        b.mnLastFrameSeen = (unsigned long)geti();
    }
    S.frames.clear(); S.kfs.clear();
    S.frames.reserve(2); S.kfs.reserve(2);
    read_views(S, false, 2, scaleFactor, nLevels);
    read_views(S, true, 2, scaleFactor, nLevels);
    // observations of the map points in the two keyframes (MapPoint::mObservations), from the keyframes' own associations
    for (int k = 0; k < 2; k++)
        for (int i = 0; i < S.kfs[k].N; i++)
            if (S.kfs[k].mvpMapPoints[i]) S.kfs[k].mvpMapPoints[i]->mObservations[&S.kfs[k]] = i;
    MapPoint::mutationLog.clear();
    const float* p = take<float>(14);
    S.th_local = p[0]; S.nnratio_local = p[1]; S.th_frame = p[2]; S.th_kf = p[3]; S.th_scw = p[4]; S.th_fuse = p[5]; S.th_sim3 = p[6];
    S.r_bird_proj = p[7]; S.r_bird_kf = p[8]; S.nnratio_bow = p[9]; S.nnratio_init = p[10]; S.nnratio_bird = p[11]; S.nnratio_tri = p[12]; S.s12 = p[13];
    S.orb_dist = geti(); S.win_init = geti(); S.win_bird = geti(); S.only_stereo = geti();
    S.Scw = matf(take<float>(16), 4, 4);
    S.R12 = matf(take<float>(9), 3, 3);
    S.t12 = matf(take<float>(3), 3, 1);
    S.F12 = matf(take<float>(9), 3, 3);
    S.already_found = geti_list(); S.scw_points = geti_list(); S.scw_matched = geti_list(); S.fuse_points = geti_list();
    S.fuse_scw_points = geti_list(); S.sim3_matches12 = geti_list(); S.bird_proj_points = geti_list();
    { const int n = geti(); const float* q = take<float>(2 * n); S.init_prev.clear(); for (int i = 0; i < n; i++) S.init_prev.push_back(cv::Point2f(q[2 * i], q[2 * i + 1])); }
    { const int n = geti(); const float* q = take<float>(2 * n); S.bird_prev.clear(); for (int i = 0; i < n; i++) S.bird_prev.push_back(cv::Point2f(q[2 * i], q[2 * i + 1])); }
}

static FILE* out;
static void put(int v) { fwrite(&v, 4, 1, out); }
static void put_tag(const char* t) { char b[8] = {0}; strncpy(b, t, 8); fwrite(b, 1, 8, out); }
static int idx_of(const Scene& S, const MapPoint* p) { return p ? (int)(p - &S.mps[0]) : -1; }
static int idx_of(const Scene& S, const MapPointBird* p) { return p ? (int)(p - &S.birds[0]) : -1; }
static void put_mps(const Scene& S, const vector<MapPoint*>& v) { put((int)v.size()); for (size_t i = 0; i < v.size(); i++) put(idx_of(S, v[i])); }
static void put_birds(const Scene& S, const vector<MapPointBird*>& v) { put((int)v.size()); for (size_t i = 0; i < v.size(); i++) put(idx_of(S, v[i])); }
static void put_ints(const vector<int>& v) { put((int)v.size()); for (size_t i = 0; i < v.size(); i++) put(v[i]); }
static void put_log(const Scene& S)
{
    put((int)MapPoint::mutationLog.size());
    for (size_t i = 0; i < MapPoint::mutationLog.size(); i++) {
        const MapPoint::Event& e = MapPoint::mutationLog[i];
        put((int)e.what); put(idx_of(S, e.a)); put(idx_of(S, e.b)); put((int)e.idx);
    }
}
static vector<MapPoint*> pick(Scene& S, const vector<int>& ids)
{
    vector<MapPoint*> v(ids.size());
    for (size_t i = 0; i < ids.size(); i++) v[i] = ids[i] >= 0 ? &S.mps[ids[i]] : static_cast<MapPoint*>(NULL);
    return v;
}

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: matcher_suite scene.bin out.bin\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    fseek(f, 0, SEEK_END); buf.resize(ftell(f)); fseek(f, 0, SEEK_SET);
    if (fread(buf.data(), 1, buf.size(), f) != buf.size()) return 2;
    fclose(f);
    out = fopen(argv[2], "wb");
    Scene S;

    // 1. Tracking::SearchLocalPoints: SearchByProjection(F, vpLocalMapPoints, th)                         src/Tracking.cc:1659
    read_scene(S);
    {
        vector<MapPoint*> all(S.mps.size());
        for (size_t i = 0; i < all.size(); i++) all[i] = &S.mps[i];
        ORBmatcher m(S.nnratio_local);
        put_tag("SBP_MPS"); put(m.SearchByProjection(S.frames[1], all, S.th_local)); put_mps(S, S.frames[1].mvpMapPoints);
    }
    // 2. Tracking::TrackWithMotionModel: SearchByProjection(Cur, Last, th, bMono), stereo and mono         src/Tracking.cc:1227
    for (int mono = 0; mono < 2; mono++) {
        read_scene(S);
        ORBmatcher m(0.9, true);
        // the tracker clears the current frame's associations first (src/Tracking.cc:1219)
        fill(S.frames[1].mvpMapPoints.begin(), S.frames[1].mvpMapPoints.end(), static_cast<MapPoint*>(NULL));
        put_tag(mono ? "SBP_FFM" : "SBP_FFS"); put(m.SearchByProjection(S.frames[1], S.frames[0], S.th_frame, mono != 0)); put_mps(S, S.frames[1].mvpMapPoints);
    }
    // 3. Tracking::Relocalization: SearchByProjection(Cur, pKF, sFound, th, ORBdist)                        src/Tracking.cc:2013
    read_scene(S);
    {
        std::set<MapPoint*> found;
        for (size_t i = 0; i < S.already_found.size(); i++) found.insert(&S.mps[S.already_found[i]]);
        ORBmatcher m(0.9, true);
        put_tag("SBP_FKF"); put(m.SearchByProjection(S.frames[1], &S.kfs[0], found, S.th_kf, S.orb_dist)); put_mps(S, S.frames[1].mvpMapPoints);
    }
    // 4. LoopClosing::ComputeSim3: SearchByProjection(pKF, Scw, vpLoopMapPoints, vpMatched, th)             src/LoopClosing.cc:375
    read_scene(S);
    {
        vector<MapPoint*> pts = pick(S, S.scw_points), matched = pick(S, S.scw_matched);
        ORBmatcher m(0.75, true);
        put_tag("SBP_SCW"); put(m.SearchByProjection(&S.kfs[0], S.Scw, pts, matched, (int)S.th_scw)); put_mps(S, matched);
    }
    // 5./6. SearchByBoW(KF, F) (Tracking::TrackReferenceKeyFrame, src/Tracking.cc:1032) and (KF, KF) (LoopClosing.cc:265)
    read_scene(S);
    {
        vector<MapPoint*> vp;
        ORBmatcher m(S.nnratio_bow, true);
        put_tag("BOW_KFF"); put(m.SearchByBoW(&S.kfs[0], S.frames[1], vp)); put_mps(S, vp);
    }
    read_scene(S);
    {
        vector<MapPoint*> vp;
        ORBmatcher m(S.nnratio_bow, true);
        put_tag("BOW_KK"); put(m.SearchByBoW(&S.kfs[0], &S.kfs[1], vp)); put_mps(S, vp);
    }
    // 7. Tracking::MonocularInitialization: SearchForInitialization(Init, Cur, prev, matches, 100)          src/Tracking.cc:739
    read_scene(S);
    {
        vector<cv::Point2f> prev = S.init_prev;
        vector<int> m12;
        ORBmatcher m(S.nnratio_init, true);
        put_tag("INIT"); put(m.SearchForInitialization(S.frames[0], S.frames[1], prev, m12, S.win_init)); put_ints(m12);
        put((int)prev.size());
        for (size_t i = 0; i < prev.size(); i++) { fwrite(&prev[i].x, 4, 1, out); fwrite(&prev[i].y, 4, 1, out); }
    }
    // 8. LocalMapping::CreateNewMapPoints: SearchForTriangulation(KF1, KF2, F12, pairs, bOnlyStereo)        src/LocalMapping.cc:278
    read_scene(S);
    {
        vector<std::pair<size_t, size_t> > pairs;
        ORBmatcher m(S.nnratio_tri, true);
        put_tag("TRIANG"); put(m.SearchForTriangulation(&S.kfs[0], &S.kfs[1], S.F12, pairs, S.only_stereo != 0));
        put((int)pairs.size());
        for (size_t i = 0; i < pairs.size(); i++) { put((int)pairs[i].first); put((int)pairs[i].second); }
    }
    // 9. LoopClosing::ComputeSim3: SearchBySim3(KF1, KF2, vpMatches12, s, R, t, 7.5)                        src/LoopClosing.cc:323
    read_scene(S);
    {
        vector<MapPoint*> m12 = pick(S, S.sim3_matches12);
        ORBmatcher m(0.75, true);
        put_tag("SIM3"); put(m.SearchBySim3(&S.kfs[0], &S.kfs[1], m12, S.s12, S.R12, S.t12, S.th_sim3)); put_mps(S, m12);
    }
    // 10. LocalMapping::SearchInNeighbors: Fuse(pKF, vpMapPoints, th)                                       src/LocalMapping.cc:499
    read_scene(S);
    {
        vector<MapPoint*> pts = pick(S, S.fuse_points);
        ORBmatcher m;
        put_tag("FUSE"); put(m.Fuse(&S.kfs[0], pts, S.th_fuse)); put_mps(S, S.kfs[0].mvpMapPoints); put_log(S);
    }
    // 11. LoopClosing::SearchAndFuse: Fuse(pKF, Scw, vpLoopMapPoints, 4, vpReplacePoints)                   src/LoopClosing.cc:599
    read_scene(S);
    {
        vector<MapPoint*> pts = pick(S, S.fuse_scw_points), repl(pts.size(), static_cast<MapPoint*>(NULL));
        ORBmatcher m(0.8);
        put_tag("FUSE_SCW"); put(m.Fuse(&S.kfs[0], S.Scw, pts, S.th_fuse, repl)); put_mps(S, repl); put_mps(S, S.kfs[0].mvpMapPoints); put_log(S);
    }
    // 12./13. BirdviewMatch with and without previous matches                                              src/Tracking.cc:744, 2158
    read_scene(S);
    {
        vector<cv::Point2f> prev = S.bird_prev;
        vector<int> m12;
        ORBmatcher m(S.nnratio_bird, true);
        put_tag("BIRD_PRV"); put(m.BirdviewMatch(S.frames[0], S.frames[1], m12, prev, S.win_bird)); put_ints(m12);
        put((int)prev.size());
        for (size_t i = 0; i < prev.size(); i++) { fwrite(&prev[i].x, 4, 1, out); fwrite(&prev[i].y, 4, 1, out); }
    }
    read_scene(S);
    {
        vector<int> m12;
        ORBmatcher m(S.nnratio_bird, true);
        put_tag("BIRD"); put(m.BirdviewMatch((const Frame&)S.frames[0], (const Frame&)S.frames[1], m12, S.win_bird)); put_ints(m12);
    }
    // 14. Tracking::SearchLocalPoints: SearchByProjectionBird(F, mvpLocalMapPointsBird)                     src/Tracking.cc:1672
    read_scene(S);
    {
        vector<MapPointBird*> pts(S.bird_proj_points.size());
        for (size_t i = 0; i < pts.size(); i++) pts[i] = &S.birds[S.bird_proj_points[i]];
        ORBmatcher m(S.nnratio_bird, true);
        put_tag("BIRD_PRJ"); put(m.SearchByProjectionBird(S.frames[1], pts, S.r_bird_proj)); put_birds(S, S.frames[1].mvpMapPointsBird);
    }
    // 15. Tracking::TrackWithMotionModel: SearchByMatchBird(Cur, Last, 15)                                  src/Tracking.cc:1241
    read_scene(S);
    {
        fill(S.frames[1].mvpMapPointsBird.begin(), S.frames[1].mvpMapPointsBird.end(), static_cast<MapPointBird*>(NULL));
        ORBmatcher m(S.nnratio_bird, true);
        put_tag("BIRD_FF"); put(m.SearchByMatchBird(S.frames[1], S.frames[0], S.win_bird)); put_birds(S, S.frames[1].mvpMapPointsBird);
    }
    // 16. Tracking::TrackReferenceKeyFrame: SearchByMatchBird(pKF, F, vpMatches, 15)                        src/Tracking.cc:1046
    read_scene(S);
    {
        vector<MapPointBird*> vp;
        ORBmatcher m(S.nnratio_bird, true);
        put_tag("BIRD_KF"); put(m.SearchByMatchBird(&S.kfs[0], S.frames[1], vp, S.r_bird_kf)); put_birds(S, vp);
    }
    // static ORBmatcher::DescriptorDistance (src/Frame.cc:737, src/MapPoint.cc:281)
    put_tag("DIST"); put(ORBmatcher::DescriptorDistance(S.frames[0].mDescriptors.row(0), S.frames[1].mDescriptors.row(0)));
    fclose(out);
    return 0;
}
