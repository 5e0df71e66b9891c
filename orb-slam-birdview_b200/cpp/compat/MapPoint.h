// Stand-in for the reference's include/MapPoint.h and include/MapPointBird.h: the members src/ORBmatcher.cc touches, with the
// reference's names and semantics (include/MapPoint.h:52-150, src/MapPoint.cc:373-417).  Used where the reference tree's own
// headers cannot be compiled (this image: no OpenCV / Eigen / DBoW2-with-OpenCV SDK): by the adapters in ../ORBmatcher_b200.cc
// and, for the oracle/_ref build, by the reference's UNMODIFIED src/ORBmatcher.cc.  Mutating calls (AddObservation, Replace,
// KeyFrame::AddMapPoint) are recorded so that a test can compare what two implementations of a matcher did to the map.
#ifndef MAPPOINT_H
#define MAPPOINT_H

#include <map>
#include <set>
#include <vector>

#include "../cv_compat.h"

namespace ORB_SLAM2
{

class KeyFrame;
class Frame;

class MapPoint
{
public:
    int Observations() { return nObs; }
    bool isBad() { return mbBad; }
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }
    cv::Mat GetWorldPos() { return mWorldPos.clone(); }
    cv::Mat GetNormal() { return mNormalVector.clone(); }
    float GetMinDistanceInvariance() { return 0.8f*mfMinDistance; }
    float GetMaxDistanceInvariance() { return 1.2f*mfMaxDistance; }
    int PredictScale(const float &currentDist, KeyFrame* pKF);
    int PredictScale(const float &currentDist, Frame* pF);
    bool IsInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) != 0; }
    int GetIndexInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF) ? (int)mObservations[pKF] : -1; }
    void AddObservation(KeyFrame* pKF, size_t idx)
    {
        if (mObservations.count(pKF)) return;                  // src/MapPoint.cc:89-91
        mObservations[pKF] = idx;
        nObs++;                                                // monocular observation (stereo adds 2: not modelled, only > comparisons matter)
        mutationLog.push_back(Event{'A', this, (MapPoint*)0, (long)idx});
    }
    void Replace(MapPoint* pMP)
    {
        if (pMP == this) return;                               // src/MapPoint.cc:187-188
        mbBad = true;
        mpReplaced = pMP;
        mutationLog.push_back(Event{'R', this, pMP, -1});
    }

    // Variables used by the tracking (include/MapPoint.h:92-97)
    float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0;
    bool mbTrackInView = false;
    int mnTrackScaleLevel = 0;
    float mTrackViewCos = 0;
    long unsigned int mnLastFrameSeen = 0;
    long unsigned int mnId = 0;

    // stand-in state behind the getters
    int nObs = 0;
    bool mbBad = false;
    cv::Mat mDescriptor, mWorldPos, mNormalVector;
    float mfMinDistance = 0, mfMaxDistance = 0;
    std::map<KeyFrame*, size_t> mObservations;
    MapPoint* mpReplaced = 0;

    struct Event { char what; MapPoint* a; MapPoint* b; long idx; };
    static std::vector<Event> mutationLog;                             // every mutation, in call order
};

class MapPointBird
{
public:
    cv::Mat GetWorldPos() { return mWorldPos.clone(); }
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }
    int Observations() { return nObs; }

    long unsigned int mnId = 0;
    long unsigned int mnLastFrameSeen = 0;

    int nObs = 0;
    cv::Mat mWorldPos, mDescriptor;
};

} // namespace ORB_SLAM2

#endif // MAPPOINT_H
