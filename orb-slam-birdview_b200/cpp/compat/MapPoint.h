// Stand-in for the reference's include/MapPoint.h and include/MapPointBird.h: ONLY the members the matcher adapters in
// ../ORBmatcher_b200.cc touch, with the reference's names (include/MapPoint.h:52-97).  Used where the reference tree is
// absent (this image: tests); inside the reference tree its own headers are found first.
#ifndef MAPPOINT_H
#define MAPPOINT_H

#include "../cv_compat.h"

namespace ORB_SLAM2
{

class MapPoint
{
public:
    int Observations() { return nObs; }
    bool isBad() { return mbBad; }
    cv::Mat GetDescriptor() { return mDescriptor; }

    // Variables used by the tracking (include/MapPoint.h:92-97)
    float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0;
    bool mbTrackInView = false;
    int mnTrackScaleLevel = 0;
    float mTrackViewCos = 0;

    // stand-in state behind the getters
    int nObs = 0;
    bool mbBad = false;
    cv::Mat mDescriptor;
};

class MapPointBird
{
public:
    long unsigned int mnId = 0;
};

} // namespace ORB_SLAM2

#endif // MAPPOINT_H
