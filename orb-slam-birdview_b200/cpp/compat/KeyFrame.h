// Stand-in for the reference's include/KeyFrame.h (and DBoW2::FeatureVector): ONLY the members
// ORBmatcher::SearchForTriangulation touches (src/ORBmatcher.cc:657-823), with the reference's names.  See compat/MapPoint.h.
#ifndef KEYFRAME_H
#define KEYFRAME_H

#include <map>
#include <vector>

#include "MapPoint.h"

namespace DBoW2
{
// Thirdparty/DBoW2/DBoW2/FeatureVector.h: node id -> indices of the local features under that node
class FeatureVector : public std::map<unsigned int, std::vector<unsigned int> > {};
}

namespace ORB_SLAM2
{

class KeyFrame
{
public:
    cv::Mat GetRotation() { return Rcw; }
    cv::Mat GetTranslation() { return tcw; }
    cv::Mat GetCameraCenter() { return Ow; }
    MapPoint* GetMapPoint(const size_t &idx) { return mvpMapPoints[idx]; }

    float fx = 0, fy = 0, cx = 0, cy = 0;
    int N = 0;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvuRight;
    cv::Mat mDescriptors;
    DBoW2::FeatureVector mFeatVec;
    std::vector<float> mvScaleFactors;
    std::vector<float> mvLevelSigma2;

    // stand-in state behind the getters
    cv::Mat Rcw, tcw, Ow;
    std::vector<MapPoint*> mvpMapPoints;
};

} // namespace ORB_SLAM2

#endif // KEYFRAME_H
