// Stand-in for the reference's include/KeyFrame.h: the members src/ORBmatcher.cc touches, with the reference's names and
// semantics (include/KeyFrame.h, src/KeyFrame.cc:586-630).  See compat/MapPoint.h.
#ifndef KEYFRAME_H
#define KEYFRAME_H

#include <map>
#include <set>
#include <vector>

#include "MapPoint.h"
#include "FeatureVector.h"

namespace ORB_SLAM2
{

#ifndef FRAME_GRID_ROWS
#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64
#endif

class KeyFrame
{
public:
    cv::Mat GetPose() { return Tcw.clone(); }
    cv::Mat GetRotation() { return Tcw.rowRange(0,3).colRange(0,3).clone(); }
    cv::Mat GetTranslation() { return Tcw.rowRange(0,3).col(3).clone(); }
    cv::Mat GetCameraCenter() { return Ow.clone(); }
    MapPoint* GetMapPoint(const size_t &idx) { return mvpMapPoints[idx]; }
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    std::set<MapPoint*> GetMapPoints()
    {
        std::set<MapPoint*> s;                                                   // src/KeyFrame.cc:248-261
        for (size_t i = 0; i < mvpMapPoints.size(); i++) if (mvpMapPoints[i] && !mvpMapPoints[i]->isBad()) s.insert(mvpMapPoints[i]);
        return s;
    }
    std::vector<MapPointBird*> GetMapPointMatchesBird() { return mvpMapPointsBird; }
    void AddMapPoint(MapPoint* pMP, const size_t &idx) { mvpMapPoints[idx] = pMP; MapPoint::mutationLog.push_back(MapPoint::Event{'K', pMP, (MapPoint*)0, (long)idx}); }
    std::vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r) const;
    bool IsInImage(const float &x, const float &y) const { return (x>=mnMinX && x<mnMaxX && y>=mnMinY && y<mnMaxY); }
    void AssignFeaturesToGrid();                                                 // what KeyFrame copies from Frame::mGrid (src/KeyFrame.cc:48-54)

    long unsigned int mnId = 0;
    float fx = 0, fy = 0, cx = 0, cy = 0, invfx = 0, invfy = 0, mbf = 0, mb = 0;
    int N = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight;
    cv::Mat mDescriptors;
    DBoW2::FeatureVector mFeatVec;
    int mnScaleLevels = 8;
    float mfScaleFactor = 1.2f, mfLogScaleFactor = 0;
    std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;                          // const int in the reference (include/KeyFrame.h:186-189)
    int mnGridCols = FRAME_GRID_COLS, mnGridRows = FRAME_GRID_ROWS;
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;

    // birdview (include/KeyFrame.h)
    std::vector<cv::KeyPoint> mvKeysBird;

    // stand-in state behind the getters
    cv::Mat Tcw, Ow;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<MapPointBird*> mvpMapPointsBird;
    std::vector< std::vector< std::vector<size_t> > > mGrid;
};

} // namespace ORB_SLAM2

#endif // KEYFRAME_H
