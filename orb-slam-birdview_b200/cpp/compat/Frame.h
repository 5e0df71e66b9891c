// Stand-in for the reference's include/Frame.h: the members src/ORBmatcher.cc touches, with the reference's names and
// semantics (include/Frame.h:96-239, src/Frame.cc:378-412, 494-559, 879-944, 1003-1010).  See compat/MapPoint.h.
#ifndef FRAME_H
#define FRAME_H

#include <vector>

#include "MapPoint.h"
#include "KeyFrame.h"

namespace ORB_SLAM2
{
#ifndef FRAME_GRID_ROWS
#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64
#endif

class Frame
{
public:
    std::vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const int minLevel=-1, const int maxLevel=-1) const;
    std::vector<size_t> GetFeaturesInAreaBirdview(const float &x, const float &y, const float &r, const int minLevel=-1, const int maxLevel=-1) const;
    void AssignFeaturesToGrid();                               // src/Frame.cc:378-412 (front and birdview grids)
    static cv::Point2f ProjectXYZ2Birdview(const cv::Point3f &p);

    long unsigned int mnId = 0;
    int N = 0;
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0, mb = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors;
    DBoW2::FeatureVector mFeatVec;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    cv::Mat mTcw;
    int mnScaleLevels = 8;
    float mfScaleFactor = 1.2f, mfLogScaleFactor = 0;
    std::vector<float> mvScaleFactors, mvInvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;

    // birdview (include/Frame.h:164-183)
    std::vector<cv::KeyPoint> mvKeysBird;
    cv::Mat mDescriptorsBird;
    std::vector<MapPointBird*> mvpMapPointsBird;
    std::vector<int> mvnBirdviewMatches;
    static cv::Mat Tbc, Tcb;
    static int birdviewRows, birdviewCols;
    static const double pixel2meter, meter2pixel, rear_axle_to_center;

    static float mfGridElementWidthInv, mfGridElementHeightInv;
    static float mfGridElementWidthInvBirdview, mfGridElementHeightInvBirdview;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;

    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    std::vector<std::size_t> mGridBirdview[FRAME_GRID_COLS][FRAME_GRID_ROWS];
};

} // namespace ORB_SLAM2

#endif // FRAME_H
