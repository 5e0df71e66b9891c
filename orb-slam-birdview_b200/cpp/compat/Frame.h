// Stand-in for the reference's include/Frame.h: ONLY the members the matcher adapters in ../ORBmatcher_b200.cc touch,
// with the reference's names (include/Frame.h:153-239).  See compat/MapPoint.h.
#ifndef FRAME_H
#define FRAME_H

#include <vector>

#include "MapPoint.h"

namespace ORB_SLAM2
{
#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

class Frame
{
public:
    int N = 0;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvuRight;
    cv::Mat mDescriptors;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<float> mvScaleFactors;

    // birdview (include/Frame.h:164-177)
    std::vector<cv::KeyPoint> mvKeysBird;
    cv::Mat mDescriptorsBird;
    std::vector<MapPointBird*> mvpMapPointsBird;

    long unsigned int mnId = 0;

    static float mfGridElementWidthInv, mfGridElementHeightInv;
    static float mfGridElementWidthInvBirdview, mfGridElementHeightInvBirdview;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
};

} // namespace ORB_SLAM2

#endif // FRAME_H
