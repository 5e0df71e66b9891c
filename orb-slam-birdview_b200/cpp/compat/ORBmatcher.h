// Stand-in for the reference's include/ORBmatcher.h: the class declaration (include/ORBmatcher.h:38-120) reduced to the
// methods ../ORBmatcher_b200.cc defines.  Inside the reference tree its own header declares them (and the rest).
#ifndef ORBMATCHER_H
#define ORBMATCHER_H

#include <vector>

#include "Frame.h"
#include "KeyFrame.h"
#include "MapPoint.h"

namespace ORB_SLAM2
{

class ORBmatcher
{
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true);

    static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b);

    int SearchByProjection(Frame &F, const std::vector<MapPoint*> &vpMapPoints, const float th = 3);

    int SearchForTriangulation(KeyFrame *pKF1, KeyFrame* pKF2, cv::Mat F12,
                               std::vector<std::pair<size_t, size_t> > &vMatchedPairs, const bool bOnlyStereo);

    int BirdviewMatch(const Frame &F1, const Frame &F2, std::vector<int> &vnMatches12, int windowSize = 10);

    int SearchByMatchBird(Frame &CurrentFrame, const Frame &LastFrame, const int windowSize = 10);

public:
    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;

protected:
    float mfNNratio;
    bool mbCheckOrientation;
};

} // namespace ORB_SLAM2

#endif // ORBMATCHER_H
