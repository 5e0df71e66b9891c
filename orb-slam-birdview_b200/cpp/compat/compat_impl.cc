// Out-of-line members of the compat object model (compat/Frame.h, KeyFrame.h, MapPoint.h): the 64x48 lookup grids and their
// window queries (semantics of src/Frame.cc:378-412, 494-559, 879-944 and src/KeyFrame.cc:586-630), MapPoint::PredictScale
// (src/MapPoint.cc:385-417) and Frame::ProjectXYZ2Birdview (src/Frame.cc:1003-1010), written once over the stand-in members.
// Test scaffolding for builds without the reference tree's dependencies; inside the reference tree its own sources are used.
#include <algorithm>
#include <cmath>

#include "Frame.h"
#include "KeyFrame.h"
#include "MapPoint.h"

namespace ORB_SLAM2
{

std::vector<MapPoint::Event> MapPoint::mutationLog;

cv::Mat Frame::Tbc, Frame::Tcb;
int Frame::birdviewRows = 0, Frame::birdviewCols = 0;
const double Frame::pixel2meter = 0.03984;                     // src/Frame.cc:39-42 with correction = 1
const double Frame::meter2pixel = 25.1;
const double Frame::rear_axle_to_center = 1.393;
float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv, Frame::mfGridElementWidthInvBirdview, Frame::mfGridElementHeightInvBirdview;
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;

namespace
{
// level = clamp(ceil(log(maxDistance / dist) / logScaleFactor), 0, nLevels - 1)
int predicted_level(float maxDistance, float dist, float logScaleFactor, int nLevels)
{
    const float ratio = maxDistance / dist;
    const int level = std::ceil(std::log(ratio) / logScaleFactor);
    return level < 0 ? 0 : (level >= nLevels ? nLevels - 1 : level);
}

// grid cell of a keypoint, or false when it falls outside the grid (PosInGrid)
bool cell_of(const cv::KeyPoint& kp, float minX, float minY, float invW, float invH, int cols, int rows, int& cx, int& cy)
{
    cx = round((kp.pt.x - minX) * invW);
    cy = round((kp.pt.y - minY) * invH);
    return cx >= 0 && cx < cols && cy >= 0 && cy < rows;
}

// Window query shared by the three grids: cells [floor((c-r)*inv), ceil((c+r)*inv)] clamped to the grid, scanned column by
// column; optional octave range with the reference's convention (a bound < 0 is "no bound", the lower bound only counts when
// it is > 0 or an upper bound is present); strict |dx| < r and |dy| < r.
template <class CellAt>
std::vector<size_t> window(CellAt cellAt, const std::vector<cv::KeyPoint>& keys, int cols, int rows, float minX, float minY, float invW, float invH,
                           float x, float y, float r, int minLevel, int maxLevel)
{
    std::vector<size_t> found;
    found.reserve(keys.size());
    const int x0 = std::max(0, (int)std::floor((x - minX - r) * invW));
    if (x0 >= cols) return found;
    const int x1 = std::min(cols - 1, (int)std::ceil((x - minX + r) * invW));
    if (x1 < 0) return found;
    const int y0 = std::max(0, (int)std::floor((y - minY - r) * invH));
    if (y0 >= rows) return found;
    const int y1 = std::min(rows - 1, (int)std::ceil((y - minY + r) * invH));
    if (y1 < 0) return found;
    const bool levels = minLevel > 0 || maxLevel >= 0;
    for (int cx = x0; cx <= x1; cx++)
        for (int cy = y0; cy <= y1; cy++)
        {
            const std::vector<size_t>& cell = cellAt(cx, cy);
            for (size_t j = 0; j < cell.size(); j++)
            {
                const cv::KeyPoint& kp = keys[cell[j]];
                if (levels && (kp.octave < minLevel || (maxLevel >= 0 && kp.octave > maxLevel))) continue;
                if (std::fabs(kp.pt.x - x) < r && std::fabs(kp.pt.y - y) < r) found.push_back(cell[j]);
            }
        }
    return found;
}
}  // namespace

int MapPoint::PredictScale(const float &currentDist, KeyFrame* pKF) { return predicted_level(mfMaxDistance, currentDist, pKF->mfLogScaleFactor, pKF->mnScaleLevels); }
int MapPoint::PredictScale(const float &currentDist, Frame* pF) { return predicted_level(mfMaxDistance, currentDist, pF->mfLogScaleFactor, pF->mnScaleLevels); }

void Frame::AssignFeaturesToGrid()
{
    for (int i = 0; i < FRAME_GRID_COLS; i++)
        for (int j = 0; j < FRAME_GRID_ROWS; j++) { mGrid[i][j].clear(); mGridBirdview[i][j].clear(); }
    int cx, cy;
    for (int i = 0; i < N; i++)
        if (cell_of(mvKeysUn[i], mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv, FRAME_GRID_COLS, FRAME_GRID_ROWS, cx, cy)) mGrid[cx][cy].push_back(i);
    for (int i = 0; i < (int)mvKeysBird.size(); i++)
        if (cell_of(mvKeysBird[i], 0.f, 0.f, mfGridElementWidthInvBirdview, mfGridElementHeightInvBirdview, FRAME_GRID_COLS, FRAME_GRID_ROWS, cx, cy))
            mGridBirdview[cx][cy].push_back(i);
}

std::vector<size_t> Frame::GetFeaturesInArea(const float &x, const float &y, const float &r, const int minLevel, const int maxLevel) const
{
    return window([this](int cx, int cy) -> const std::vector<size_t>& { return mGrid[cx][cy]; }, mvKeysUn, FRAME_GRID_COLS, FRAME_GRID_ROWS,
                  mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv, x, y, r, minLevel, maxLevel);
}

std::vector<size_t> Frame::GetFeaturesInAreaBirdview(const float &x, const float &y, const float &r, const int minLevel, const int maxLevel) const
{
    return window([this](int cx, int cy) -> const std::vector<size_t>& { return mGridBirdview[cx][cy]; }, mvKeysBird, FRAME_GRID_COLS, FRAME_GRID_ROWS,
                  0.f, 0.f, mfGridElementWidthInvBirdview, mfGridElementHeightInvBirdview, x, y, r, minLevel, maxLevel);
}

cv::Point2f Frame::ProjectXYZ2Birdview(const cv::Point3f &p)
{
    // vehicle frame (x forward, y left) -> birdview pixel; the int halves and the double products are the reference's
    cv::Point2f pt;
    pt.x = birdviewCols/2-p.y*meter2pixel;
    pt.y = birdviewRows/2-(p.x-rear_axle_to_center)*meter2pixel;
    return pt;
}

void KeyFrame::AssignFeaturesToGrid()
{
    // a KeyFrame's grid is a copy of its Frame's (src/KeyFrame.cc:48-54): built with the Frame's float origin
    mGrid.assign(mnGridCols, std::vector< std::vector<size_t> >(mnGridRows));
    int cx, cy;
    for (int i = 0; i < N; i++)
        if (cell_of(mvKeysUn[i], Frame::mnMinX, Frame::mnMinY, mfGridElementWidthInv, mfGridElementHeightInv, mnGridCols, mnGridRows, cx, cy)) mGrid[cx][cy].push_back(i);
}

std::vector<size_t> KeyFrame::GetFeaturesInArea(const float &x, const float &y, const float &r) const
{
    // queried with the KeyFrame's own int origin and without a level filter (src/KeyFrame.cc:586-630)
    return window([this](int cx, int cy) -> const std::vector<size_t>& { return mGrid[cx][cy]; }, mvKeysUn, mnGridCols, mnGridRows,
                  (float)mnMinX, (float)mnMinY, mfGridElementWidthInv, mfGridElementHeightInv, x, y, r, -1, -1);
}

} // namespace ORB_SLAM2
