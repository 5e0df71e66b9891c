/* ORACLE -- TEST INFRASTRUCTURE ONLY (see orb_oracle.h).  Matcher half.
 *
 * Restates /root/reference/src/ORBmatcher.cc (Hamming scans) and the Frame lookup grid
 * (/root/reference/src/Frame.cc:378-412,494-559,879-944) on plain arrays.  The object graph
 * (MapPoint*, KeyFrame*, cv::Mat) the reference walks is flattened by the caller; every loop below keeps
 * the reference's iteration order, comparison operators and loop-carried state.
 */
#include "orb_oracle.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

namespace {

const int TH_HIGH = 100;      /* src/ORBmatcher.cc:37 */
const int TH_LOW = 50;        /* :38 */
const int HISTO_LENGTH = 30;  /* :39 */
const int FRAME_GRID_ROWS = 48;  /* include/Frame.h:39 */
const int FRAME_GRID_COLS = 64;  /* include/Frame.h:40 */

/* ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:1647-1663 */
inline int DescriptorDistance(const uint8_t* a, const uint8_t* b)
{
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t wa, wb;
        memcpy(&wa, a + 4 * i, 4);
        memcpy(&wb, b + 4 * i, 4);
        unsigned int v = wa ^ wb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

/* ORBmatcher::ComputeThreeMaxima, src/ORBmatcher.cc:1601-1642 */
void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3)
{
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

inline int rot_bin(float a1, float a2)
{
    const float factor = 1.0f / HISTO_LENGTH;
    float rot = a1 - a2;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)round(rot * factor);
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

/* ORBmatcher::RadiusByViewingCos, src/ORBmatcher.cc:131-137 */
inline float RadiusByViewingCos(float viewCos) { return viewCos > 0.998 ? 2.5f : 4.0f; }

}  // namespace

struct oracle_frame {
    int N;
    std::vector<oracle_kp_t> keys;
    std::vector<uint8_t> desc;
    std::vector<float> uRight;
    float minX, minY, invW, invH;
    std::vector<int> grid[FRAME_GRID_COLS][FRAME_GRID_ROWS];

    /* Frame::PosInGrid / PosInGridBirdview (src/Frame.cc:549-559, 879-889); birdview = minX=minY=0 */
    bool PosInGrid(const oracle_kp_t& kp, int& posX, int& posY) const
    {
        posX = (int)round((kp.x - minX) * invW);
        posY = (int)round((kp.y - minY) * invH);
        if (posX < 0 || posX >= FRAME_GRID_COLS || posY < 0 || posY >= FRAME_GRID_ROWS) return false;
        return true;
    }

    /* Frame::GetFeaturesInArea / GetFeaturesInAreaBirdview (src/Frame.cc:494-547, 891-944) */
    void GetFeaturesInArea(float x, float y, float r, int minLevel, int maxLevel, std::vector<int>& vIndices) const
    {
        vIndices.clear();
        const int nMinCellX = std::max(0, (int)floor((x - minX - r) * invW));
        if (nMinCellX >= FRAME_GRID_COLS) return;
        const int nMaxCellX = std::min((int)FRAME_GRID_COLS - 1, (int)ceil((x - minX + r) * invW));
        if (nMaxCellX < 0) return;
        const int nMinCellY = std::max(0, (int)floor((y - minY - r) * invH));
        if (nMinCellY >= FRAME_GRID_ROWS) return;
        const int nMaxCellY = std::min((int)FRAME_GRID_ROWS - 1, (int)ceil((y - minY + r) * invH));
        if (nMaxCellY < 0) return;
        const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
            for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
                const std::vector<int>& vCell = grid[ix][iy];
                for (size_t j = 0, jend = vCell.size(); j < jend; j++) {
                    const oracle_kp_t& kpUn = keys[vCell[j]];
                    if (bCheckLevels) {
                        if (kpUn.octave < minLevel) continue;
                        if (maxLevel >= 0)
                            if (kpUn.octave > maxLevel) continue;
                    }
                    const float distx = kpUn.x - x;
                    const float disty = kpUn.y - y;
                    if (fabs(distx) < r && fabs(disty) < r) vIndices.push_back(vCell[j]);
                }
            }
    }
};

extern "C" {

int oracle_descriptor_distance(const uint8_t* a, const uint8_t* b) { return DescriptorDistance(a, b); }

void oracle_hamming_knn2(const uint8_t* q, int nq, const uint8_t* m, int nm,
                         int32_t* best_idx, int32_t* best_d, int32_t* second_d)
{
    for (int i = 0; i < nq; i++) {
        int bestDist = 256, bestDist2 = 256, bestIdx = -1;
        for (int j = 0; j < nm; j++) {
            const int dist = DescriptorDistance(q + (size_t)i * 32, m + (size_t)j * 32);
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx = j; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        best_idx[i] = bestIdx; best_d[i] = bestDist; second_d[i] = bestDist2;
    }
}

oracle_frame* oracle_frame_create(const oracle_kp_t* kps, const uint8_t* desc, int n,
                                  float minX, float minY, float invW, float invH, const float* uRight)
{
    oracle_frame* F = new oracle_frame;
    F->N = n;
    F->keys.assign(kps, kps + n);
    F->desc.assign(desc, desc + (size_t)n * 32);
    F->uRight.assign(n, -1.f);
    if (uRight) F->uRight.assign(uRight, uRight + n);
    F->minX = minX; F->minY = minY; F->invW = invW; F->invH = invH;
    /* Frame::AssignFeaturesToGrid, src/Frame.cc:378-412 */
    for (int i = 0; i < n; i++) {
        int gx, gy;
        if (F->PosInGrid(F->keys[i], gx, gy)) F->grid[gx][gy].push_back(i);
    }
    return F;
}

void oracle_frame_destroy(oracle_frame* F) { delete F; }

int oracle_frame_features_in_area(const oracle_frame* F, float x, float y, float r, int minLevel, int maxLevel,
                                  int32_t* out, int cap)
{
    std::vector<int> v;
    F->GetFeaturesInArea(x, y, r, minLevel, maxLevel, v);
    for (int i = 0; i < (int)v.size() && i < cap; i++) out[i] = v[i];
    return (int)v.size();
}

/* src/ORBmatcher.cc:45-129 */
int oracle_search_by_projection(const oracle_frame* F, const float* scaleFactors, int nq,
                                const uint8_t* q_valid, const float* q_u, const float* q_v, const float* q_uR,
                                const int32_t* q_level, const float* q_viewcos, const uint8_t* q_desc,
                                const uint8_t* q_obs_pos, const uint8_t* kp_blocked,
                                float th, float nnratio,
                                int32_t* out_best_idx, int32_t* out_best_dist, int32_t* out_query_of_kp)
{
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    /* F.mvpMapPoints[idx] && ->Observations()>0, as it evolves through the loop */
    std::vector<uint8_t> blocked(F->N, 0);
    if (kp_blocked) blocked.assign(kp_blocked, kp_blocked + F->N);
    for (int i = 0; i < F->N; i++) out_query_of_kp[i] = -1;
    std::vector<int> vIndices;
    for (int iMP = 0; iMP < nq; iMP++) {
        out_best_idx[iMP] = -1; out_best_dist[iMP] = 256;
        if (!q_valid[iMP]) continue;
        const int nPredictedLevel = q_level[iMP];
        float r = RadiusByViewingCos(q_viewcos[iMP]);
        if (bFactor) r *= th;
        F->GetFeaturesInArea(q_u[iMP], q_v[iMP], r * scaleFactors[nPredictedLevel], nPredictedLevel - 1, nPredictedLevel, vIndices);
        if (vIndices.empty()) continue;
        const uint8_t* MPdescriptor = q_desc + (size_t)iMP * 32;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (size_t k = 0; k < vIndices.size(); k++) {
            const int idx = vIndices[k];
            if (blocked[idx]) continue;
            if (F->uRight[idx] > 0) {
                const float er = fabs(q_uR[iMP] - F->uRight[idx]);
                if (er > r * scaleFactors[nPredictedLevel]) continue;
            }
            const int dist = DescriptorDistance(MPdescriptor, &F->desc[(size_t)idx * 32]);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist;
                bestLevel2 = bestLevel; bestLevel = F->keys[idx].octave;
                bestIdx = idx;
            } else if (dist < bestDist2) {
                bestLevel2 = F->keys[idx].octave;
                bestDist2 = dist;
            }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            out_query_of_kp[bestIdx] = iMP;
            blocked[bestIdx] = q_obs_pos ? q_obs_pos[iMP] : 1;
            out_best_idx[iMP] = bestIdx; out_best_dist[iMP] = bestDist;
            nmatches++;
        }
    }
    return nmatches;
}

/* src/ORBmatcher.cc:1328-1470 (projection of Last's map points done by the caller) */
int oracle_search_by_projection_frame(const oracle_frame* Cur, const float* scaleFactors, int nq,
                                      const uint8_t* q_valid, const float* q_u, const float* q_v,
                                      const float* q_invz, const int32_t* q_octave, const float* q_angle,
                                      const uint8_t* q_desc, const uint8_t* q_obs_pos,
                                      const uint8_t* kp_blocked, float th, float mbf, int mode, int checkOri,
                                      int32_t* out_query_of_kp)
{
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    std::vector<uint8_t> blocked(Cur->N, 0);
    if (kp_blocked) blocked.assign(kp_blocked, kp_blocked + Cur->N);
    for (int i = 0; i < Cur->N; i++) out_query_of_kp[i] = -1;
    std::vector<int> vIndices2;
    for (int i = 0; i < nq; i++) {
        if (!q_valid[i]) continue;
        const float u = q_u[i], v = q_v[i], invzc = q_invz[i];
        int nLastOctave = q_octave[i];
        float radius = th * scaleFactors[nLastOctave];
        if (mode == 1) Cur->GetFeaturesInArea(u, v, radius, nLastOctave, -1, vIndices2);
        else if (mode == 2) Cur->GetFeaturesInArea(u, v, radius, 0, nLastOctave, vIndices2);
        else Cur->GetFeaturesInArea(u, v, radius, nLastOctave - 1, nLastOctave + 1, vIndices2);
        if (vIndices2.empty()) continue;
        const uint8_t* dMP = q_desc + (size_t)i * 32;
        int bestDist = 256, bestIdx2 = -1;
        for (size_t k = 0; k < vIndices2.size(); k++) {
            const int i2 = vIndices2[k];
            if (blocked[i2]) continue;
            if (Cur->uRight[i2] > 0) {
                const float ur = u - mbf * invzc;
                const float er = fabs(ur - Cur->uRight[i2]);
                if (er > radius) continue;
            }
            const int dist = DescriptorDistance(dMP, &Cur->desc[(size_t)i2 * 32]);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= TH_HIGH) {
            out_query_of_kp[bestIdx2] = i;
            blocked[bestIdx2] = q_obs_pos ? q_obs_pos[i] : 1;
            nmatches++;
            if (checkOri) rotHist[rot_bin(q_angle[i], Cur->keys[bestIdx2].angle)].push_back(bestIdx2);
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (size_t j = 0; j < rotHist[i].size(); j++) { out_query_of_kp[rotHist[i][j]] = -1; nmatches--; }
    }
    return nmatches;
}

/* src/ORBmatcher.cc:1667-1786 (prev_xy != NULL) and :1788-1899 (prev_xy == NULL) */
int oracle_birdview_match(const oracle_kp_t* kps1, const uint8_t* desc1, int n1, const oracle_frame* F2,
                          float* prev_xy, int windowSize, float nnratio, int checkOri, int32_t* vnMatches12)
{
    int nmatches = 0;
    for (int i = 0; i < n1; i++) vnMatches12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    std::vector<int> vMatchedDistance(F2->N, INT_MAX);
    std::vector<int> vnMatches21(F2->N, -1);
    std::vector<int> vIndices2;
    for (int i1 = 0; i1 < n1; i1++) {
        const oracle_kp_t& kp1 = kps1[i1];
        int level1 = kp1.octave;
        if (prev_xy) {
            if (level1 > 0) continue;
            F2->GetFeaturesInArea(prev_xy[2 * i1], prev_xy[2 * i1 + 1], (float)windowSize, level1, level1, vIndices2);
        } else {
            F2->GetFeaturesInArea(kp1.x, kp1.y, (float)windowSize, level1, level1, vIndices2);
        }
        if (vIndices2.empty()) continue;
        const uint8_t* d1 = desc1 + (size_t)i1 * 32;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (size_t k = 0; k < vIndices2.size(); k++) {
            const int i2 = vIndices2[k];
            int dist = DescriptorDistance(d1, &F2->desc[(size_t)i2 * 32]);
            if (vMatchedDistance[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= TH_LOW) {
            if (bestDist < (float)bestDist2 * nnratio) {
                if (vnMatches21[bestIdx2] >= 0) { vnMatches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
                vnMatches12[i1] = bestIdx2;
                vnMatches21[bestIdx2] = i1;
                vMatchedDistance[bestIdx2] = bestDist;
                nmatches++;
                if (checkOri) rotHist[rot_bin(kp1.angle, F2->keys[bestIdx2].angle)].push_back(i1);
            }
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) {
                int idx1 = rotHist[i][j];
                if (vnMatches12[idx1] >= 0) { vnMatches21[vnMatches12[idx1]] = -1; vnMatches12[idx1] = -1; nmatches--; }
            }
        }
    }
    if (prev_xy)
        for (int i1 = 0; i1 < n1; i1++)
            if (vnMatches12[i1] >= 0) { prev_xy[2 * i1] = F2->keys[vnMatches12[i1]].x; prev_xy[2 * i1 + 1] = F2->keys[vnMatches12[i1]].y; }
    return nmatches;
}

/* src/ORBmatcher.cc:2000-2114 */
int oracle_search_by_match_bird_kf(const oracle_kp_t* kf_kps, const uint8_t* has_mp, const uint8_t* mp_desc, int nk,
                                   const oracle_frame* F, float r, float nnratio, int checkOri, int32_t* out_mp_of_kp)
{
    for (int i = 0; i < F->N; i++) out_mp_of_kp[i] = -1;
    std::vector<int> vMatchedDistanceBird(F->N, INT_MAX);
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    std::vector<int> vIndices;
    for (int k = 0; k < nk; k++) {
        if (!has_mp[k]) continue;
        const oracle_kp_t& kp = kf_kps[k];
        F->GetFeaturesInArea(kp.x, kp.y, r, -1, -1, vIndices);
        if (vIndices.empty()) continue;
        const uint8_t* d1 = mp_desc + (size_t)k * 32;
        int bestDist = INT_MAX, bestLevel = -1, bestDist2 = INT_MAX, bestLevel2 = -1, bestIdx = -1;
        for (size_t j = 0; j < vIndices.size(); j++) {
            const int idx = vIndices[j];
            const int dist = DescriptorDistance(d1, &F->desc[(size_t)idx * 32]);
            if (vMatchedDistanceBird[idx] <= dist) continue;
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist;
                bestLevel2 = bestLevel; bestLevel = F->keys[idx].octave;
                bestIdx = idx;
            } else if (dist < bestDist2) {
                bestLevel2 = F->keys[idx].octave;
                bestDist2 = dist;
            }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel != bestLevel2 || bestDist < (float)bestDist2 * nnratio) {
                out_mp_of_kp[bestIdx] = k;
                vMatchedDistanceBird[bestIdx] = bestDist;
                if (checkOri) rotHist[rot_bin(kp.angle, F->keys[bestIdx].angle)].push_back(bestIdx);
                nmatches++;
            }
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) { out_mp_of_kp[rotHist[i][j]] = -1; nmatches--; }
        }
    }
    return nmatches;
}

/* src/ORBmatcher.cc:1923-1998 (projection through Tbc*Tcw done by the caller) */
int oracle_search_by_projection_bird(const oracle_frame* F, int nq, const uint8_t* q_valid,
                                     const float* q_x, const float* q_y, const uint8_t* q_desc,
                                     const uint8_t* q_obs_pos, const uint8_t* kp_blocked, float r, float nnratio,
                                     int32_t* out_query_of_kp)
{
    int nmatches = 0;
    std::vector<uint8_t> blocked(F->N, 0);
    if (kp_blocked) blocked.assign(kp_blocked, kp_blocked + F->N);
    for (int i = 0; i < F->N; i++) out_query_of_kp[i] = -1;
    std::vector<int> vIndices;
    for (int iMP = 0; iMP < nq; iMP++) {
        if (!q_valid[iMP]) continue;
        F->GetFeaturesInArea(q_x[iMP], q_y[iMP], r, -1, -1, vIndices);
        if (vIndices.empty()) continue;
        const uint8_t* MPdescriptor = q_desc + (size_t)iMP * 32;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (size_t k = 0; k < vIndices.size(); k++) {
            const int idx = vIndices[k];
            if (blocked[idx]) continue;
            const int dist = DescriptorDistance(MPdescriptor, &F->desc[(size_t)idx * 32]);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist;
                bestLevel2 = bestLevel; bestLevel = F->keys[idx].octave;
                bestIdx = idx;
            } else if (dist < bestDist2) {
                bestLevel2 = F->keys[idx].octave;
                bestDist2 = dist;
            }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            out_query_of_kp[bestIdx] = iMP;
            blocked[bestIdx] = q_obs_pos ? q_obs_pos[iMP] : 1;
            nmatches++;
        }
    }
    return nmatches;
}

/* src/ORBmatcher.cc:405-520: the loop BirdviewMatch(F1,F2,..,vPrevMatched,..) was copied from (identical
 * statements on mvKeysUn/mDescriptors/GetFeaturesInArea instead of the *Bird members) */
int oracle_search_for_initialization(const oracle_kp_t* kps1, const uint8_t* desc1, int n1, const oracle_frame* F2,
                                     float* prev_xy, int windowSize, float nnratio, int checkOri, int32_t* vnMatches12)
{
    return oracle_birdview_match(kps1, desc1, n1, F2, prev_xy, windowSize, nnratio, checkOri, vnMatches12);
}

/* Generic best-only windowed search; with the host-side geometry factored out this is the common loop of
 * SearchByProjection(Frame&,KeyFrame*,set,th,ORBdist) (:1472-1599), SearchByProjection(KeyFrame*,Scw,...) (:290-403),
 * Fuse (:825-975, :977-1100) and each direction of SearchBySim3 (:1102-1326).  flags: 1 BLOCK, 2 URCHECK, 4 CHI2, 8 ORI */
int oracle_search_window_best(const oracle_frame* F, int nq, const uint8_t* q_valid, const float* q_x, const float* q_y,
                              const float* q_r, const int32_t* q_minL, const int32_t* q_maxL, const uint8_t* q_desc,
                              const float* q_aux, const float* q_angle, const uint8_t* q_obs_pos, const uint8_t* kp_blocked,
                              const float* invLevelSigma2, int accTh, int flags,
                              int32_t* out_best_idx, int32_t* out_best_dist, int32_t* out_query_of_kp)
{
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    std::vector<uint8_t> blocked(F->N, 0);
    if (kp_blocked) blocked.assign(kp_blocked, kp_blocked + F->N);
    for (int i = 0; i < F->N; i++) out_query_of_kp[i] = -1;
    std::vector<int> vIndices;
    for (int i = 0; i < nq; i++) {
        out_best_idx[i] = -1; out_best_dist[i] = 256;
        if (q_valid && !q_valid[i]) continue;
        const float u = q_x[i], v = q_y[i], radius = q_r[i];
        F->GetFeaturesInArea(u, v, radius, q_minL[i], q_maxL[i], vIndices);
        if (vIndices.empty()) continue;
        const uint8_t* dMP = q_desc + (size_t)i * 32;
        int bestDist = 256, bestIdx = -1;
        for (size_t k = 0; k < vIndices.size(); k++) {
            const int idx = vIndices[k];
            if (blocked[idx]) continue;
            if ((flags & 2) && F->uRight[idx] > 0) {
                const float er = fabs(q_aux[i] - F->uRight[idx]);
                if (er > radius) continue;
            }
            if (flags & 4) {
                const oracle_kp_t& kp = F->keys[idx];
                const int kpLevel = kp.octave;
                if (F->uRight[idx] >= 0) {
                    const float ex = u - kp.x, ey = v - kp.y, er = q_aux[i] - F->uRight[idx];
                    const float e2 = ex * ex + ey * ey + er * er;
                    if (e2 * invLevelSigma2[kpLevel] > 7.8) continue;
                } else {
                    const float ex = u - kp.x, ey = v - kp.y;
                    const float e2 = ex * ex + ey * ey;
                    if (e2 * invLevelSigma2[kpLevel] > 5.99) continue;
                }
            }
            const int dist = DescriptorDistance(dMP, &F->desc[(size_t)idx * 32]);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        if (bestDist <= accTh) {
            out_best_idx[i] = bestIdx; out_best_dist[i] = bestDist;
            out_query_of_kp[bestIdx] = i;
            if (flags & 1) blocked[bestIdx] = q_obs_pos ? q_obs_pos[i] : 1;
            nmatches++;
            if (flags & 8) rotHist[rot_bin(q_angle[i], F->keys[bestIdx].angle)].push_back(bestIdx);
        }
    }
    if (flags & 8) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (size_t j = 0; j < rotHist[i].size(); j++) { out_query_of_kp[rotHist[i][j]] = -1; nmatches--; }
    }
    return nmatches;
}

/* src/ORBmatcher.cc:159-288 (kf_kf == 0) and :522-655 (kf_kf != 0) */
int oracle_search_by_bow(const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int n1,
                         const oracle_frame* F2, const uint8_t* valid2,
                         const int32_t* fv1_node, const int32_t* fv1_ptr, const int32_t* fv1_idx, int nn1,
                         const int32_t* fv2_node, const int32_t* fv2_ptr, const int32_t* fv2_idx, int nn2,
                         float nnratio, int checkOri, int kf_kf, int32_t* out)
{
    int nmatches = 0;
    const int n2 = F2->N;
    std::vector<int> vpMapPointMatches(n2, -1);     /* (KF,F): KF index whose MapPoint the F keypoint received */
    std::vector<int> vpMatches12(n1, -1);           /* (KF,KF) */
    std::vector<bool> vbMatched2(n2, false);
    std::vector<int> rotHist[HISTO_LENGTH];
    int f1 = 0, f2 = 0;
    while (f1 < nn1 && f2 < nn2) {
        if (fv1_node[f1] == fv2_node[f2]) {
            for (int i1 = fv1_ptr[f1]; i1 < fv1_ptr[f1 + 1]; i1++) {
                const int idx1 = fv1_idx[i1];
                if (!valid1[idx1]) continue;
                const uint8_t* d1 = desc1 + (size_t)idx1 * 32;
                int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                for (int i2 = fv2_ptr[f2]; i2 < fv2_ptr[f2 + 1]; i2++) {
                    const int idx2 = fv2_idx[i2];
                    if (kf_kf) {
                        if (vbMatched2[idx2] || !(valid2 ? valid2[idx2] : 1)) continue;
                    } else {
                        if (vpMapPointMatches[idx2] >= 0) continue;
                    }
                    const int dist = DescriptorDistance(d1, &F2->desc[(size_t)idx2 * 32]);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                const bool th = kf_kf ? (bestDist1 < TH_LOW) : (bestDist1 <= TH_LOW);
                if (th) {
                    if (static_cast<float>(bestDist1) < nnratio * static_cast<float>(bestDist2)) {
                        if (kf_kf) { vpMatches12[idx1] = bestIdx2; vbMatched2[bestIdx2] = true; }
                        else vpMapPointMatches[bestIdx2] = idx1;
                        if (checkOri) rotHist[rot_bin(angle1[idx1], F2->keys[bestIdx2].angle)].push_back(kf_kf ? idx1 : bestIdx2);
                        nmatches++;
                    }
                }
            }
            f1++; f2++;
        } else if (fv1_node[f1] < fv2_node[f2]) {
            while (f1 < nn1 && fv1_node[f1] < fv2_node[f2]) f1++;
        } else {
            while (f2 < nn2 && fv2_node[f2] < fv1_node[f1]) f2++;
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) {
                if (kf_kf) vpMatches12[rotHist[i][j]] = -1; else vpMapPointMatches[rotHist[i][j]] = -1;
                nmatches--;
            }
        }
    }
    if (kf_kf) for (int i = 0; i < n1; i++) out[i] = vpMatches12[i];
    else for (int i = 0; i < n2; i++) out[i] = vpMapPointMatches[i];
    return nmatches;
}

/* ORBmatcher::CheckDistEpipolarLine, src/ORBmatcher.cc:140-157 */
static bool CheckDistEpipolarLine(const oracle_kp_t& kp1, const oracle_kp_t& kp2, const float* F12, const float* levelSigma2_2)
{
    const float a = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
    const float b = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
    const float c = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
    const float num = a * kp2.x + b * kp2.y + c;
    const float den = a * a + b * b;
    if (den == 0) return false;
    const float dsqr = num * num / den;
    return dsqr < 3.84 * levelSigma2_2[kp2.octave];
}

/* src/ORBmatcher.cc:657-823 */
int oracle_search_for_triangulation(const oracle_kp_t* kps1, const uint8_t* desc1, const float* uR1, const uint8_t* has_mp1, int n1,
                                    const oracle_kp_t* kps2, const uint8_t* desc2, const float* uR2, const uint8_t* has_mp2, int n2,
                                    const int32_t* fv1_node, const int32_t* fv1_ptr, const int32_t* fv1_idx, int nn1,
                                    const int32_t* fv2_node, const int32_t* fv2_ptr, const int32_t* fv2_idx, int nn2,
                                    const float* F12, float ex, float ey,
                                    const float* scaleFactors2, const float* levelSigma2_2,
                                    int onlyStereo, int checkOri, int32_t* pairs)
{
    int nmatches = 0;
    std::vector<bool> vbMatched2(n2, false);
    std::vector<int> vMatches12(n1, -1);
    std::vector<int> rotHist[HISTO_LENGTH];
    int f1 = 0, f2 = 0;
    while (f1 < nn1 && f2 < nn2) {
        if (fv1_node[f1] == fv2_node[f2]) {
            for (int i1 = fv1_ptr[f1]; i1 < fv1_ptr[f1 + 1]; i1++) {
                const int idx1 = fv1_idx[i1];
                if (has_mp1[idx1]) continue;
                const bool bStereo1 = uR1 ? uR1[idx1] >= 0 : false;
                if (onlyStereo)
                    if (!bStereo1) continue;
                const oracle_kp_t& kp1 = kps1[idx1];
                const uint8_t* d1 = desc1 + (size_t)idx1 * 32;
                int bestDist = TH_LOW;
                int bestIdx2 = -1;
                for (int i2 = fv2_ptr[f2]; i2 < fv2_ptr[f2 + 1]; i2++) {
                    const int idx2 = fv2_idx[i2];
                    if (vbMatched2[idx2] || has_mp2[idx2]) continue;
                    const bool bStereo2 = uR2 ? uR2[idx2] >= 0 : false;
                    if (onlyStereo)
                        if (!bStereo2) continue;
                    const uint8_t* d2 = desc2 + (size_t)idx2 * 32;
                    const int dist = DescriptorDistance(d1, d2);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    const oracle_kp_t& kp2 = kps2[idx2];
                    if (!bStereo1 && !bStereo2) {
                        const float distex = ex - kp2.x;
                        const float distey = ey - kp2.y;
                        if (distex * distex + distey * distey < 100 * scaleFactors2[kp2.octave]) continue;
                    }
                    if (CheckDistEpipolarLine(kp1, kp2, F12, levelSigma2_2)) { bestIdx2 = idx2; bestDist = dist; }
                }
                if (bestIdx2 >= 0) {
                    const oracle_kp_t& kp2 = kps2[bestIdx2];
                    vMatches12[idx1] = bestIdx2;
                    nmatches++;
                    if (checkOri) rotHist[rot_bin(kp1.angle, kp2.angle)].push_back(idx1);
                }
            }
            f1++; f2++;
        } else if (fv1_node[f1] < fv2_node[f2]) {
            while (f1 < nn1 && fv1_node[f1] < fv2_node[f2]) f1++;   /* lower_bound */
        } else {
            while (f2 < nn2 && fv2_node[f2] < fv1_node[f1]) f2++;
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) { vMatches12[rotHist[i][j]] = -1; nmatches--; }
        }
    }
    int np = 0;
    for (int i = 0; i < n1; i++) {
        if (vMatches12[i] < 0) continue;
        pairs[2 * np] = i; pairs[2 * np + 1] = vMatches12[i];
        np++;
    }
    return nmatches;
}

}  // extern "C"


/* ---- DBoW2 vocabulary transform (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1139-1203, 1230-1271; FORB::distance
 * FORB.cpp:81-101; BowVector.cpp:34-46,62-84; FeatureVector.cpp:31-45), TF_IDF weighting + L1 scoring as used by
 * ORBVocabulary.  The tree is passed flattened: children CSR, node descriptors, leaf word ids and weights. ---- */
struct oracle_voc {
    int nNodes, L;
    std::vector<int> childPtr, childIdx, wordId;
    std::vector<uint8_t> desc;
    std::vector<double> weight;
};

extern "C" {

oracle_voc* oracle_voc_create(int nNodes, const int32_t* child_ptr, const int32_t* child_idx, const uint8_t* node_desc,
                              const int32_t* word_id, const double* weight, int L)
{
    oracle_voc* v = new oracle_voc;
    v->nNodes = nNodes; v->L = L;
    v->childPtr.assign(child_ptr, child_ptr + nNodes + 1);
    v->childIdx.assign(child_idx, child_idx + child_ptr[nNodes]);
    v->desc.assign(node_desc, node_desc + (size_t)nNodes * 32);
    v->wordId.assign(word_id, word_id + nNodes);
    v->weight.assign(weight, weight + nNodes);
    return v;
}

void oracle_voc_destroy(oracle_voc* v) { delete v; }

/* transform(features, BowVector, FeatureVector, levelsup); outputs: per-feature word/node, the BowVector as ascending
 * (word, value) and the FeatureVector as CSR over ascending node ids */
int oracle_voc_transform(const oracle_voc* V, const uint8_t* features, int n, int levelsup,
                         int32_t* out_word, int32_t* out_node,
                         int32_t* bow_word, double* bow_value, int32_t* n_words,
                         int32_t* fv_node, int32_t* fv_ptr, int32_t* fv_idx, int32_t* n_fv)
{
    std::map<unsigned, double> v;
    std::map<unsigned, std::vector<unsigned> > fv;
    for (int i_feature = 0; i_feature < n; i_feature++) {
        const uint8_t* feature = features + (size_t)i_feature * 32;
        const int nid_level = V->L - levelsup;
        unsigned nid = 0;
        unsigned final_id = 0;
        int current_level = 0;
        do {
            ++current_level;
            const int cb = V->childPtr[final_id], ce = V->childPtr[final_id + 1];
            final_id = V->childIdx[cb];
            double best_d = DescriptorDistance(feature, &V->desc[(size_t)final_id * 32]);
            for (int c = cb + 1; c < ce; c++) {
                const unsigned id = V->childIdx[c];
                double d = DescriptorDistance(feature, &V->desc[(size_t)id * 32]);
                if (d < best_d) { best_d = d; final_id = id; }
            }
            if (current_level == nid_level) nid = final_id;
        } while (V->childPtr[final_id] != V->childPtr[final_id + 1]);
        const unsigned word_id = V->wordId[final_id];
        const double w = V->weight[final_id];
        out_word[i_feature] = (int32_t)word_id; out_node[i_feature] = (int32_t)nid;
        if (w > 0) {
            v[word_id] += w;                 /* BowVector::addWeight */
            fv[nid].push_back(i_feature);    /* FeatureVector::addFeature */
        }
    }
    double norm = 0.0;                        /* BowVector::normalize(L1) */
    for (auto it = v.begin(); it != v.end(); ++it) norm += fabs(it->second);
    if (norm > 0.0)
        for (auto it = v.begin(); it != v.end(); ++it) it->second /= norm;
    int k = 0;
    for (auto it = v.begin(); it != v.end(); ++it, ++k) { bow_word[k] = (int32_t)it->first; bow_value[k] = it->second; }
    *n_words = k;
    int nn = 0, p = 0;
    for (auto it = fv.begin(); it != fv.end(); ++it, ++nn) {
        fv_node[nn] = (int32_t)it->first; fv_ptr[nn] = p;
        for (unsigned idx : it->second) fv_idx[p++] = (int32_t)idx;
    }
    fv_ptr[nn] = p;
    *n_fv = nn;
    return n;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// Frame::isInFrustum (reference src/Frame.cc:436-492) + MapPoint::PredictScale (src/MapPoint.cc:402-417) for an
// array of map points: the query generation of Tracking::SearchLocalPoints (src/Tracking.cc:1634-1648).
// cv::Mat arithmetic as OpenCV 4.x evaluates it for these shapes (checked against cv2.gemm / cv2.norm by
// tests/ref_py/frustum_py_ref.py): 3x3 * 3x1 + 3x1 through gemm's small-matrix path = float products summed left to
// right in float, then (float)((double)sum + (double)c); cv::norm = sqrt of a double sum of squares; Mat::dot
// accumulates in double.  Scalar float expressions as written, no FMA contraction (oracle policy).
// ---------------------------------------------------------------------------------------------------
extern "C" {

struct oracle_camera_pose {
    float Rcw[9], tcw[3], Ow[3];
    float fx, fy, cx, cy, mbf;
    float min_x, max_x, min_y, max_y;
    float log_scale_factor;
    int32_t n_levels;
};

int oracle_is_in_frustum(int n, const float* pos, const float* normal, const float* max_distance, const float* min_distance,
                         const uint8_t* candidate, const oracle_camera_pose* F, float viewingCosLimit,
                         uint8_t* out_in_view, float* out_u, float* out_v, float* out_uR, int32_t* out_level, float* out_viewcos)
{
    int nToMatch = 0;
    for (int i = 0; i < n; i++) {
        out_in_view[i] = 0; out_u[i] = 0; out_v[i] = 0; out_uR[i] = 0; out_level[i] = 0; out_viewcos[i] = 0;
        if (candidate && !candidate[i]) continue;                 // mnLastFrameSeen == current frame, or isBad()
        const float* P = pos + 3 * i;
        // 3D in camera coordinates: Pc = mRcw*P + mtcw
        float Pc[3];
        for (int r = 0; r < 3; r++) {
            float t = F->Rcw[3 * r] * P[0] + F->Rcw[3 * r + 1] * P[1];
            t = t + F->Rcw[3 * r + 2] * P[2];
            Pc[r] = (float)((double)t + (double)F->tcw[r]);
        }
        const float PcX = Pc[0], PcY = Pc[1], PcZ = Pc[2];
        if (PcZ < 0.0f) continue;
        const float invz = 1.0f / PcZ;
        const float u = F->fx * PcX * invz + F->cx;
        const float v = F->fy * PcY * invz + F->cy;
        if (u < F->min_x || u > F->max_x) continue;
        if (v < F->min_y || v > F->max_y) continue;
        const float maxDistance = 1.2f * max_distance[i];         // GetMaxDistanceInvariance, src/MapPoint.cc:380-383
        const float minDistance = 0.8f * min_distance[i];         // GetMinDistanceInvariance, :374-377
        const float PO[3] = {P[0] - F->Ow[0], P[1] - F->Ow[1], P[2] - F->Ow[2]};
        const float dist = (float)std::sqrt((double)PO[0] * PO[0] + (double)PO[1] * PO[1] + (double)PO[2] * PO[2]);
        if (dist < minDistance || dist > maxDistance) continue;
        const float* Pn = normal + 3 * i;
        const double dot = (double)PO[0] * Pn[0] + (double)PO[1] * Pn[1] + (double)PO[2] * Pn[2];
        const float viewCos = (float)(dot / dist);
        if (viewCos < viewingCosLimit) continue;
        // PredictScale
        const float ratio = max_distance[i] / dist;
        const float c = std::ceil(std::log(ratio) / F->log_scale_factor);
        // int conversion as x86 cvttss2si: out-of-range / NaN -> INT_MIN
        int nScale = (c > -2147483648.f && c < 2147483648.f) ? (int)c : INT32_MIN;
        if (nScale < 0) nScale = 0;
        else if (nScale >= F->n_levels) nScale = F->n_levels - 1;
        out_in_view[i] = 1;
        out_u[i] = u;
        out_uR[i] = u - F->mbf * invz;
        out_v[i] = v;
        out_level[i] = nScale;
        out_viewcos[i] = viewCos;
        nToMatch++;
    }
    return nToMatch;
}

// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:242-307) and MapPointBird::ComputeDistinctiveDescriptors
// (src/MapPointBird.cc:87-152), the selection step: for every landmark the descriptors of its observations form one
// group (CSR `group_ptr`); Distances[i][j] = DescriptorDistance, Distances[i][i] = 0; per row the sorted distances'
// element [0.5*(N-1)] is the median; the FIRST row with the strictly smallest median wins (:288-301).
void oracle_distinctive_descriptors(const uint8_t* desc, const int32_t* group_ptr, int n_groups,
                                    int32_t* best_idx, int32_t* best_median)
{
    std::vector<int> row;
    for (int g = 0; g < n_groups; g++) {
        const int b = group_ptr[g], N = group_ptr[g + 1] - b;
        best_idx[g] = -1; best_median[g] = -1;
        if (N <= 0) continue;                                   // observations.empty() / vDescriptors.empty(): nothing chosen
        int BestMedian = INT32_MAX, BestIdx = 0;
        for (int i = 0; i < N; i++) {
            row.assign(N, 0);
            for (int j = 0; j < N; j++) row[j] = i == j ? 0 : DescriptorDistance(desc + (size_t)(b + i) * 32, desc + (size_t)(b + j) * 32);
            std::sort(row.begin(), row.end());
            const int median = row[(size_t)(0.5 * (N - 1))];
            if (median < BestMedian) { BestMedian = median; BestIdx = i; }
        }
        best_idx[g] = BestIdx; best_median[g] = BestMedian;
    }
}

}  // extern "C"
