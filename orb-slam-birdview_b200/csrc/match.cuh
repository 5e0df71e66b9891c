// Matcher job descriptors shared by match.cu and api.cu.
#pragma once
#include "ctx.cuh"

namespace orbb200 {

enum WinMode { WM_PROJ = 0, WM_PROJ_FRAME = 1, WM_BIRD = 2, WM_BIRD_KF = 3, WM_PROJ_BIRD = 4, WM_BOW_KF_F = 5, WM_BOW_KF_KF = 6, WM_BEST = 7 };
enum WinFlags { WF_BLOCK = 1, WF_URCHECK = 2, WF_CHI2 = 4 };

// One windowed-search call (all pointers are device pointers).
struct WinJob {
    const FrameDev* frame;
    int nq, mode, levelMode, checkOri, kpCap;
    float th, nnratio, mbf;
    const float* scaleFactors;
    int nLevels;                // entries of scaleFactors / invLevelSigma2: queries and keypoints with a level outside [0, nLevels) are skipped
    const uint8_t* q_valid;     // may be null (all valid)
    const float* q_x;
    const float* q_y;
    const float* q_aux;         // WM_PROJ: projected uR; WM_PROJ_FRAME: 1/z
    int accTh, flags;           // WM_BEST: acceptance threshold, WinFlags
    const float* q_r;           // WM_BEST: per-query radius
    const int32_t* q_maxlevel;  // WM_BEST: max level (q_level = min level); BoW modes: end of the candidate range (q_level = begin)
    const int32_t* cand_idx;    // BoW modes: keypoint indices of the second frame grouped by vocabulary node
    const float* invLevelSigma2;   // WF_CHI2
    const int32_t* q_level;     // predicted level / octave
    const float* q_viewcos;     // WM_PROJ
    const float* q_angle;       // rotation histogram
    const uint8_t* q_desc;
    const uint8_t* q_obs_pos;   // may be null (all 1)
    const uint8_t* kp_blocked;  // may be null (none)
    int* scratch;               // win_scratch_ints(kpCap, nq) ints
    int32_t* out_best_idx;      // [nq] or null
    int32_t* out_best_dist;     // [nq] or null
    int32_t* out_per_kp;        // [n] query assigned to each keypoint (all modes but WM_BIRD)
    int32_t* out_per_query;     // [nq] WM_BIRD: vnMatches12
    int32_t* out_nmatches;
};

constexpr int WM_LISTCAP = 24;   // cached candidates per query (more -> that query re-scans its window)
// owner, lastOwner [kpCap]; choice, cdist, newChoice, newCdist, nextq, qbin, ccount [nq]; clist int2[nq][WM_LISTCAP]
__host__ __device__ static inline size_t win_clist_offset(int kpCap, int nq) { return ((size_t)2 * kpCap + (size_t)7 * nq + 1) & ~(size_t)1; }
static inline size_t win_scratch_ints(int kpCap, int nq) { return win_clist_offset(kpCap, nq) + (size_t)2 * WM_LISTCAP * nq + 2; }

struct TriJob {
    const orbb200_kp_t* kps1; const uint8_t* desc1; const float* uR1; const uint8_t* has_mp1; int n1;
    const orbb200_kp_t* kps2; const uint8_t* desc2; const float* uR2; const uint8_t* has_mp2; int n2;
    const int32_t* fv2_idx;
    const int32_t* item_idx1; const int32_t* item_b2; const int32_t* item_e2; int nItems;
    const float* F12; float ex, ey;
    const float* scaleFactors2; const float* levelSigma2_2; int nLevels2;
    int onlyStereo, checkOri;
    int32_t* match12;           // [n1]
    int32_t* pairs;             // [n1][2]
    int32_t* npairs;
};

int knn2_queries_per_thread(int nq, int nm);
void launch_knn2(Ctx& c, const uint8_t* d_q, int nq, const uint8_t* d_m, int nm, int nsplit, int4* d_partial,
                 int32_t* bi, int32_t* bd, int32_t* sd);
void launch_popc_peak(Ctx& c, uint32_t* d_out, int blocks, int iters);
void launch_distinctive(Ctx& c, const uint8_t* d_desc, const int32_t* d_groupPtr, int nGroups, int32_t* d_bestIdx, int32_t* d_bestMedian);
void launch_grid_build(Ctx& c, const FrameDev* d_frames, int nframes);
void launch_features_in_area(Ctx& c, const FrameDev* d_frame, float x, float y, float r, int minLevel, int maxLevel,
                             int32_t* d_out, int cap, int32_t* d_count);
void launch_window_match(Ctx& c, const WinJob* d_jobs, int njobs, int maxNq, int maxKpCap);
void launch_triangulation(Ctx& c, const TriJob& J);

// Frame::isInFrustum over a device-resident map
struct FrustumJob {
    orbb200_camera_pose pose;
    const orbb200_camera_pose* poses;   // device array of nFrames poses (batched form), or null: `pose`
    int nFrames;
    float cosLimit;
    int n;
    const float* pos; const float* normal; const float* maxDist; const float* minDist; const uint8_t* candidate;
    uint8_t* inView; float* u; float* v; float* uR; int32_t* level; float* viewcos; int32_t* count;
};
void launch_frustum(Ctx& c, const FrustumJob& J);

}  // namespace orbb200
