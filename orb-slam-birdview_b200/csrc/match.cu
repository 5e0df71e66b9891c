// Matching kernels: 256-bit Hamming brute force (best / second best), device lookup grid, and the
// windowed greedy searches of ORBmatcher restated as parallel fixed-point iterations.
#include <algorithm>

#include "match.cuh"

namespace orbb200 {

__device__ __forceinline__ int hamming256(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1)
{
    // ORBmatcher::DescriptorDistance (reference src/ORBmatcher.cc:1647-1663): popcount of the 256-bit XOR
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// ---------------------------------------------------------------------------------------------------
// Brute force knn2: grid = (map splits, query tiles).  Each thread keeps KN_QPT query descriptors in
// registers; map descriptors are staged in shared memory tiles and read as warp-wide broadcasts.
// Partial (best_d, best_idx, second_d) per (split, query) are merged by knn2_merge_kernel in split order
// so that the first minimum in index order wins, exactly like a sequential scan with strict '<'.
// ---------------------------------------------------------------------------------------------------
constexpr int KN_THREADS = 128;
constexpr int KN_QPT = 4;        // queries per thread when there is enough work to fill the GPU with it
constexpr int KN_MT = 128;       // map descriptors per shared-memory tile

// KN_QPT queries per thread amortise a shared-memory broadcast over four distances; small problems (2k x 2k: 7 us of POPC work) do not
// fill 148 SMs that way and run with one query per thread instead -- four times the CTAs, a quarter of the chain per thread.
template <int KN_QPT>
__global__ void __launch_bounds__(KN_THREADS) knn2_kernel(const uint4* __restrict__ q, int nq, const uint4* __restrict__ m, int nm,
                                                          int mPerSplit, int4* __restrict__ partial)
{
    __shared__ uint4 sM[2][KN_MT][2];
    pdl_launch_dependents();
    const int tid = threadIdx.x;
    const int split = blockIdx.x;
    const int mBeg = split * mPerSplit, mEnd = min(nm, mBeg + mPerSplit);
    const int qBase = blockIdx.y * (KN_THREADS * KN_QPT);

    uint4 qa[KN_QPT], qb[KN_QPT];
    int best[KN_QPT], second[KN_QPT], bidx[KN_QPT];
#pragma unroll
    for (int j = 0; j < KN_QPT; j++) {
        const int qi = min(qBase + j * KN_THREADS + tid, nq - 1);
        qa[j] = __ldg(q + 2 * (size_t)qi);
        qb[j] = __ldg(q + 2 * (size_t)qi + 1);
        best[j] = 256; second[j] = 256; bidx[j] = -1;
    }
    int buf = 0;
    // prefetch first tile
    {
        const int mi = mBeg + tid;
        if (mi < mEnd) { sM[0][tid][0] = __ldg(m + 2 * (size_t)mi); sM[0][tid][1] = __ldg(m + 2 * (size_t)mi + 1); }
    }
    __syncthreads();
    for (int t0 = mBeg; t0 < mEnd; t0 += KN_MT) {
        const int cnt = min(KN_MT, mEnd - t0);
        // prefetch next tile into the other buffer
        const int nmi = t0 + KN_MT + tid;
        uint4 n0, n1;
        const bool havNext = nmi < mEnd;
        if (havNext) { n0 = __ldg(m + 2 * (size_t)nmi); n1 = __ldg(m + 2 * (size_t)nmi + 1); }
#pragma unroll 4
        for (int i = 0; i < cnt; i++) {
            const uint4 m0 = sM[buf][i][0], m1 = sM[buf][i][1];
#pragma unroll
            for (int j = 0; j < KN_QPT; j++) {
                const int d = hamming256(qa[j], qb[j], m0, m1);
                if (d < best[j]) { second[j] = best[j]; best[j] = d; bidx[j] = t0 + i; }
                else second[j] = min(second[j], d);
            }
        }
        if (havNext) { sM[buf ^ 1][tid][0] = n0; sM[buf ^ 1][tid][1] = n1; }
        __syncthreads();
        buf ^= 1;
    }
#pragma unroll
    for (int j = 0; j < KN_QPT; j++) {
        const int qi = qBase + j * KN_THREADS + tid;
        if (qi < nq) partial[(size_t)split * nq + qi] = make_int4(best[j], bidx[j], second[j], 0);
    }
}

// One warp per query: lane l folds the contiguous split range [l*c, (l+1)*c) in order, then the lanes are folded pairwise with
// the earlier range on the left -- (best, idx, second) with "first minimum wins" is an associative fold, so the result equals the
// sequential scan over the splits (and therefore over the map descriptors) with strict '<'.
__global__ void __launch_bounds__(256) knn2_merge_kernel(const int4* __restrict__ partial, int nq, int nsplit,
                                                         int32_t* __restrict__ bi, int32_t* __restrict__ bd, int32_t* __restrict__ sd)
{
    const int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    pdl_wait();                                             // launched programmatically behind knn2_kernel
    if (qi >= nq) return;                                   // whole warps leave together
    const int per = (nsplit + 31) >> 5;
    int best = 256, second = 256, idx = -1;
    for (int s = lane * per, e = min(s + per, nsplit); s < e; s++) {
        const int4 p = __ldg(partial + (size_t)s * nq + qi);
        if (p.x < best) { second = min(best, p.z); best = p.x; idx = p.y; }
        else second = min(second, p.x);     // p.z >= p.x
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {                      // lane l (earlier splits) <- lane l + o (later splits)
        const int b2 = __shfl_down_sync(0xffffffffu, best, o), s2 = __shfl_down_sync(0xffffffffu, second, o), i2 = __shfl_down_sync(0xffffffffu, idx, o);
        if (lane + o < 32) {
            if (b2 < best) { second = min(best, s2); best = b2; idx = i2; }
            else second = min(second, b2);
        }
    }
    if (lane == 0) { bi[qi] = idx; bd[qi] = best; sd[qi] = second; }
}

// one query per thread below ~64 M descriptor pairs (see knn2_kernel)
int knn2_queries_per_thread(int nq, int nm) { return (long long)nq * nm < (64ll << 20) ? 1 : KN_QPT; }

void launch_knn2(Ctx& c, const uint8_t* d_q, int nq, const uint8_t* d_m, int nm, int nsplit, int4* d_partial,
                 int32_t* bi, int32_t* bd, int32_t* sd)
{
    const int mPerSplit = (nm + nsplit - 1) / nsplit;
    const int qpt = knn2_queries_per_thread(nq, nm);
    dim3 grid(nsplit, (nq + KN_THREADS * qpt - 1) / (KN_THREADS * qpt));
    if (qpt == 1)
        knn2_kernel<1><<<grid, KN_THREADS, 0, c.stream>>>(reinterpret_cast<const uint4*>(d_q), nq, reinterpret_cast<const uint4*>(d_m), nm, mPerSplit, d_partial);
    else
        knn2_kernel<KN_QPT><<<grid, KN_THREADS, 0, c.stream>>>(reinterpret_cast<const uint4*>(d_q), nq, reinterpret_cast<const uint4*>(d_m), nm, mPerSplit, d_partial);
    launch_chain(c.pdl, knn2_merge_kernel, dim3((nq + 7) / 8), dim3(256), 0, c.stream, (const int4*)d_partial, nq, nsplit, bi, bd, sd);
    c.launches += 2;
}

// POPC issue-rate micro-benchmark: the denominator of the matching roofline.
__global__ void __launch_bounds__(256) popc_peak_kernel(uint32_t* out, int iters, uint32_t seed)
{
    uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x, c0 = a ^ b, d0 = a + b;
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            s0 += __popc(a); s1 += __popc(b); s2 += __popc(c0); s3 += __popc(d0);
            a += s3; b += s0; c0 += s1; d0 += s2;     // keep the chain data dependent, 4 independent streams
        }
    }
    if ((s0 ^ s1 ^ s2 ^ s3) == 0x12345678u) out[0] = a;
}

void launch_popc_peak(Ctx& c, uint32_t* d_out, int blocks, int iters)
{
    popc_peak_kernel<<<blocks, 256, 0, c.stream>>>(d_out, iters, 12345u);
    c.launches++;
}

// ---------------------------------------------------------------------------------------------------
// Lookup grid: Frame::AssignFeaturesToGrid / PosInGrid[Birdview] (reference src/Frame.cc:378-412,
// 549-559, 879-889).  CSR over 64x48 cells, column-major cell id = ix*48+iy (the scan order of
// GetFeaturesInArea), items in ascending keypoint index inside each cell.  One CTA per frame.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int frame_count(const FrameDev& F) { return F.n_ptr ? min(*F.n_ptr, F.n) : F.n; }

__device__ __forceinline__ int grid_cell_of(const FrameDev& F, const orbb200_kp_t& kp)
{
    const int px = (int)roundf(__fmul_rn(__fsub_rn(kp.x, F.minX), F.invW));
    const int py = (int)roundf(__fmul_rn(__fsub_rn(kp.y, F.minY), F.invH));
    if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) return -1;
    return px * GRID_ROWS + py;
}

__global__ void __launch_bounds__(1024) grid_build_kernel(const FrameDev* __restrict__ frames)
{
    __shared__ int sCnt[GRID_CELLS];
    __shared__ int warpTot[32];
    __shared__ int sTotal;                  // keypoints that fell inside the grid (<= n: PosInGrid rejects the rest)
    // the per-cell item lists are put in order by insertion sorts: chains of dependent reads and writes, in shared memory when the
    // frame's keypoints fit (they do for every extractor budget in use), in the global list otherwise
    constexpr int GB_SMEM_ITEMS = 6144;
    __shared__ int sItems[GB_SMEM_ITEMS];
    const FrameDev F = frames[blockIdx.x];
    const int n = frame_count(F);
    const int tid = threadIdx.x;
    int* const items = n <= GB_SMEM_ITEMS ? sItems : F.cellItems;
    for (int i = tid; i < GRID_CELLS; i += 1024) sCnt[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 1024) {
        const int cid = grid_cell_of(F, F.kps[i]);
        if (cid >= 0) atomicAdd(&sCnt[cid], 1);
    }
    __syncthreads();
    // exclusive scan over 3072 cells: 3 per thread
    int v[3], s = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) { v[k] = sCnt[tid * 3 + k]; s += v[k]; }
    int x = s;
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) warpTot[wid] = x;
    __syncthreads();
    if (wid == 0) {
        int t = warpTot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
        warpTot[lane] = t;
    }
    __syncthreads();
    int ex = (wid ? warpTot[wid - 1] : 0) + x - s;
#pragma unroll
    for (int k = 0; k < 3; k++) { F.cellStart[tid * 3 + k] = ex; sCnt[tid * 3 + k] = ex; ex += v[k]; }
    if (tid == 1023) { F.cellStart[GRID_CELLS] = ex; sTotal = ex; }
    __syncthreads();
    for (int i = tid; i < n; i += 1024) {
        const int cid = grid_cell_of(F, F.kps[i]);
        if (cid >= 0) items[atomicAdd(&sCnt[cid], 1)] = i;
    }
    __syncthreads();
    // restore ascending index order inside each cell (lists are a few items long)
    for (int cidx = tid; cidx < GRID_CELLS; cidx += 1024) {
        const int e = sCnt[cidx], b = cidx ? sCnt[cidx - 1] : 0;      // after the scatter sCnt[c] is the end of cell c = the start of cell c + 1
        for (int i = b + 1; i < e; i++) {
            const int key = items[i];
            int j = i - 1;
            while (j >= b && items[j] > key) { items[j + 1] = items[j]; j--; }
            items[j + 1] = key;
        }
    }
    __syncthreads();
    // position-sorted copy of what a window scan needs of each keypoint (no index -> keypoint indirection in the scan)
    // (only the sTotal items the CSR holds: cellItems beyond them is unwritten memory)
    const int nItems = sTotal;
    for (int k = tid; k < nItems; k += 1024) {
        const int idx = items[k];
        if (items != F.cellItems) F.cellItems[k] = idx;
        const orbb200_kp_t kp = F.kps[idx];
        F.cellKp[k] = make_int4(__float_as_int(kp.x), __float_as_int(kp.y), kp.octave, idx);
    }
}

void launch_grid_build(Ctx& c, const FrameDev* d_frames, int nframes)
{
    grid_build_kernel<<<nframes, 1024, 0, c.stream>>>(d_frames);
    c.launches++;
}

// ---------------------------------------------------------------------------------------------------
// Windowed searches.  A "job" is one reference matcher call: a frame (grid) and nq queries.  One CTA per
// job.  The reference loops are sequential because later queries see earlier assignments:
//   BLOCK semantics  (SearchByProjection x2, SearchByProjectionBird): a keypoint that an earlier accepted
//                    query with Observations()>0 took is skipped;
//   DIST semantics   (BirdviewMatch x2, SearchByMatchBird(KF,F)): a keypoint is skipped when
//                    vMatchedDistance[idx] <= dist, i.e. when an earlier accepted query matched it at <= dist.
// Both make query q a function of the final choices of queries < q only, so iterating "recompute every
// query from the previous round's choices" reaches the unique fixed point = the sequential result;
// rounds needed = longest dependency chain + 1 (2-4 in practice).
// ---------------------------------------------------------------------------------------------------
struct Sel { int best, second, bestIdx, bestLevel, secondLevel; };

template <class Visit>
__device__ __forceinline__ void scan_window(const FrameDev& F, float x, float y, float r, int minLevel, int maxLevel, Visit&& visit)
{
    // Frame::GetFeaturesInArea[Birdview] (reference src/Frame.cc:494-547, 891-944)
    const float fx = __fsub_rn(x, F.minX), fy = __fsub_rn(y, F.minY);
    const int nMinCellX = max(0, (int)floorf(__fmul_rn(__fsub_rn(fx, r), F.invW)));
    if (nMinCellX >= GRID_COLS) return;
    const int nMaxCellX = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(fx, r), F.invW)));
    if (nMaxCellX < 0) return;
    const int nMinCellY = max(0, (int)floorf(__fmul_rn(__fsub_rn(fy, r), F.invH)));
    if (nMinCellY >= GRID_ROWS) return;
    const int nMaxCellY = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(fy, r), F.invH)));
    if (nMaxCellY < 0) return;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++) {
        // cells (ix, nMinCellY..nMaxCellY) are contiguous in the column-major CSR
        const int b = F.cellStart[ix * GRID_ROWS + nMinCellY], e = F.cellStart[ix * GRID_ROWS + nMaxCellY + 1];
        auto item = [&](const int4 it) {
            orbb200_kp_t kp;                      // the fields a visitor may read: position and octave
            kp.x = __int_as_float(it.x); kp.y = __int_as_float(it.y); kp.octave = it.z;
            kp.size = 0.f; kp.angle = 0.f; kp.response = 0.f; kp.class_id = -1;
            if (bCheckLevels) {
                if (kp.octave < minLevel) return;
                if (maxLevel >= 0 && kp.octave > maxLevel) return;
            }
            const float dx = __fsub_rn(kp.x, x), dy = __fsub_rn(kp.y, y);
            if (fabsf(dx) < r && fabsf(dy) < r) visit(it.w, kp);
        };
        // two items in flight: the slice is contiguous, so the second load does not depend on the first
        for (int k = b; k < e; k += 2) {
            const int4 it0 = F.cellKp[k];
            const int4 it1 = F.cellKp[min(k + 1, e - 1)];
            item(it0);
            if (k + 1 < e) item(it1);
        }
    }
}

__global__ void __launch_bounds__(256) features_in_area_kernel(const FrameDev* frames, float x, float y, float r, int minLevel, int maxLevel,
                                                               int32_t* out, int cap, int32_t* count)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const FrameDev F = frames[0];
    int n = 0;
    scan_window(F, x, y, r, minLevel, maxLevel, [&](int idx, const orbb200_kp_t&) { if (n < cap) out[n] = idx; n++; });
    *count = n;
}

void launch_features_in_area(Ctx& c, const FrameDev* d_frame, float x, float y, float r, int minLevel, int maxLevel,
                             int32_t* d_out, int cap, int32_t* d_count)
{
    features_in_area_kernel<<<1, 32, 0, c.stream>>>(d_frame, x, y, r, minLevel, maxLevel, d_out, cap, d_count);
    c.launches++;
}

__device__ __forceinline__ int rot_bin(float a1, float a2)
{
    // rot = a1 - a2 (+360 if <0); bin = round(rot * (1/HISTO_LENGTH)) (reference src/ORBmatcher.cc:236-246)
    constexpr float factor = 1.0f / HISTO_LENGTH;
    float rot = __fsub_rn(a1, a2);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, factor));
    if (bin == HISTO_LENGTH) bin = 0;
    return min(max(bin, 0), HISTO_LENGTH - 1);
}

// ORBmatcher::ComputeThreeMaxima (reference src/ORBmatcher.cc:1601-1642)
__device__ void three_maxima(const int* histo, int& ind1, int& ind2, int& ind3)
{
    int max1 = 0, max2 = 0, max3 = 0;
    ind1 = ind2 = ind3 = -1;
    for (int i = 0; i < HISTO_LENGTH; i++) {
        const int s = histo[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
}

constexpr int WM_THREADS = 1024;     // one CTA per call: the rounds are latency chains, more threads = more of them in flight (512: 0.129 ms, 1024: 0.102 ms per step)
constexpr int WC_THREADS = 128;

__device__ __forceinline__ bool is_dist_sem(int mode) { return mode == WM_BIRD || mode == WM_BIRD_KF; }
__device__ __forceinline__ bool is_bow(int mode) { return mode == WM_BOW_KF_F || mode == WM_BOW_KF_KF; }

// Per-mode search window of query q (reference: SearchByProjection :62-69, :1385-1398; BirdviewMatch :1684-1688,
// :1807; SearchByMatchBird :2029; SearchByProjectionBird :1948).  Returns false when the query is skipped.
struct Window { float x, y, r, urRef, urTol; int minL, maxL; bool urCheck; };

__device__ __forceinline__ bool query_window(const WinJob& J, int q, Window& W)
{
    if (J.q_valid != nullptr && !J.q_valid[q]) return false;
    W.minL = -1; W.maxL = -1; W.urCheck = false; W.urRef = 0.f; W.urTol = 0.f;
    const int mode = J.mode;
    if (is_bow(mode)) return true;
    W.x = J.q_x[q]; W.y = J.q_y[q];
    if (mode == WM_PROJ) {
        const int lvl = J.q_level[q];
        if ((unsigned)lvl >= (unsigned)J.nLevels) return false;         // outside the pyramid: no scale factor to read
        float r = J.q_viewcos[q] > 0.998 ? 2.5f : 4.0f;                 // RadiusByViewingCos, :131-137
        if (J.th != 1.0f) r = __fmul_rn(r, J.th);
        W.r = __fmul_rn(r, J.scaleFactors[lvl]);
        W.minL = lvl - 1; W.maxL = lvl;
        W.urCheck = true; W.urRef = J.q_aux[q]; W.urTol = W.r;
    } else if (mode == WM_PROJ_FRAME) {
        const int oct = J.q_level[q];
        if ((unsigned)oct >= (unsigned)J.nLevels) return false;
        W.r = __fmul_rn(J.th, J.scaleFactors[oct]);
        if (J.levelMode == 1) { W.minL = oct; W.maxL = -1; }
        else if (J.levelMode == 2) { W.minL = 0; W.maxL = oct; }
        else { W.minL = oct - 1; W.maxL = oct + 1; }
        W.urCheck = true; W.urRef = __fsub_rn(W.x, __fmul_rn(J.mbf, J.q_aux[q])); W.urTol = W.r;
    } else if (mode == WM_BIRD) {
        const int lvl = J.q_level[q];
        if (J.levelMode == 1 && lvl > 0) return false;                  // prevMatched variant: octave 0 only
        W.r = J.th; W.minL = lvl; W.maxL = lvl;
    } else if (mode == WM_BEST) {
        // generic best-only search: per-query radius and level range from the host adapter
        W.r = J.q_r[q]; W.minL = J.q_level[q]; W.maxL = J.q_maxlevel[q];
        if (J.flags & WF_URCHECK) { W.urCheck = true; W.urRef = J.q_aux[q]; W.urTol = W.r; }
    } else {   // WM_BIRD_KF, WM_PROJ_BIRD
        W.r = J.th;
    }
    return true;
}

// The filters of a window candidate that depend neither on the loop-carried state nor on the descriptor (the modes without
// "distance" semantics): blocked keypoints, the stereo-consistency test, Fuse's reprojection gate.
__device__ __forceinline__ bool keypoint_passes(const WinJob& J, const FrameDev& F, int q, const Window& W, int mode, int idx, float kx, float ky, int octave)
{
    if (J.kp_blocked && J.kp_blocked[idx]) return false;
    if (W.urCheck && F.uRight) {
        const float ur = F.uRight[idx];
        if (ur > 0 && fabsf(__fsub_rn(W.urRef, ur)) > W.urTol) return false;
    }
    if (mode == WM_BEST && (J.flags & WF_CHI2)) {
        // reprojection gate of Fuse (:913-936): 7.8 with a stereo observation, 5.99 without
        const float ex = __fsub_rn(W.x, kx), ey = __fsub_rn(W.y, ky);
        float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
        const float kpr = F.uRight ? F.uRight[idx] : -1.f;
        if ((unsigned)octave >= (unsigned)J.nLevels) return false;  // keypoint outside the pyramid: no sigma to gate with
        const float inv = J.invLevelSigma2[octave];
        if (kpr >= 0) {
            const float er = __fsub_rn(J.q_aux[q], kpr);
            e2 = __fadd_rn(e2, __fmul_rn(er, er));
            if ((double)__fmul_rn(e2, inv) > 7.8) return false;
        } else {
            if ((double)__fmul_rn(e2, inv) > 5.99) return false;
        }
    }
    return true;
}

// Enumerate, in the reference's scan order, the candidates of query q that pass every filter which does not
// depend on the loop-carried state, with their Hamming distance: visit(idx, dist, octave).
template <class Visit>
__device__ __forceinline__ void static_candidates(const WinJob& J, const FrameDev& F, int q, const Window& W, Visit&& visit)
{
    const int mode = J.mode;
    const uint4* qd = reinterpret_cast<const uint4*>(J.q_desc) + 2 * (size_t)q;
    const uint4 qa = __ldg(qd), qb = __ldg(qd + 1);
    if (is_bow(mode)) {
        // SearchByBoW (:159-288, :522-655): every keypoint of the same vocabulary node, node list order
        for (int p = J.q_level[q]; p < J.q_maxlevel[q]; p++) {
            const int idx = J.cand_idx[p];
            if (J.kp_blocked && J.kp_blocked[idx]) continue;         // (KF,KF): no / bad MapPoint on the KF2 keypoint
            const uint4* kd = reinterpret_cast<const uint4*>(F.desc) + 2 * (size_t)idx;
            visit(idx, hamming256(qa, qb, kd[0], kd[1]), 0);
        }
        return;
    }
    const bool distSem = is_dist_sem(mode);
    scan_window(F, W.x, W.y, W.r, W.minL, W.maxL, [&](int idx, const orbb200_kp_t& kp) {
        if (!distSem && !keypoint_passes(J, F, q, W, mode, idx, kp.x, kp.y, kp.octave)) return;
        const uint4* kd = reinterpret_cast<const uint4*>(F.desc) + 2 * (size_t)idx;
        visit(idx, hamming256(qa, qb, kd[0], kd[1]), kp.octave);
    });
}

// Phase 1 (fully parallel, thread per query): enumerate the static candidates once and cache
// (index, distance | level<<16) in scan order.  Queries with more than WM_LISTCAP survivors are re-scanned
// in phase 2.
__global__ void __launch_bounds__(WC_THREADS) window_cands_kernel(const WinJob* __restrict__ jobs)
{
    const WinJob J = jobs[blockIdx.y];
    const int q = blockIdx.x * WC_THREADS + threadIdx.x;
    if (q >= J.nq) return;
    const FrameDev F = *J.frame;
    int* ccount = J.scratch + 2 * J.kpCap + 6 * J.nq;
    int2* clist = reinterpret_cast<int2*>(J.scratch + win_clist_offset(J.kpCap, J.nq)) + (size_t)q * WM_LISTCAP;
    Window W;
    int n = -1;     // skipped query
    if (query_window(J, q, W)) {
        n = 0;
        static_candidates(J, F, q, W, [&](int idx, int d, int level) {
            if (n < WM_LISTCAP) clist[n] = make_int2(idx, d | (level << 16));
            n++;
        });
    }
    ccount[q] = n;
}

// The same phase for a few calls (one frame per call: a few thousand queries on an otherwise idle GPU), ONE WARP per query.  A
// thread-per-query scan is a chain of dependent L2 round trips -- cell bounds, items two at a time, per-keypoint flags, a 32-byte
// descriptor per surviving item -- and the slowest thread sets the kernel's duration (40 us for 3000 map points).  Here the lanes take
// 32 consecutive items of a column slice at once: one round trip for the items, one for the flags, one for the descriptors, and a
// ballot appends the survivors in scan order (the list is the one the thread-per-query kernel writes).
constexpr int WCW_WARPS = 8;

__global__ void __launch_bounds__(WCW_WARPS * 32) window_cands_warp_kernel(const WinJob* __restrict__ jobs)
{
    const WinJob J = jobs[blockIdx.y];
    const int q = blockIdx.x * WCW_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= J.nq) return;                                  // whole warps leave together
    const FrameDev F = *J.frame;
    int* ccount = J.scratch + 2 * J.kpCap + 6 * J.nq;
    int2* clist = reinterpret_cast<int2*>(J.scratch + win_clist_offset(J.kpCap, J.nq)) + (size_t)q * WM_LISTCAP;
    const unsigned ltmask = (1u << lane) - 1u;
    Window W;
    int n = -1;     // skipped query
    if (query_window(J, q, W)) {                            // (warp-uniform: every lane evaluates the same query)
        n = 0;
        const int mode = J.mode;
        const uint4* qd = reinterpret_cast<const uint4*>(J.q_desc) + 2 * (size_t)q;
        const uint4 qa = __ldg(qd), qb = __ldg(qd + 1);
        auto append = [&](bool pass, int idx, int level) {
            int d = 0;
            if (pass) {
                const uint4* kd = reinterpret_cast<const uint4*>(F.desc) + 2 * (size_t)idx;
                d = hamming256(qa, qb, kd[0], kd[1]);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, pass);
            const int pos = n + __popc(bal & ltmask);
            if (pass && pos < WM_LISTCAP) clist[pos] = make_int2(idx, d | (level << 16));
            n += __popc(bal);
        };
        if (is_bow(mode)) {
            const int p0 = J.q_level[q], p1 = J.q_maxlevel[q];
            for (int pb = p0; pb < p1; pb += 32) {
                const int p = pb + lane;
                const int idx = p < p1 ? J.cand_idx[p] : 0;
                const bool pass = p < p1 && !(J.kp_blocked && J.kp_blocked[idx]);
                append(pass, idx, 0);
            }
        } else {
            // Frame::GetFeaturesInArea[Birdview] (reference src/Frame.cc:494-547, 891-944), as in scan_window
            const float x = W.x, y = W.y, r = W.r;
            const int minLevel = W.minL, maxLevel = W.maxL;
            const bool distSem = is_dist_sem(mode);
            const float fx = __fsub_rn(x, F.minX), fy = __fsub_rn(y, F.minY);
            const int nMinCellX = max(0, (int)floorf(__fmul_rn(__fsub_rn(fx, r), F.invW)));
            const int nMaxCellX = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(fx, r), F.invW)));
            const int nMinCellY = max(0, (int)floorf(__fmul_rn(__fsub_rn(fy, r), F.invH)));
            const int nMaxCellY = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(fy, r), F.invH)));
            if (nMinCellX < GRID_COLS && nMaxCellX >= 0 && nMinCellY < GRID_ROWS && nMaxCellY >= 0) {
                const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
                // the slice bounds of every column of the window in one round trip (lane = column; windows wider than 32 columns loop)
                for (int ixb = nMinCellX; ixb <= nMaxCellX; ixb += 32) {
                    const int myIx = ixb + lane;
                    int myB = 0, myE = 0;
                    if (myIx <= nMaxCellX) { myB = F.cellStart[myIx * GRID_ROWS + nMinCellY]; myE = F.cellStart[myIx * GRID_ROWS + nMaxCellY + 1]; }
                    const int ncol = min(32, nMaxCellX - ixb + 1);
                    for (int cidx = 0; cidx < ncol; cidx++) {
                        const int b = __shfl_sync(0xffffffffu, myB, cidx), e = __shfl_sync(0xffffffffu, myE, cidx);
                        for (int kb = b; kb < e; kb += 32) {
                            const int k = kb + lane;
                            bool pass = k < e;
                            int idx = 0, octave = 0;
                            if (pass) {
                                const int4 it = F.cellKp[k];
                                const float kx = __int_as_float(it.x), ky = __int_as_float(it.y);
                                octave = it.z; idx = it.w;
                                if (bCheckLevels) {
                                    if (octave < minLevel) pass = false;
                                    if (maxLevel >= 0 && octave > maxLevel) pass = false;
                                }
                                const float dx = __fsub_rn(kx, x), dy = __fsub_rn(ky, y);
                                if (!(fabsf(dx) < r && fabsf(dy) < r)) pass = false;
                                if (pass && !distSem && !keypoint_passes(J, F, q, W, mode, idx, kx, ky, octave)) pass = false;
                            }
                            append(pass, idx, octave);
                        }
                    }
                }
            }
        }
    }
    if (lane == 0) ccount[q] = n;
}

__global__ void __launch_bounds__(WM_THREADS) window_match_kernel(const WinJob* __restrict__ jobs, int smemInts)
{
    // the per-call state (2 * kpCap + 6 * nq ints) lives in shared memory when it fits: the rounds are chains of
    // dependent reads of it, and with one CTA per call there is nothing else on the SM to cover L2 latency
    extern __shared__ int wmState[];
    __shared__ int sHist[HISTO_LENGTH];
    __shared__ int sKeep[3];
    __shared__ int sCount, sRemoved;
    const WinJob J = jobs[blockIdx.x];
    const FrameDev F = *J.frame;
    const int nkp = frame_count(F);
    const int nq = J.nq;
    const int tid = threadIdx.x;
    const int mode = J.mode;
    const bool distSem = is_dist_sem(mode);
    const bool independent = (mode == WM_BEST) && !(J.flags & WF_BLOCK);    // Fuse / SearchBySim3: no loop-carried state

    int* state = (2 * (size_t)J.kpCap + 6 * (size_t)nq <= (size_t)smemInts) ? wmState : J.scratch;
    int* owner = state;                     // [kpCap] BLOCK: min blocking query; DIST: head of claimant list
    int* lastOwner = owner + J.kpCap;       // [kpCap] max accepted query per keypoint
    int* choice = lastOwner + J.kpCap;      // [nq]    published: accepted keypoint or -1
    int* cdist = choice + nq;               // [nq]    published: its distance
    int* newChoice = cdist + nq;            // [nq]    staged during a sweep
    int* newCdist = newChoice + nq;         // [nq]
    int* nextq = newCdist + nq;             // [nq]    DIST: claimant list link
    int* qbin = nextq + nq;                 // [nq]    rotation bin of an accepted query
    const int* ccount = J.scratch + 2 * (size_t)J.kpCap + 6 * (size_t)nq;   // [nq] phase-1 candidate counts (global: written by window_cands_kernel)
    const int2* clistAll = reinterpret_cast<const int2*>(J.scratch + win_clist_offset(J.kpCap, nq));

    for (int i = tid; i < nkp; i += WM_THREADS) { owner[i] = distSem ? -1 : 0x7fffffff; lastOwner[i] = -1; }
    for (int i = tid; i < nq; i += WM_THREADS) { choice[i] = -1; cdist[i] = 0; }
    __syncthreads();

    for (int round = 0; round <= nq; round++) {
        int changed = 0;
        for (int q = tid; q < nq; q += WM_THREADS) {
            int selIdx = -1, selDist = 0;
            const int cnt = ccount[q];
            if (cnt > 0) {
                const int big = distSem ? 0x7fffffff : 256;
                int best = big, second = big, bestIdx = -1, bestLevel = -1, secondLevel = -1;
                auto consider = [&](int idx, int d, int level) {
                    if (!distSem) {
                        if (owner[idx] < q) return;
                    } else {
                        // vMatchedDistance[idx] <= d as left by the accepted queries < q
                        for (int c = owner[idx]; c >= 0; c = nextq[c])
                            if (c < q && cdist[c] <= d) return;
                    }
                    if (d < best) { second = best; best = d; secondLevel = bestLevel; bestLevel = level; bestIdx = idx; }
                    else if (d < second) { secondLevel = level; second = d; }
                };
                if (cnt <= WM_LISTCAP) {
                    const int2* cl = clistAll + (size_t)q * WM_LISTCAP;
                    for (int k = 0; k < cnt; k++) {
                        const int2 e = cl[k];
                        consider(e.x, e.y & 0xffff, e.y >> 16);
                    }
                } else {
                    Window W;               // overflow: enumerate again (same filters as phase 1)
                    query_window(J, q, W);
                    static_candidates(J, F, q, W, consider);
                }
                bool acc;
                if (mode == WM_PROJ || mode == WM_PROJ_BIRD)
                    acc = best <= TH_HIGH && !(bestLevel == secondLevel && (float)best > __fmul_rn(J.nnratio, (float)second));
                else if (mode == WM_PROJ_FRAME)
                    acc = best <= TH_HIGH;
                else if (mode == WM_BEST)
                    acc = best <= J.accTh;
                else if (mode == WM_BIRD)
                    acc = best <= TH_LOW && (float)best < __fmul_rn((float)second, J.nnratio);
                else if (mode == WM_BOW_KF_F)
                    acc = best <= TH_LOW && (float)best < __fmul_rn(J.nnratio, (float)second);     // :228-230
                else if (mode == WM_BOW_KF_KF)
                    acc = best < TH_LOW && (float)best < __fmul_rn(J.nnratio, (float)second);      // :593-595
                else
                    acc = best <= TH_HIGH && (bestLevel != secondLevel || (float)best < __fmul_rn((float)second, J.nnratio));
                if (acc) { selIdx = bestIdx; selDist = best; }
            }
            if (selIdx != choice[q] || selDist != cdist[q]) changed = 1;
            newChoice[q] = selIdx;      // published after the barrier: other threads still read choice/cdist
            newCdist[q] = selDist;
        }
        const int any = __syncthreads_or(changed);
        if (!any) break;
        for (int q = tid; q < nq; q += WM_THREADS) { choice[q] = newChoice[q]; cdist[q] = newCdist[q]; }
        if (independent) { __syncthreads(); break; }
        for (int i = tid; i < nkp; i += WM_THREADS) owner[i] = distSem ? -1 : 0x7fffffff;
        __syncthreads();
        for (int q = tid; q < nq; q += WM_THREADS) {
            const int cidx = choice[q];
            if (cidx < 0) continue;
            if (distSem) nextq[q] = atomicExch(&owner[cidx], q);
            else if (J.q_obs_pos == nullptr || J.q_obs_pos[q]) atomicMin(&owner[cidx], q);
        }
        __syncthreads();
    }

    // ---- outputs ----
    if (tid < HISTO_LENGTH) sHist[tid] = 0;
    if (tid == 0) { sCount = 0; sRemoved = 0; }
    __syncthreads();
    int nAcc = 0;
    for (int q = tid; q < nq; q += WM_THREADS) {
        const int cidx = choice[q];
        qbin[q] = -1;
        if (cidx < 0) continue;
        nAcc++;
        atomicMax(&lastOwner[cidx], q);
        if (J.checkOri) {
            const int b = rot_bin(J.q_angle[q], F.kps[cidx].angle);
            qbin[q] = b;
            atomicAdd(&sHist[b], 1);
        }
    }
    if (nAcc) atomicAdd(&sCount, nAcc);
    __syncthreads();
    if (tid == 0) {
        int i1 = -1, i2 = -1, i3 = -1;
        if (J.checkOri) three_maxima(sHist, i1, i2, i3);
        sKeep[0] = i1; sKeep[1] = i2; sKeep[2] = i3;
    }
    __syncthreads();
    const int k1 = sKeep[0], k2 = sKeep[1], k3 = sKeep[2];
    if (J.out_best_idx)
        for (int q = tid; q < nq; q += WM_THREADS) { J.out_best_idx[q] = choice[q]; if (J.out_best_dist) J.out_best_dist[q] = choice[q] >= 0 ? cdist[q] : 256; }
    if (mode == WM_BIRD) {
        // vnMatches12 (:1725-1733): the last accepted claimant owns the keypoint; rotation filter only
        // un-matches queries that still hold their match (:1758-1768)
        int removed = 0, owned = 0;
        for (int q = tid; q < nq; q += WM_THREADS) {
            const int cidx = choice[q];
            int m = -1;
            if (cidx >= 0 && lastOwner[cidx] == q) {
                owned++;
                m = cidx;
                if (J.checkOri) { const int b = qbin[q]; if (b != k1 && b != k2 && b != k3) { m = -1; removed++; } }
            }
            J.out_per_query[q] = m;
        }
        __syncthreads();              // sCount (accepted) has been read by nobody yet: reuse as "owned"
        if (tid == 0) sCount = 0;
        __syncthreads();
        if (owned) atomicAdd(&sCount, owned);
        if (removed) atomicAdd(&sRemoved, removed);
        __syncthreads();
        if (tid == 0) *J.out_nmatches = sCount - sRemoved;
        return;
    }
    if (mode == WM_BOW_KF_KF) {
        // vpMatches12[idx1] (:597, :643-647): per query, cleared when its rotation bin is discarded
        int removed = 0;
        for (int q = tid; q < nq; q += WM_THREADS) {
            int m = choice[q];
            if (m >= 0 && J.checkOri) { const int b = qbin[q]; if (b != k1 && b != k2 && b != k3) { m = -1; removed++; } }
            J.out_per_query[q] = m;
        }
        if (removed) atomicAdd(&sRemoved, removed);
        __syncthreads();
        if (tid == 0) *J.out_nmatches = sCount - sRemoved;
        return;
    }
    // BLOCK modes and BIRD_KF: per keypoint the last accepted query, unless an accepted query on it sits in a
    // discarded rotation bin (:1455-1463, :2098-2108); nmatches = accepted - entries in discarded bins
    for (int i = tid; i < nkp; i += WM_THREADS) J.out_per_kp[i] = lastOwner[i];
    __syncthreads();
    if (J.checkOri) {
        int removed = 0;
        for (int q = tid; q < nq; q += WM_THREADS) {
            const int cidx = choice[q];
            if (cidx < 0) continue;
            const int b = qbin[q];
            if (b != k1 && b != k2 && b != k3) { J.out_per_kp[cidx] = -1; removed++; }
        }
        if (removed) atomicAdd(&sRemoved, removed);
    }
    __syncthreads();
    if (tid == 0) *J.out_nmatches = sCount - sRemoved;
}

void launch_window_match(Ctx& c, const WinJob* d_jobs, int njobs, int maxNq, int maxKpCap)
{
    // shared-memory budget of window_match_kernel's per-call state; larger calls keep it in their global scratch
    constexpr size_t WM_SMEM_MAX = 200 * 1024;
    size_t smem = sizeof(int) * (2 * (size_t)std::max(maxKpCap, 0) + 6 * (size_t)std::max(maxNq, 0));
    if (smem > WM_SMEM_MAX) smem = 0;
    if (smem > 48 * 1024 && smem > ensure_max_dynamic_smem(c.device, (const void*)window_match_kernel, SMEM_WINDOW_MATCH)) smem = 0;
    if (maxNq > 0) {
        if (c.warpCands && (long long)maxNq * njobs <= 32768) {     // a few calls: one warp per query (see window_cands_warp_kernel)
            window_cands_warp_kernel<<<dim3((maxNq + WCW_WARPS - 1) / WCW_WARPS, njobs), WCW_WARPS * 32, 0, c.stream>>>(d_jobs);
        } else {
            dim3 grid((maxNq + WC_THREADS - 1) / WC_THREADS, njobs);
            window_cands_kernel<<<grid, WC_THREADS, 0, c.stream>>>(d_jobs);
        }
        c.launches++;
    }
    window_match_kernel<<<njobs, WM_THREADS, smem, c.stream>>>(d_jobs, (int)(smem / sizeof(int)));
    c.launches++;
}

// ---------------------------------------------------------------------------------------------------
// SearchForTriangulation (reference src/ORBmatcher.cc:657-823): no loop-carried state (vbMatched2 is never
// set), so every keypoint of KF1 is independent.  One warp per KF1 keypoint of a shared vocabulary node;
// lanes stride over the node's KF2 keypoints; the running "dist <= bestDist" rule makes the winner the
// minimum distance among candidates passing the epipolar tests, last in scan order on ties.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) triangulation_kernel(TriJob J)
{
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= J.nItems) return;
    const int idx1 = J.item_idx1[w];
    const int b2 = J.item_b2[w], e2 = J.item_e2[w];
    J.match12[idx1] = -1;
    if (J.has_mp1[idx1]) return;
    const bool st1 = J.uR1 ? J.uR1[idx1] >= 0 : false;
    if (J.onlyStereo && !st1) return;
    const orbb200_kp_t kp1 = J.kps1[idx1];
    const uint4* d1 = reinterpret_cast<const uint4*>(J.desc1) + 2 * (size_t)idx1;
    const uint4 qa = d1[0], qb = d1[1];
    // epipolar line of kp1 in image 2 (CheckDistEpipolarLine, :140-157)
    const float* Fm = J.F12;
    const float a = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, Fm[0]), __fmul_rn(kp1.y, Fm[3])), Fm[6]);
    const float b = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, Fm[1]), __fmul_rn(kp1.y, Fm[4])), Fm[7]);
    const float cc = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, Fm[2]), __fmul_rn(kp1.y, Fm[5])), Fm[8]);
    const float den = __fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b));
    int best = 0x7fffffff, bestPos = -1;   // minimise (dist, -pos)
    for (int p = b2 + lane; p < e2; p += 32) {
        const int idx2 = J.fv2_idx[p];
        if (J.has_mp2[idx2]) continue;
        const bool st2 = J.uR2 ? J.uR2[idx2] >= 0 : false;
        if (J.onlyStereo && !st2) continue;
        const uint4* d2 = reinterpret_cast<const uint4*>(J.desc2) + 2 * (size_t)idx2;
        const int dist = hamming256(qa, qb, d2[0], d2[1]);
        if (dist > TH_LOW) continue;
        const orbb200_kp_t kp2 = J.kps2[idx2];
        if ((unsigned)kp2.octave >= (unsigned)J.nLevels2) continue;       // outside the keyframe's pyramid tables
        if (!st1 && !st2) {
            const float dx = __fsub_rn(J.ex, kp2.x), dy = __fsub_rn(J.ey, kp2.y);
            if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.f, J.scaleFactors2[kp2.octave])) continue;
        }
        if (den == 0) continue;
        const float num = __fadd_rn(__fadd_rn(__fmul_rn(a, kp2.x), __fmul_rn(b, kp2.y)), cc);
        const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
        if (!((double)dsqr < 3.84 * (double)J.levelSigma2_2[kp2.octave])) continue;
        if (dist < best || (dist == best && p > bestPos)) { best = dist; bestPos = p; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int ob = __shfl_xor_sync(0xffffffffu, best, o), op = __shfl_xor_sync(0xffffffffu, bestPos, o);
        if (ob < best || (ob == best && op > bestPos)) { best = ob; bestPos = op; }
    }
    if (lane == 0 && bestPos >= 0) J.match12[idx1] = J.fv2_idx[bestPos];
}

// rotation-histogram filter + pair compaction (:791-822); single CTA
__global__ void __launch_bounds__(1024) triangulation_finish_kernel(TriJob J)
{
    __shared__ int sHist[HISTO_LENGTH];
    __shared__ int sKeep[3];
    __shared__ int warpTot[32];
    const int tid = threadIdx.x;
    if (tid < HISTO_LENGTH) sHist[tid] = 0;
    __syncthreads();
    if (J.checkOri) {
        for (int i = tid; i < J.n1; i += 1024) {
            const int m = J.match12[i];
            if (m >= 0) atomicAdd(&sHist[rot_bin(J.kps1[i].angle, J.kps2[m].angle)], 1);
        }
        __syncthreads();
        if (tid == 0) { int a, b, c; three_maxima(sHist, a, b, c); sKeep[0] = a; sKeep[1] = b; sKeep[2] = c; }
        __syncthreads();
        for (int i = tid; i < J.n1; i += 1024) {
            const int m = J.match12[i];
            if (m < 0) continue;
            const int bin = rot_bin(J.kps1[i].angle, J.kps2[m].angle);
            if (bin != sKeep[0] && bin != sKeep[1] && bin != sKeep[2]) J.match12[i] = -1;
        }
        __syncthreads();
    }
    // ordered compaction over idx1
    int carry = 0;
    const int lane = tid & 31, wid = tid >> 5;
    for (int base = 0; base < J.n1; base += 1024) {
        const int i = base + tid;
        const int f = (i < J.n1 && J.match12[i] >= 0) ? 1 : 0;
        int x = f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) warpTot[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int t = warpTot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
            warpTot[lane] = t;
        }
        __syncthreads();
        const int ex = carry + (wid ? warpTot[wid - 1] : 0) + x - f;
        if (f) { J.pairs[2 * ex] = i; J.pairs[2 * ex + 1] = J.match12[i]; }
        carry += warpTot[31];
        __syncthreads();
    }
    if (tid == 0) *J.npairs = carry;
}

void launch_triangulation(Ctx& c, const TriJob& J)
{
    if (J.n1 > 0) cudaMemsetAsync(J.match12, 0xff, sizeof(int32_t) * J.n1, c.stream);
    if (J.nItems > 0) {
        triangulation_kernel<<<(J.nItems * 32 + 255) / 256, 256, 0, c.stream>>>(J);
        c.launches++;
    }
    triangulation_finish_kernel<<<1, 1024, 0, c.stream>>>(J);
    c.launches++;
}

// ---------------------------------------------------------------------------------------------------
// Frame::isInFrustum (reference src/Frame.cc:436-492) + MapPoint::PredictScale (src/MapPoint.cc:402-417),
// one thread per map point.  cv::Mat arithmetic as OpenCV evaluates it for these shapes: mRcw*P+mtcw = float
// products summed left to right, then the translation added (the double add + float rounding gemm performs equals
// one float add); cv::norm and Mat::dot accumulate in double (the products of two floats are exact in double, so
// DFMA contraction cannot change them).  logf is evaluated in double and rounded (host libm's logf is correctly
// rounded for all but ~2^-10 of inputs, and the level only changes when log(ratio)/log(1.2) sits on an integer).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) frustum_kernel(FrustumJob J)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = false;
    float u = 0.f, v = 0.f, uR = 0.f, viewCos = 0.f;
    int nScale = 0;
    // batched form (J.poses != null): blockIdx.y = frame, its pose from device memory, its outputs at [frame][n]
    const size_t fo = J.poses ? (size_t)blockIdx.y * J.n : 0;
    const orbb200_camera_pose Fp = J.poses ? J.poses[blockIdx.y] : J.pose;
    if (i < J.n && (!J.candidate || J.candidate[i])) {
        const orbb200_camera_pose& F = Fp;
        const float Px = J.pos[3 * i], Py = J.pos[3 * i + 1], Pz = J.pos[3 * i + 2];
        float Pc[3];
#pragma unroll
        for (int r = 0; r < 3; r++) {
            float t = __fadd_rn(__fmul_rn(F.Rcw[3 * r], Px), __fmul_rn(F.Rcw[3 * r + 1], Py));
            t = __fadd_rn(t, __fmul_rn(F.Rcw[3 * r + 2], Pz));
            Pc[r] = __fadd_rn(t, F.tcw[r]);
        }
        do {
            if (Pc[2] < 0.0f) break;
            const float invz = __fdiv_rn(1.0f, Pc[2]);
            u = __fadd_rn(__fmul_rn(__fmul_rn(F.fx, Pc[0]), invz), F.cx);
            v = __fadd_rn(__fmul_rn(__fmul_rn(F.fy, Pc[1]), invz), F.cy);
            if (u < F.min_x || u > F.max_x) break;
            if (v < F.min_y || v > F.max_y) break;
            const float maxD = J.maxDist[i];
            const float maxDistance = __fmul_rn(1.2f, maxD), minDistance = __fmul_rn(0.8f, J.minDist[i]);
            const float POx = __fsub_rn(Px, F.Ow[0]), POy = __fsub_rn(Py, F.Ow[1]), POz = __fsub_rn(Pz, F.Ow[2]);
            const float dist = (float)sqrt((double)POx * POx + (double)POy * POy + (double)POz * POz);
            if (dist < minDistance || dist > maxDistance) break;
            const double dot = (double)POx * J.normal[3 * i] + (double)POy * J.normal[3 * i + 1] + (double)POz * J.normal[3 * i + 2];
            viewCos = (float)(dot / (double)dist);
            if (viewCos < J.cosLimit) break;
            const float ratio = __fdiv_rn(maxD, dist);
            const float cl = ceilf(__fdiv_rn((float)log((double)ratio), F.log_scale_factor));
            // (int) as the host's cvttss2si: NaN / out of range -> INT_MIN
            nScale = (cl > -2147483648.f && cl < 2147483648.f) ? (int)cl : INT_MIN;
            nScale = nScale < 0 ? 0 : (nScale >= F.n_levels ? F.n_levels - 1 : nScale);
            uR = __fsub_rn(u, __fmul_rn(F.mbf, invz));
            ok = true;
        } while (false);
    }
    if (i < J.n) {
        J.inView[fo + i] = ok ? 1 : 0;
        J.u[fo + i] = ok ? u : 0.f; J.v[fo + i] = ok ? v : 0.f; J.uR[fo + i] = ok ? uR : 0.f;
        J.level[fo + i] = ok ? nScale : 0; J.viewcos[fo + i] = ok ? viewCos : 0.f;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && bal) atomicAdd(J.count + (J.poses ? blockIdx.y : 0), __popc(bal));
}

void launch_frustum(Ctx& c, const FrustumJob& J)
{
    const int frames = J.poses ? J.nFrames : 1;
    cudaMemsetAsync(J.count, 0, sizeof(int32_t) * frames, c.stream);
    if (J.n > 0) {
        frustum_kernel<<<dim3((J.n + 127) / 128, frames), 128, 0, c.stream>>>(J);
        c.launches++;
    }
}

// ---------------------------------------------------------------------------------------------------
// MapPoint[Bird]::ComputeDistinctiveDescriptors selection (src/MapPoint.cc:272-301): one CTA per landmark, one warp
// per row of the distance matrix.  The median of a row is a rank selection, not a sort: the row's distances (0..256)
// are counted into a per-warp histogram in shared memory and the bin where the running count first exceeds
// k = (int)(0.5*(N-1)) is the k-th smallest.  First row with the smallest median = lexicographic min of (median, row).
// ---------------------------------------------------------------------------------------------------
constexpr int DD_WARPS = 4;
constexpr int DD_BINS = 288;      // 257 used; 9 bins per lane

__global__ void __launch_bounds__(DD_WARPS * 32) distinctive_kernel(const uint8_t* __restrict__ desc, const int32_t* __restrict__ groupPtr,
                                                                    int32_t* __restrict__ bestIdx, int32_t* __restrict__ bestMedian)
{
    __shared__ int hist[DD_WARPS][DD_BINS];
    __shared__ int sKey[DD_WARPS];
    const int g = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int b = groupPtr[g], N = groupPtr[g + 1] - b;
    if (N <= 0) {
        if (threadIdx.x == 0) { bestIdx[g] = -1; bestMedian[g] = -1; }
        return;
    }
    const uint4* D = reinterpret_cast<const uint4*>(desc) + 2 * (size_t)b;
    const int k = (int)(0.5 * (N - 1));
    int* h = hist[wid];
    for (int i = lane; i < DD_BINS; i += 32) h[i] = 0;
    __syncwarp();
    int best = 0x7fffffff;                                          // median << 20 | row  (N < 2^20 checked by the host)
    for (int i = wid; i < N; i += DD_WARPS) {
        const uint4 a0 = __ldg(D + 2 * i), a1 = __ldg(D + 2 * i + 1);
        for (int j = lane; j < N; j += 32) {
            const int d = j == i ? 0 : hamming256(a0, a1, __ldg(D + 2 * j), __ldg(D + 2 * j + 1));
            atomicAdd(&h[d], 1);
        }
        __syncwarp();
        int c[9], sum = 0;
#pragma unroll
        for (int t = 0; t < 9; t++) { c[t] = h[9 * lane + t]; sum += c[t]; }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int run = incl - sum, med = -1;                             // elements in the bins before this lane's
        const bool mine = run <= k && k < incl;                     // exactly one lane
        if (mine) {
#pragma unroll
            for (int t = 0; t < 9; t++) {
                if (med < 0 && k < run + c[t]) med = 9 * lane + t;
                run += c[t];
            }
        }
        const unsigned who = __ballot_sync(0xffffffffu, mine);
        med = __shfl_sync(0xffffffffu, med, __ffs(who) - 1);
#pragma unroll
        for (int t = 0; t < 9; t++) h[9 * lane + t] = 0;
        __syncwarp();
        best = min(best, (med << 20) | i);
    }
    if (lane == 0) sKey[wid] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        int r = sKey[0];
#pragma unroll
        for (int w = 1; w < DD_WARPS; w++) r = min(r, sKey[w]);
        bestIdx[g] = r & 0xfffff; bestMedian[g] = r >> 20;
    }
}

void launch_distinctive(Ctx& c, const uint8_t* d_desc, const int32_t* d_groupPtr, int nGroups, int32_t* d_bestIdx, int32_t* d_bestMedian)
{
    if (nGroups <= 0) return;
    distinctive_kernel<<<nGroups, DD_WARPS * 32, 0, c.stream>>>(d_desc, d_groupPtr, d_bestIdx, d_bestMedian);
    c.launches++;
}

}  // namespace orbb200
