// Frame::ComputeStereoMatches (reference src/Frame.cc:662-836) on the device: row-band descriptor search in the
// right image, 11x11 SAD sliding-window refinement on the un-blurred pyramid level, parabola sub-pixel fit,
// and the median-based outlier cut.  Works on the results of the last extraction (keypoints, descriptors and
// pyramids of both images are already resident), which also removes the mvImagePyramid download the
// reference's stereo path needs.
#include "ctx.cuh"

namespace orbb200 {

__device__ __forceinline__ int hamming256s(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1)
{
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

constexpr int SM_WARPS = 8;

// vRowIndices (:673-686): for every image row the right keypoints whose band [floor(y-r), ceil(y+r)], r = 2*scale[octave],
// contains it.  CSR per frame, built by one CTA; the lists are unordered (the match reduction is order-free).
__global__ void __launch_bounds__(1024) stereo_rows_kernel(Geom g, const orbb200_kp_t* __restrict__ kps, const int32_t* __restrict__ counts,
                                                           int right0, int strideImgs, int32_t* __restrict__ rowStart, int32_t* __restrict__ rowItems,
                                                           int rowCap, int itemCap)
{
    extern __shared__ int sRow[];            // [nRows + 1]
    __shared__ int warpTot[32];
    const int frame = blockIdx.x, tid = threadIdx.x;
    const int imgR = right0 + frame * strideImgs;
    const int nR = min(counts[imgR], g.kpPerImg);
    const int nRows = g.h;
    const orbb200_kp_t* K = kps + (size_t)imgR * g.kpPerImg;
    int32_t* RS = rowStart + (size_t)frame * rowCap;
    int32_t* RI = rowItems + (size_t)frame * itemCap;
    for (int i = tid; i <= nRows; i += 1024) sRow[i] = 0;
    __syncthreads();
    for (int i = tid; i < nR; i += 1024) {
        const float y = K[i].y, r = __fmul_rn(2.0f, g.lv[K[i].octave].scale);
        const int lo = max((int)floorf(__fsub_rn(y, r)), 0), hi = min((int)ceilf(__fadd_rn(y, r)), nRows - 1);
        for (int yy = lo; yy <= hi; yy++) atomicAdd(&sRow[yy], 1);
    }
    __syncthreads();
    // exclusive scan over the rows (chunks of 1024)
    int carry = 0;
    const int lane = tid & 31, wid = tid >> 5;
    for (int base = 0; base < nRows; base += 1024) {
        const int i = base + tid;
        const int v = i < nRows ? sRow[i] : 0;
        int x = v;
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
        const int ex = carry + (wid ? warpTot[wid - 1] : 0) + x - v;
        if (i < nRows) { RS[i] = ex; sRow[i] = ex; }
        carry += warpTot[31];
        __syncthreads();
    }
    if (tid == 0) RS[nRows] = carry;
    __syncthreads();
    for (int i = tid; i < nR; i += 1024) {
        const float y = K[i].y, r = __fmul_rn(2.0f, g.lv[K[i].octave].scale);
        const int lo = max((int)floorf(__fsub_rn(y, r)), 0), hi = min((int)ceilf(__fadd_rn(y, r)), nRows - 1);
        for (int yy = lo; yy <= hi; yy++) {
            const int p = atomicAdd(&sRow[yy], 1);
            if (p < itemCap) RI[p] = i;
        }
    }
}

// One warp per left keypoint.  The reference walks vRowIndices[vL] in ascending right index and keeps the first
// minimum; the lexicographic (distance, right index) minimum over the unordered row list is the same element.
__global__ void __launch_bounds__(SM_WARPS * 32) stereo_match_kernel(Geom g, const uint8_t* __restrict__ pyr,
                                                                    const orbb200_kp_t* __restrict__ kps, const uint8_t* __restrict__ desc,
                                                                    const int32_t* __restrict__ counts, int left0, int right0, int strideImgs,
                                                                    const float* __restrict__ invScale, float mb, float mbf,
                                                                    const int32_t* __restrict__ rowStart, const int32_t* __restrict__ rowItems,
                                                                    int rowCap, int itemCap,
                                                                    float* __restrict__ uRight, float* __restrict__ depth, int32_t* __restrict__ sad)
{
    const int frame = blockIdx.y;
    const int imgL = left0 + frame * strideImgs, imgR = right0 + frame * strideImgs;
    const int lane = threadIdx.x & 31;
    const int iL = blockIdx.x * SM_WARPS + (threadIdx.x >> 5);
    const int nL = min(counts[imgL], g.kpPerImg);
    if (iL >= nL) return;
    const size_t oL = (size_t)imgL * g.kpPerImg, oR = (size_t)imgR * g.kpPerImg;
    float outU = -1.0f, outD = -1.0f;
    int outS = -1;

    const orbb200_kp_t kpL = kps[oL + iL];
    const int levelL = kpL.octave;
    const float vL = kpL.y, uL = kpL.x;
    const int row = (int)vL;
    const float maxD = __fdiv_rn(mbf, mb);               // :687-689: minZ = mb, minD = 0, maxD = mbf/minZ
    const float minU = __fsub_rn(uL, maxD), maxU = uL;   // uL - minD
    int best = TH_HIGH, bestR = 0x7fffffff;
    if (!(maxU < 0)) {
        const uint4* dl = reinterpret_cast<const uint4*>(desc) + 2 * (oL + iL);
        const uint4 qa = dl[0], qb = dl[1];
        if (row >= 0 && row < g.h) {
            const int32_t* RS = rowStart + (size_t)frame * rowCap;
            const int32_t* RI = rowItems + (size_t)frame * itemCap;
            const int pb = RS[row], pe = min(RS[row + 1], itemCap);
            for (int p = pb + lane; p < pe; p += 32) {
                const int iR = RI[p];
                const orbb200_kp_t* kr = kps + oR + iR;
                const int octR = kr->octave;
                if (octR < levelL - 1 || octR > levelL + 1) continue;
                const float uR = kr->x;
                if (!(uR >= minU && uR <= maxU)) continue;
                const uint4* dr = reinterpret_cast<const uint4*>(desc) + 2 * (oR + iR);
                const int d = hamming256s(qa, qb, dr[0], dr[1]);
                if (d < best || (d == best && iR < bestR)) { best = d; bestR = iR; }   // first minimum in right-index order
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int ob = __shfl_xor_sync(0xffffffffu, best, o), oi = __shfl_xor_sync(0xffffffffu, bestR, o);
        if (ob < best || (ob == best && oi < bestR)) { best = ob; bestR = oi; }
    }
    constexpr int thOrbDist = (TH_HIGH + TH_LOW) / 2;
    if (best < thOrbDist) {
        // ---- sub-pixel match by correlation (:739-815), integer SAD: every term is a small integer ----
        const float uR0 = kps[oR + bestR].x;
        const float sf = invScale[levelL];
        const float suL = roundf(__fmul_rn(kpL.x, sf)), svL = roundf(__fmul_rn(kpL.y, sf)), suR0 = roundf(__fmul_rn(uR0, sf));
        const LevelGeom L = g.lv[levelL];
        constexpr int w = 5, LL = 5;
        const float iniu = suR0 + LL - w, endu = suR0 + LL + w + 1;
        if (!(iniu < 0 || endu >= (float)L.w)) {
            const int y0 = (int)(svL - w), xL0 = (int)(suL - w), xR0 = (int)(suR0 - w);   // xR0: incR = 0
            const uint8_t* PL = pyr + (size_t)imgL * g.pyrBytes + L.off;
            const uint8_t* PR = pyr + (size_t)imgR * g.pyrBytes + L.off;
            const int cL = PL[(size_t)(y0 + w) * L.pitch + xL0 + w];
            int cR[2 * LL + 1];
#pragma unroll
            for (int k = 0; k < 2 * LL + 1; k++) cR[k] = PR[(size_t)(y0 + w) * L.pitch + xR0 + (k - LL) + w];
            int acc[2 * LL + 1];
#pragma unroll
            for (int k = 0; k < 2 * LL + 1; k++) acc[k] = 0;
            for (int p = lane; p < 121; p += 32) {
                const int dy = p / 11, dx = p - dy * 11;
                const int a = (int)PL[(size_t)(y0 + dy) * L.pitch + xL0 + dx] - cL;
                const uint8_t* rr = PR + (size_t)(y0 + dy) * L.pitch + xR0 + dx;
#pragma unroll
                for (int k = 0; k < 2 * LL + 1; k++) acc[k] += abs(a - ((int)rr[k - LL] - cR[k]));
            }
#pragma unroll
            for (int k = 0; k < 2 * LL + 1; k++) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
            }
            int bestS = 0x7fffffff, bestInc = 0;
#pragma unroll
            for (int k = 0; k < 2 * LL + 1; k++)
                if (acc[k] < bestS) { bestS = acc[k]; bestInc = k - LL; }
            if (bestInc != -LL && bestInc != LL) {
                float d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
                for (int k = 1; k < 2 * LL; k++)
                    if (k - LL == bestInc) { d1 = (float)acc[k - 1]; d2 = (float)acc[k]; d3 = (float)acc[k + 1]; }
                // deltaR = (dist1-dist3)/(2.0f*(dist1+dist3-2.0f*dist2)) (:798)
                const float den = __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2)));
                const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), den);
                if (!(deltaR < -1 || deltaR > 1)) {
                    float bestuR = __fmul_rn(L.scale, __fadd_rn(__fadd_rn(suR0, (float)bestInc), deltaR));
                    float disparity = __fsub_rn(uL, bestuR);
                    if (disparity >= 0.f && disparity < maxD) {
                        if (disparity <= 0) { disparity = 0.01f; bestuR = (float)((double)uL - 0.01); }
                        outD = __fdiv_rn(mbf, disparity);
                        outU = bestuR;
                        outS = bestS;
                    }
                }
            }
        }
    }
    if (lane == 0) { uRight[oL + iL] = outU; depth[oL + iL] = outD; sad[oL + iL] = outS; }
}

// Outlier cut (:817-835): thDist = 1.5*1.4*median of the SAD distances (median = element size/2 of the sorted
// list); every match with dist >= thDist is dropped.  One CTA per frame; the order statistic comes from a
// two-level histogram (SAD <= 121*510 < 2^16).
__global__ void __launch_bounds__(256) stereo_median_kernel(int kpPerImg, const int32_t* __restrict__ counts, int left0, int strideImgs,
                                                            float* __restrict__ uRight, float* __restrict__ depth,
                                                            const int32_t* __restrict__ sad, int32_t* __restrict__ nKept)
{
    __shared__ int hist[256];
    __shared__ int sTotal, sBin, sBefore, sMedian;
    const int frame = blockIdx.x, tid = threadIdx.x;
    const int imgL = left0 + frame * strideImgs;
    const int n = min(counts[imgL], kpPerImg);
    const size_t o = (size_t)imgL * kpPerImg;
    hist[tid] = 0;
    if (tid == 0) sTotal = 0;
    __syncthreads();
    int mine = 0;
    for (int i = tid; i < n; i += 256) {
        const int s = sad[o + i];
        if (s >= 0) { atomicAdd(&hist[min(s >> 8, 255)], 1); mine++; }
    }
    if (mine) atomicAdd(&sTotal, mine);
    __syncthreads();
    const int total = sTotal;
    if (total == 0) { if (tid == 0) nKept[frame] = 0; return; }
    const int k = total / 2;
    if (tid == 0) {
        int acc = 0, b = 0;
        for (; b < 256; b++) { if (acc + hist[b] > k) break; acc += hist[b]; }
        sBin = b; sBefore = acc;
    }
    __syncthreads();
    const int bin = sBin, before = sBefore;
    __syncthreads();
    hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        const int s = sad[o + i];
        if (s >= 0 && min(s >> 8, 255) == bin) atomicAdd(&hist[s & 255], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int acc = before, b = 0;
        for (; b < 256; b++) { if (acc + hist[b] > k) break; acc += hist[b]; }
        sMedian = (bin << 8) | b;
    }
    __syncthreads();
    const float thDist = __fmul_rn(1.5f * 1.4f, (float)sMedian);
    int kept = 0;
    for (int i = tid; i < n; i += 256) {
        const int s = sad[o + i];
        if (s < 0) continue;
        if ((float)s < thDist) kept++;
        else { uRight[o + i] = -1.0f; depth[o + i] = -1.0f; }
    }
    __syncthreads();
    if (tid == 0) sTotal = 0;
    __syncthreads();
    if (kept) atomicAdd(&sTotal, kept);
    __syncthreads();
    if (tid == 0) nKept[frame] = sTotal;
}

void launch_stereo(Ctx& c, int n_frames, int left0, int right0, int strideImgs, float mb, float mbf, const float* d_invScale, int32_t* d_nKept)
{
    const Geom& g = c.cur->g;
    const int rowCap = g.h + 1, itemCap = c.stereoItemCap;
    const size_t smem = sizeof(int) * (size_t)(g.h + 1);
    stereo_rows_kernel<<<n_frames, 1024, smem, c.stream>>>(g, c.d_kps, c.d_counts, right0, strideImgs, c.d_rowStart, c.d_rowItems, rowCap, itemCap);
    dim3 grid((g.kpPerImg + SM_WARPS - 1) / SM_WARPS, n_frames);
    stereo_match_kernel<<<grid, SM_WARPS * 32, 0, c.stream>>>(g, c.d_pyr, c.d_kps, c.d_desc, c.d_counts, left0, right0, strideImgs, d_invScale, mb, mbf,
                                                              c.d_rowStart, c.d_rowItems, rowCap, itemCap, c.d_uRight, c.d_depth, c.d_sad);
    stereo_median_kernel<<<n_frames, 256, 0, c.stream>>>(g.kpPerImg, c.d_counts, left0, strideImgs, c.d_uRight, c.d_depth, c.d_sad, d_nKept);
    c.launches += 3;
}

}  // namespace orbb200
