// C ABI of liborbb200.so (include/orbb200.h): context, geometry, host/device entry points.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "match.cuh"

using namespace orbb200;


struct orbb200_frame {
    Ctx* ctx = nullptr;
    FrameDev h{};
    FrameDev* d_self = nullptr;
    void* owned[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int cap = 0;
};

// Local map snapshot on the device + the per-point projection results of the last isInFrustum pass.
struct orbb200_map {
    Ctx* ctx = nullptr;
    int n = 0;
    float* d_pos = nullptr; float* d_normal = nullptr; float* d_maxDist = nullptr; float* d_minDist = nullptr;
    uint8_t* d_desc = nullptr;
    uint8_t* d_candidate = nullptr; uint8_t* d_inView = nullptr;
    float* d_u = nullptr; float* d_v = nullptr; float* d_uR = nullptr; float* d_viewcos = nullptr;
    int32_t* d_level = nullptr; int32_t* d_count = nullptr;
};

namespace orbb200 {

// last orbb200_create() failure of THIS host thread (Tracking, LocalMapping and LoopClosing each create their own contexts)
static thread_local std::string g_create_err;
static std::mutex g_create_mu;

static inline int cvRoundF(float v) { return (int)lrintf(v); }
static inline int cvFloorF(float v) { int i = (int)v; return i - (i > v); }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Geometry of every level for an image shape: level sizes (ComputePyramid, reference
// src/ORBextractor.cc:1111-1112), FAST grid (:776-787), octree roots (:543-546), pool offsets.
bool build_geom(const Ctx& c, int w, int h, Geom& g, std::string& err)
{
    memset(&g, 0, sizeof(g));
    g.nlevels = c.nlevels; g.w = w; g.h = h; g.iniTh = c.iniTh; g.minTh = c.minTh;
    size_t off = 0, candOff = 0;
    int kpOff = 0, cellBase = 0, xt = 0, yt = 0;
    for (int l = 0; l < c.nlevels; l++) {
        LevelGeom& L = g.lv[l];
        L.w = std::max(cvRoundF((float)w * c.invScale[l]), 0);
        L.h = std::max(cvRoundF((float)h * c.invScale[l]), 0);
        // bordered storage (cf. the reference's copyMakeBorder, src/ORBextractor.cc:1122-1128): PYR_MARGIN_Y
        // rows above/below and PYR_MARGIN_X bytes left of every row hold the reflect-101 border, so that the
        // stencil kernels never branch on image edges; L.off addresses pixel (0,0)
        L.pitch = (int)align_up((size_t)PYR_MARGIN_X + std::max(L.w, 1) + 16, 128);
        L.off = (unsigned)(off + (size_t)PYR_MARGIN_Y * L.pitch + PYR_MARGIN_X);
        off += align_up((size_t)L.pitch * (std::max(L.h, 1) + 2 * PYR_MARGIN_Y), 256);
        L.scale = c.scale[l];
        L.quota = c.quota[l];
        L.patchSize = (int)(PATCH_SIZE * c.scale[l]);
        L.maxBX = L.w - EDGE_THRESHOLD + 3;
        L.maxBY = L.h - EDGE_THRESHOLD + 3;
        const float width = (float)(L.maxBX - FAST_BORDER), height = (float)(L.maxBY - FAST_BORDER);
        L.nCols = width > 0 ? (int)(width / 30.f) : 0;
        L.nRows = height > 0 ? (int)(height / 30.f) : 0;
        L.cellBase = cellBase;
        long candCap = 0;
        if (L.nCols > 0 && L.nRows > 0) {
            L.wCell = (int)ceilf(width / L.nCols);
            L.hCell = (int)ceilf(height / L.nRows);
            if (L.wCell + 6 > 65 || L.hCell + 6 > 65) { err = "FAST cell larger than 59 px: unsupported geometry"; return false; }
            for (int i = 0; i < L.nRows; i++) {
                const float iniY = (float)(FAST_BORDER + i * L.hCell);
                float maxY = iniY + L.hCell + 6;
                if (iniY >= L.maxBY - 3) continue;
                if (maxY > L.maxBY) maxY = (float)L.maxBY;
                for (int j = 0; j < L.nCols; j++) {
                    const float iniX = (float)(FAST_BORDER + j * L.wCell);
                    float maxX = iniX + L.wCell + 6;
                    if (iniX >= L.maxBX - 6) continue;
                    if (maxX > L.maxBX) maxX = (float)L.maxBX;
                    const int wi = (int)maxX - (int)iniX - 6, hi = (int)maxY - (int)iniY - 6;
                    if (wi > 0 && hi > 0) candCap += (long)((wi + 1) / 2) * ((hi + 1) / 2);
                    L.nCells++;
                }
            }
        } else {
            L.nCols = L.nRows = 0; L.wCell = L.hCell = 1;
        }
        cellBase += L.nCells;
        L.candCap = (int)candCap;
        L.candOff = (unsigned)candOff;
        candOff += align_up((size_t)candCap, 64);
        const int regW = L.maxBX - FAST_BORDER, regH = L.maxBY - FAST_BORDER;
        L.nIni = (regW > 0 && regH > 0) ? (int)roundf((float)regW / (float)regH) : 0;
        L.hX = L.nIni > 0 ? (float)regW / (float)L.nIni : 1.f;
        if (L.nCells > 0 && L.nIni <= 0) { err = "portrait level with nIni == 0: the reference divides by zero here"; return false; }
        L.maxNodes = std::max(L.quota + 3, 4 * std::max(L.nIni, 1)) + 1;
        if (L.maxNodes > 65535 || L.maxBX > 4096 + FAST_BORDER || L.maxBY > 4096 + FAST_BORDER) { err = "image or quota too large (limits: 4096 px, 65535 nodes/level)"; return false; }
        L.kpCap = L.maxNodes;
        L.kpOff = kpOff;
        kpOff += L.kpCap;
        L.xtabOff = xt; L.ytabOff = yt;
        if (l > 0) { xt += L.w; yt += L.h; }
    }
    g.pyrBytes = (unsigned)off;
    g.candPerImg = (unsigned)candOff;
    g.kpPerImg = kpOff;
    g.totalCells = cellBase;
    return true;
}

bool ensure_scratch(Ctx& c, size_t dev_bytes, size_t host_bytes)
{
    if (dev_bytes > c.d_scratch_bytes) {
        c.allocEpoch++;
        cudaStreamSynchronize(c.stream);
        if (c.d_scratch) cudaFree(c.d_scratch);
        c.d_scratch = nullptr; c.d_scratch_bytes = 0;
        const size_t nb = align_up(dev_bytes + dev_bytes / 4, 1 << 20);
        if (cudaMalloc(&c.d_scratch, nb) != cudaSuccess) { c.err = "cudaMalloc(scratch) failed"; return false; }
        c.d_scratch_bytes = nb;
    }
    if (host_bytes > c.h_scratch_bytes) {
        c.allocEpoch++;
        cudaStreamSynchronize(c.stream);
        if (c.h_scratch) cudaFreeHost(c.h_scratch);
        c.h_scratch = nullptr; c.h_scratch_bytes = 0;
        const size_t nb = align_up(host_bytes + host_bytes / 4, 1 << 20);
        if (cudaMallocHost(&c.h_scratch, nb) != cudaSuccess) { c.err = "cudaMallocHost(scratch) failed"; return false; }
        c.h_scratch_bytes = nb;
    }
    return true;
}

// Per-shape tables: resize coefficients (OpenCV resize.cpp, INTER_RESIZE_COEF_BITS = 11) and FAST cells.
const ShapeTables* get_shape(Ctx& c, int w, int h)
{
    auto it = c.shapes.find({w, h});
    if (it != c.shapes.end()) return &it->second;
    ShapeTables st;
    // max_w / max_h of orbb200_create size per-row tables too (the stereo row index has max_h + 2 entries per frame): a taller or
    // wider image is refused even when its pools would fit
    if (w > c.maxW || h > c.maxH) { c.err = "image shape exceeds the context's max_w/max_h"; return nullptr; }
    if (!build_geom(c, w, h, st.g, c.err)) return nullptr;
    const Geom& g = st.g;
    if (g.pyrBytes > c.gmax.pyrBytes || g.candPerImg > c.gmax.candPerImg + c.gmax.candPerImg / 8 || g.kpPerImg > c.gmax.kpPerImg) {
        c.err = "image shape exceeds the context's max_w/max_h";
        return nullptr;
    }
    std::vector<int2> xtab;
    std::vector<int4> ytab;
    for (int l = 1; l < g.nlevels; l++) {
        const LevelGeom &S = g.lv[l - 1], &D = g.lv[l];
        if (D.w <= 0 || D.h <= 0) continue;
        const double sx_ = (double)S.w / D.w, sy_ = (double)S.h / D.h;
        const size_t xbase = xtab.size();
        for (int dx = 0; dx < D.w; dx++) {
            float fx = (float)((dx + 0.5) * sx_ - 0.5);
            int sx = cvFloorF(fx);
            fx -= sx;
            if (sx < 0) { fx = 0; sx = 0; }
            if (sx >= S.w - 1) { fx = 0; sx = S.w - 1; }
            const int a0 = std::min(std::max(cvRoundF((1.f - fx) * 2048), -32768), 32767);
            const int a1 = std::min(std::max(cvRoundF(fx * 2048), -32768), 32767);
            xtab.push_back(make_int2(sx, (a0 & 0xffff) | (a1 << 16)));
        }
        bool narrow = true;     // resize_kernel<NARROW>: groups of 4 output columns (x4 = 0, 4, ...) within an 8-byte source window
        for (int x4 = 0; x4 < D.w; x4 += 4) {
            const int first = xtab[xbase + x4].x, last = std::min(xtab[xbase + std::min(x4 + 3, D.w - 1)].x + 1, S.w - 1);
            if (last - first > 7) narrow = false;
        }
        st.resizeNarrow[l] = narrow;
        for (int dy = 0; dy < D.h; dy++) {
            float fy = (float)((dy + 0.5) * sy_ - 0.5);
            int sy = cvFloorF(fy);
            fy -= sy;
            const int b0 = std::min(std::max(cvRoundF((1.f - fy) * 2048), -32768), 32767);
            const int b1 = std::min(std::max(cvRoundF(fy * 2048), -32768), 32767);
            ytab.push_back(make_int4(std::min(std::max(sy, 0), S.h - 1), std::min(std::max(sy + 1, 0), S.h - 1), b0, b1));
        }
    }
    std::vector<int4> cells;
    for (int l = 0; l < g.nlevels; l++) {
        const LevelGeom& L = g.lv[l];
        if (L.nCells == 0) continue;
        for (int i = 0; i < L.nRows; i++) {
            const float iniY = (float)(FAST_BORDER + i * L.hCell);
            float maxY = iniY + L.hCell + 6;
            if (iniY >= L.maxBY - 3) continue;
            if (maxY > L.maxBY) maxY = (float)L.maxBY;
            for (int j = 0; j < L.nCols; j++) {
                const float iniX = (float)(FAST_BORDER + j * L.wCell);
                float maxX = iniX + L.wCell + 6;
                if (iniX >= L.maxBX - 6) continue;
                if (maxX > L.maxBX) maxX = (float)L.maxBX;
                push_fast_cell(cells, st.fastSmem, (int)iniX, (int)iniY, (int)maxX, (int)maxY, l, L.off, L.pitch, L.candOff, L.candCap);
            }
        }
    }
    // the same cells, grouped for fast_strip_kernel: the cells of a row that the reference keeps (a prefix of the columns), split
    // evenly into groups whose image (first iniX rounded down to 4 .. last maxX) fits the 136-pixel tile
    std::vector<int4> groups;
    for (int l = 0; l < g.nlevels; l++) {
        const LevelGeom& L = g.lv[l];
        if (L.nCells == 0) continue;
        int kept = 0;                                       // columns with iniX < maxBX - 6
        while (kept < L.nCols && FAST_BORDER + kept * L.wCell < L.maxBX - 6) kept++;
        if (kept == 0) continue;
        const int gmax = std::max(1, std::min(FS_MAX_CELLS, (2 * (FS_PITCH - 3) - 9) / L.wCell));
        const int ngr = (kept + gmax - 1) / gmax;
        for (int i = 0; i < L.nRows; i++) {
            const int iniY = FAST_BORDER + i * L.hCell;
            if (iniY >= L.maxBY - 3) continue;
            const int maxY = std::min(iniY + L.hCell + 6, L.maxBY);
            for (int q = 0; q < ngr; q++) {
                const int j0 = (int)((long)q * kept / ngr), j1 = (int)((long)(q + 1) * kept / ngr);     // cells [j0, j1)
                const int x0 = FAST_BORDER + j0 * L.wCell;
                const int x1 = std::min(FAST_BORDER + (j1 - 1) * L.wCell + L.wCell + 6, L.maxBX);
                push_fast_group(groups, st.fastStrip, x0, iniY, x1, maxY, l, j1 - j0, L.wCell, L.off, L.pitch, L.candOff, L.candCap);
            }
        }
    }
    st.nFastGroups = (int)(groups.size() / 3);
    {   // groups are level-major: first group of each level
        int gi = 0;
        for (int l = 0; l <= g.nlevels; l++) {
            while (gi < st.nFastGroups && groups[3 * (size_t)gi].z < l) gi++;
            st.fastGroupBase[l] = gi;
        }
    }
    std::vector<int4> btiles, rtiles;
    for (int l = 0; l < g.nlevels; l++) {
        const LevelGeom& L = g.lv[l];
        for (int y0 = 0; y0 < L.h; y0 += BL_ROWS)
            for (int x0 = 0; x0 < L.w; x0 += 128) btiles.push_back(make_int4(l, x0, y0, 0));
        if (l > 0)
            for (int y0 = 0; y0 < L.h; y0 += RS_ROWS)
                for (int x0 = 0; x0 < L.w; x0 += 128) {
                    // {x0, y0, first staged source row, rows}, {first staged source column (16-aligned), 16-byte vectors per
                    // row, 2^16 / vectors + 1, -}: the window of resize_kernel, so that its staging loads depend on one load
                    const LevelGeom& Sg = g.lv[l - 1];
                    const int4* yt = ytab.data() + L.ytabOff;
                    const int2* xt = xtab.data() + L.xtabOff;
                    const int yLast = std::min(y0 + RS_ROWS, L.h) - 1, xLast = std::min(x0 + 127, L.w - 1);
                    const int ry0 = yt[y0].x, ry1 = yt[yLast].y;
                    const int cx0 = xt[x0].x & ~15, cx1 = std::min(xt[xLast].x + 1, Sg.w - 1);
                    const int nvec = (cx1 - cx0) / 16 + 1;
                    rtiles.push_back(make_int4(x0, y0, ry0, ry1 - ry0 + 1));
                    rtiles.push_back(make_int4(cx0, nvec, (int)(65536u / (unsigned)nvec + 1u), 0));
                    st.resizeTileCount[l]++;
                }
        st.resizeTileBase[l + 1] = (int)rtiles.size() / 2;
        if (l > 0 && L.w > 0 && L.h > 0) {
            // shared-memory window of a 128 x RS_ROWS tile: source columns (16-byte aligned start) and rows it touches
            const LevelGeom& Sg = g.lv[l - 1];
            const int span = (int)ceil(128.0 * Sg.w / L.w) + 3 + 15, rows = (int)ceil((double)RS_ROWS * Sg.h / L.h) + 3;
            st.resizeSmemPitch[l] = (int)align_up((size_t)span + 16, 16);
            st.resizeSmemRows[l] = rows;
        }
    }
    // TMA descriptors of the blurred pool (describe_kernel fetches a keypoint's 64 x 39-byte box with one bulk tensor copy)
    {
        typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
            c.err = "cuTensorMapEncodeTiled is not available from this driver";
            return nullptr;
        }
        memset(&st.dmaps, 0, sizeof(st.dmaps));
        memset(&st.rmaps, 0, sizeof(st.rmaps));
        for (int l = 1; l < g.nlevels; l++) {             // resize level l stages a window of level l-1
            const LevelGeom& S = g.lv[l - 1];
            if (S.w <= 0 || S.h <= 0 || st.resizeTileCount[l] == 0 || st.resizeSmemPitch[l] > 256 || st.resizeSmemRows[l] > 256) continue;
            const cuuint64_t dims[3] = {(cuuint64_t)(S.pitch - PYR_MARGIN_X), (cuuint64_t)(S.h + PYR_MARGIN_Y), (cuuint64_t)c.maxBatch};
            const cuuint64_t strides[2] = {(cuuint64_t)S.pitch, (cuuint64_t)g.pyrBytes};
            const cuuint32_t box[3] = {(cuuint32_t)st.resizeSmemPitch[l], (cuuint32_t)st.resizeSmemRows[l], 1}, estr[3] = {1, 1, 1};
            st.resizeMapOk[l] = ((EncodeTiled)fn)(&st.rmaps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, c.d_pyr + S.off, dims, strides, box, estr,
                                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        }
        for (int l = 0; l < g.nlevels; l++) {
            const LevelGeom& L = g.lv[l];
            if (L.w <= 0 || L.h <= 0) continue;
            const cuuint64_t dims[3] = {(cuuint64_t)(L.pitch - PYR_MARGIN_X), (cuuint64_t)(L.h + PYR_MARGIN_Y), (cuuint64_t)c.maxBatch};
            const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)g.pyrBytes};
            const cuuint32_t box[3] = {DS_TBOX_W, DS_TBOX_H, 1}, estr[3] = {1, 1, 1};
            const CUresult r = ((EncodeTiled)fn)(&st.dmaps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, c.d_blur + L.off, dims, strides, box, estr,
                                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS && L.nCells > 0) {      // levels without FAST cells never hold a keypoint
                c.err = "cuTensorMapEncodeTiled failed for pyramid level " + std::to_string(l) + " (" + std::to_string((int)r) + ")";
                return nullptr;
            }
        }
    }
    st.nBlurTiles = (int)btiles.size();
    st.nFastCells = (int)(cells.size() / 3);
    auto up = [&](void** dptr, const void* src, size_t bytes) -> bool {
        if (bytes == 0) bytes = 16;
        if (cudaMalloc(dptr, bytes) != cudaSuccess) return false;
        if (src && cudaMemcpyAsync(*dptr, src, bytes, cudaMemcpyHostToDevice, c.stream) != cudaSuccess) return false;
        return true;
    };
    bool ok = up((void**)&st.d_xtab, xtab.data(), xtab.size() * sizeof(int2)) &&
              up((void**)&st.d_ytab, ytab.data(), ytab.size() * sizeof(int4)) &&
              up((void**)&st.d_cells, cells.data(), cells.size() * sizeof(int4)) &&
              up((void**)&st.d_groups, groups.data(), groups.size() * sizeof(int4)) &&
              up((void**)&st.d_blurTiles, btiles.data(), btiles.size() * sizeof(int4)) &&
              up((void**)&st.d_resizeTiles, rtiles.data(), rtiles.size() * sizeof(int4)) &&
              up((void**)&st.d_dmaps, &st.dmaps, sizeof(st.dmaps)) &&
              up((void**)&st.d_rmaps, &st.rmaps, sizeof(st.rmaps));
    cudaStreamSynchronize(c.stream);   // host vectors go out of scope
    if (!ok) { c.err = "cudaMalloc(shape tables) failed"; return nullptr; }
    auto res = c.shapes.emplace(std::make_pair(w, h), st);
    return &res.first->second;
}

StageTimer::StageTimer(Ctx& c_, int stage) : c(c_), on(c_.timing)
{
    if (!on) return;
    auto get = [&]() {
        cudaEvent_t e;
        if (!c.freeEvents.empty()) { e = c.freeEvents.back(); c.freeEvents.pop_back(); }
        else cudaEventCreate(&e);
        return e;
    };
    p.stage = stage; p.e0 = get(); p.e1 = get();
    cudaEventRecord(p.e0, c.stream);
}

StageTimer::~StageTimer()
{
    if (!on) return;
    cudaEventRecord(p.e1, c.stream);
    c.pending.push_back(p);
    if (c.pending.size() >= 2048) drain_stage_events(c);
}

void drain_stage_events(Ctx& c)
{
    if (c.pending.empty()) return;
    cudaStreamSynchronize(c.stream);
    for (auto& p : c.pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) { c.stageMs[p.stage] += ms; c.stageGroups[p.stage]++; }
        c.freeEvents.push_back(p.e0); c.freeEvents.push_back(p.e1);
    }
    c.pending.clear();
}

// fork == true: the blur (FMA pipe + memory bound; only the descriptors need it) runs on a second stream beside
// FAST + octree (integer-ALU bound) and joins before the descriptor kernel.  In a captured graph this becomes
// two parallel branches.  With per-stage timing the stages run back to back on one stream.
//
// One or two images (the drop-in's one-frame-per-call pattern) are latency bound: seven dependent resize launches, then FAST, the
// octree (one CTA per level) and the descriptors.  Grid FAST reads no border pixel and level 0 needs no resize, so for them FAST +
// octree of level 0 -- the longest of the per-level CTAs -- run on a third branch beside the resize chain; the other levels follow
// the chain, the border fill + blur take the second branch, and the descriptor kernel joins all three.  Within a branch the
// kernels are programmatic dependent launches (see pdl_wait): a successor's CTAs are resident and past their prologue when the
// predecessor's last CTA retires.
static void enqueue_extract_kernels(Ctx& c, int n, bool fork, const HostStage* hs = nullptr)
{
    const Geom& g = c.cur->g;
    const HostMirror* mirror = hs ? &hs->mirror : nullptr;
    launch_clear_counters(c, n);
    uint8_t* hostPyr = hs ? hs->hostPyr : nullptr;
    if (hs) launch_import_host(c, hs->imgs, hs->imgBytes, hs->rowBytes, n, hostPyr);
    if (fork && n <= 2 && c.splitLevel0 && !c.fastCells && c.cur->nFastGroups > 0 && g.nlevels > 1) {
        cudaEventRecord(c.evFork0, c.stream);
        cudaStreamWaitEvent(c.stream3, c.evFork0, 0);
        launch_fast_levels(c, n, 0, 1, c.stream3, false);
        launch_octree_levels(c, n, 0, 1, c.stream3, true);
        cudaEventRecord(c.evJoin0, c.stream3);
        launch_resizes(c, n, c.stream, hostPyr);
        cudaEventRecord(c.evFork, c.stream);
        cudaStreamWaitEvent(c.stream2, c.evFork, 0);
        launch_border(c, n, c.stream2, false);
        // a full dependency, not a programmatic one: the border bytes share cache lines with pixels that FAST CTAs on the other branch
        // may have pulled into an SM's L1 before border_kernel wrote them; only a kernel boundary is documented to drop such lines
        launch_blur(c, n, c.stream2, false);
        cudaEventRecord(c.evJoin, c.stream2);
        launch_fast_levels(c, n, 1, g.nlevels, c.stream, true);
        if (std::getenv("ORBB200_OCTREE_PER_LEVEL")) {       // profiling aid: one launch per level (per-level durations under ncu)
            for (int l = 1; l < g.nlevels; l++) launch_octree_levels(c, n, l, l + 1, c.stream, false);
        } else
        launch_octree_levels(c, n, 1, g.nlevels, c.stream, true);
        cudaStreamWaitEvent(c.stream, c.evJoin0, 0);
        cudaStreamWaitEvent(c.stream, c.evJoin, 0);
        launch_describe(c, n, false, mirror);
        return;
    }
    { StageTimer t(c, 1); launch_pyramid(c, n, hostPyr); }   // includes the border fill
    if (fork) {
        cudaEventRecord(c.evFork, c.stream);
        cudaStreamWaitEvent(c.stream2, c.evFork, 0);
        launch_blur(c, n, c.stream2);
        cudaEventRecord(c.evJoin, c.stream2);
        launch_fast(c, n, true);
        launch_octree(c, n);
        cudaStreamWaitEvent(c.stream, c.evJoin, 0);
        launch_describe(c, n, false, mirror);
        return;
    }
    { StageTimer t(c, 2); launch_fast(c, n); }
    { StageTimer t(c, 3); launch_blur(c, n, c.stream); }
    { StageTimer t(c, 4); launch_octree(c, n); }
    { StageTimer t(c, 5); launch_describe(c, n, false, mirror); }
}

// The ~20 launches of one extraction are captured once per (shape, image count) into a CUDA graph and replayed:
// for single frames (the real-time use of the drop-in) the CPU launch cost is a large part of the latency.
// Per-stage timing needs events between the kernels, so it uses the plain launches.
//
// hs != nullptr (a small host call): the images are in the pinned staging block and the results are wanted in its mirror half; the
// upload (import_host_kernel) and the delivery (describe_kernel's second set of stores) are then part of the same graph, keyed
// separately and re-captured if the staging block has moved.
static int run_extract(Ctx& c, int n, const HostStage* hs = nullptr)
{
    c.stereoValid = false;
    c.pyrMirrorFresh = nullptr;
    if (c.timing || !c.useGraphs) {
        enqueue_extract_kernels(c, n, !c.timing && c.forkBlur, hs);
        ORBB200_CUDA_OK(c, cudaGetLastError());
        return ORBB200_OK;
    }
    {   // inside somebody else's capture (a whole frame step being recorded): the kernels become nodes of that graph
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(c.stream, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusActive) {
            enqueue_extract_kernels(c, n, c.forkBlur, hs);
            return ORBB200_OK;
        }
    }
    ShapeTables* st = const_cast<ShapeTables*>(c.cur);
    const int key = hs ? n + 65536 + (hs->hostPyr ? 131072 : 0) : n;
    auto it = st->graphs.find(key);
    if (it != st->graphs.end() && hs && it->second.stage != hs->imgs) {       // the pinned block was reallocated since the capture
        cudaGraphExecDestroy(it->second.exec);
        st->graphs.erase(it);
        it = st->graphs.end();
    }
    if (it == st->graphs.end()) {
        // warm the lazily configured kernel attributes outside the capture, then capture
        const long long before = c.launches;
        enqueue_extract_kernels(c, n, c.forkBlur, hs);
        ORBB200_CUDA_OK(c, cudaGetLastError());
        const long long perRun = c.launches - before;
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        ORBB200_CUDA_OK(c, cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeThreadLocal));
        enqueue_extract_kernels(c, n, c.forkBlur, hs);
        c.launches -= perRun;                      // the capture pass enqueues nothing
        cudaError_t e = cudaStreamEndCapture(c.stream, &graph);
        if (e == cudaSuccess) e = cudaGraphInstantiate(&exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (e != cudaSuccess) {                    // fall back to plain launches for this context (still the CUDA path)
            cudaGetLastError();
            c.useGraphs = false;
            return ORBB200_OK;                     // the warm-up run above already produced this call's result
        }
        st->graphs[key] = {exec, perRun, hs ? hs->imgs : nullptr};
        return ORBB200_OK;                         // ditto
    }
    ORBB200_CUDA_OK(c, cudaGraphLaunch(it->second.exec, c.stream));
    c.launches += it->second.launches;
    return ORBB200_OK;
}

// after the stream has drained: copy staged results of a small host-buffer frame step into the caller's buffers
void deliver_host_copies(Ctx& c)
{
    for (const Ctx::HostCopy& h : c.hostCopies)
        for (size_t r = 0; r < h.rows; r++)
            memcpy(static_cast<uint8_t*>(h.dst) + r * h.dpitch, static_cast<const uint8_t*>(h.src) + r * h.width, h.width);
    c.hostCopies.clear();
}

static int check_status(Ctx& c)
{
    int32_t st = 0;
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(&st, c.d_status, sizeof(st), cudaMemcpyDeviceToHost, c.stream));
    ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
    if (st != 0) {
        cudaMemsetAsync(c.d_status, 0, sizeof(int32_t), c.stream);
        c.err = "device-side overflow in octree distribution (status " + std::to_string(st) + ")";
        return ORBB200_ERR_UNSUPPORTED;
    }
    return ORBB200_OK;
}

}  // namespace orbb200

namespace {
struct Arena {
    uint8_t* base;
    size_t cap, off = 0;
    Arena(uint8_t* b, size_t c) : base(b), cap(c) {}
    template <class T> T* take(size_t n) { off = align_up(off, 256); T* p = reinterpret_cast<T*>(base + off); off += n * sizeof(T); return p; }
};
}  // namespace

#define CTX_ENTER(ctx)                                             \
    if (!(ctx)) return ORBB200_ERR_ARG;                            \
    Ctx& c = (ctx)->c;                                             \
    ORBB200_CUDA_OK(c, cudaSetDevice(c.device))

extern "C" {

int orbb200_create(orbb200_ctx** out, int device, int nfeatures, float scaleFactor, int nlevels,
                   int iniThFAST, int minThFAST, int max_w, int max_h, int max_batch)
{
    std::lock_guard<std::mutex> lk(g_create_mu);
    if (!out || nfeatures <= 0 || nlevels <= 0 || nlevels > MAX_LEVELS || max_w <= 0 || max_h <= 0 || max_batch <= 0 ||
        !(scaleFactor > 1.0f) || iniThFAST <= 0 || minThFAST <= 0 || iniThFAST > 254 || minThFAST > iniThFAST) {
        g_create_err = "orbb200_create: bad argument";
        return ORBB200_ERR_ARG;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev) {
        g_create_err = std::string("orbb200_create: no usable CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "bad index") + ")";
        return ORBB200_ERR_CUDA;
    }
    orbb200_ctx* h = new orbb200_ctx;
    Ctx& c = h->c;
    c.device = device;
    auto fail = [&](const std::string& m, int code) { g_create_err = m; orbb200_destroy(h); return code; };
    if (cudaSetDevice(device) != cudaSuccess) return fail("cudaSetDevice failed", ORBB200_ERR_CUDA);
    if (cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c.stream2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c.evFork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c.evJoin, cudaEventDisableTiming) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c.stream3, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c.stream4, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c.evBirdCarry, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c.evFork4, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c.evJoin4, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c.evFork0, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c.evJoin0, cudaEventDisableTiming) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c.streamBird, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c.evBirdFork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c.evBirdJoin, cudaEventDisableTiming) != cudaSuccess)
        return fail("cudaStreamCreate failed", ORBB200_ERR_CUDA);
    c.forkBlur = std::getenv("ORBB200_SERIAL") == nullptr;
    c.pdl = std::getenv("ORBB200_NO_PDL") == nullptr;
    c.splitLevel0 = std::getenv("ORBB200_NO_SPLIT") == nullptr;
    c.octreeSmemCand = std::getenv("ORBB200_NO_OCTREE_SMEM") == nullptr;
    c.hostGraph = std::getenv("ORBB200_NO_HOST_GRAPH") == nullptr;
    c.frameGraph = std::getenv("ORBB200_NO_FRAME_GRAPH") == nullptr;
    c.subpixGenericWarp = std::getenv("ORBB200_SUBPIX_GENERIC") != nullptr;
    c.selectTiers = std::getenv("ORBB200_SELECT_TIERS") != nullptr;
    c.stageMatch = std::getenv("ORBB200_NO_STAGED_MATCH") == nullptr;
    c.warpCands = std::getenv("ORBB200_NO_WARP_CANDS") == nullptr;
    if (const char* e = std::getenv("ORBB200_STAGE_MAX_KB")) c.stageMaxBytes = (size_t)std::max(0, atoi(e)) << 10;
    c.forkBird = std::getenv("ORBB200_FORK_BIRD") != nullptr;      // measured: beside the front extraction it is 3 % SLOWER than after it (6.86 vs 6.65 ms per 128 frames)
    c.fastCells = std::getenv("ORBB200_FAST_CELLS") != nullptr;
    c.stageUploads = std::getenv("ORBB200_NO_STAGED_UPLOAD") == nullptr;
    if (const char* e = std::getenv("ORBB200_SUBPIX_CTAS")) c.subpixCtasPerSm = std::max(1, std::min(4, atoi(e)));
    // ORBextractor::ORBextractor (reference src/ORBextractor.cc:410-446)
    c.nfeatures = nfeatures; c.scaleFactor = scaleFactor; c.nlevels = nlevels; c.iniTh = iniThFAST; c.minTh = minThFAST;
    c.scale.resize(nlevels); c.sigma2.resize(nlevels); c.invScale.resize(nlevels); c.invSigma2.resize(nlevels); c.quota.resize(nlevels);
    c.scale[0] = 1.0f; c.sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        c.scale[i] = (float)(c.scale[i - 1] * c.scaleFactor);
        c.sigma2[i] = c.scale[i] * c.scale[i];
    }
    for (int i = 0; i < nlevels; i++) { c.invScale[i] = 1.0f / c.scale[i]; c.invSigma2[i] = 1.0f / c.sigma2[i]; }
    {
        float factor = (float)(1.0f / c.scaleFactor);
        float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
        int sum = 0;
        for (int l = 0; l < nlevels - 1; l++) {
            c.quota[l] = cvRoundF(nDesired);
            sum += c.quota[l];
            nDesired *= factor;
        }
        c.quota[nlevels - 1] = std::max(nfeatures - sum, 0);
    }
    c.maxW = max_w; c.maxH = max_h; c.maxBatch = max_batch;
    std::string err;
    if (!build_geom(c, max_w, max_h, c.gmax, err)) return fail("orbb200_create: " + err, ORBB200_ERR_UNSUPPORTED);
    const Geom& g = c.gmax;
    const size_t candSlots = (size_t)g.candPerImg + g.candPerImg / 8 + 64;   // slack: smaller shapes may tile slightly worse
    const size_t nb = (size_t)max_batch;
    bool ok = true;
    auto alloc = [&](void** p, size_t bytes) { if (ok && cudaMalloc(p, std::max<size_t>(bytes, 256)) != cudaSuccess) ok = false; };
    alloc((void**)&c.d_pyr, nb * g.pyrBytes + 4096);
    alloc((void**)&c.d_blur, nb * g.pyrBytes + 4096);
    alloc((void**)&c.d_cand, nb * candSlots * sizeof(uint32_t));
    alloc((void**)&c.d_nodeOf, nb * candSlots * sizeof(uint16_t));
    alloc((void**)&c.d_candCount, nb * MAX_LEVELS * sizeof(int32_t));
    alloc((void**)&c.d_lvlKp, nb * g.kpPerImg * sizeof(uint32_t));
    alloc((void**)&c.d_lvlCount, nb * MAX_LEVELS * sizeof(int32_t));
    alloc((void**)&c.d_kps, nb * g.kpPerImg * sizeof(orbb200_kp_t));
    alloc((void**)&c.d_desc, nb * g.kpPerImg * 32);
    alloc((void**)&c.d_counts, nb * sizeof(int32_t));
    alloc((void**)&c.d_status, 256);
    alloc((void**)&c.d_uRight, nb * g.kpPerImg * sizeof(float));
    alloc((void**)&c.d_depth, nb * g.kpPerImg * sizeof(float));
    alloc((void**)&c.d_sad, nb * g.kpPerImg * sizeof(int32_t));
    alloc((void**)&c.d_nKept, nb * sizeof(int32_t));
    alloc((void**)&c.d_invScale, MAX_LEVELS * sizeof(float));
    {   // stereo row table: a right keypoint of octave o spans at most 2*ceil(2*scale[o]) + 2 rows
        const int span = 2 * (int)ceilf(2.0f * c.scale[nlevels - 1]) + 3;
        c.stereoItemCap = g.kpPerImg * span;
        const size_t nf = nb / 2 + 1;
        alloc((void**)&c.d_rowStart, nf * (size_t)(max_h + 2) * sizeof(int32_t));
        alloc((void**)&c.d_rowItems, nf * (size_t)c.stereoItemCap * sizeof(int32_t));
    }
    if (!ok) return fail("orbb200_create: cudaMalloc of the device pools failed", ORBB200_ERR_CUDA);
    cudaMemsetAsync(c.d_status, 0, 256, c.stream);
    cudaMemcpyAsync(c.d_invScale, c.invScale.data(), sizeof(float) * nlevels, cudaMemcpyHostToDevice, c.stream);
    cudaMemsetAsync(c.d_counts, 0, nb * sizeof(int32_t), c.stream);
    cudaMemsetAsync(c.d_lvlCount, 0, nb * MAX_LEVELS * sizeof(int32_t), c.stream);
    // pyramid pools are read up to the row pitch (never past it): start from defined bytes
    cudaMemsetAsync(c.d_pyr, 0, nb * g.pyrBytes, c.stream);
    cudaMemsetAsync(c.d_blur, 0, nb * g.pyrBytes, c.stream);
    if (cudaStreamSynchronize(c.stream) != cudaSuccess) return fail("orbb200_create: device initialisation failed", ORBB200_ERR_CUDA);
    *out = h;
    return ORBB200_OK;
}

void orbb200_destroy(orbb200_ctx* ctx)
{
    if (!ctx) return;
    Ctx& c = ctx->c;
    cudaSetDevice(c.device);
    if (c.stream) cudaStreamSynchronize(c.stream);
    drain_stage_events(c);
    bird_destroy(c);
    for (cudaEvent_t e : c.freeEvents) cudaEventDestroy(e);
    for (auto& p : c.plans) cudaFree(p.block);
    for (auto& p : c.framePlans) cudaFree(p.block);
    if (c.d_fstep) cudaFree(c.d_fstep);
    void* ptrs[] = {c.d_pyr, c.d_blur, c.d_cand, c.d_nodeOf, c.d_candCount, c.d_lvlKp, c.d_lvlCount, c.d_kps, c.d_desc, c.d_counts, c.d_status, c.d_scratch, c.d_step,
                    c.d_uRight, c.d_depth, c.d_sad, c.d_nKept, c.d_invScale, c.d_rowStart, c.d_rowItems};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (auto& fg : c.frameGraphs) if (fg.exec) cudaGraphExecDestroy(fg.exec);
    for (auto& kv : c.shapes) { cudaFree(kv.second.d_xtab); cudaFree(kv.second.d_ytab); cudaFree(kv.second.d_cells); cudaFree(kv.second.d_groups); cudaFree(kv.second.d_blurTiles); cudaFree(kv.second.d_resizeTiles); cudaFree(kv.second.d_dmaps); cudaFree(kv.second.d_rmaps);
                               for (auto& gk : kv.second.graphs) cudaGraphExecDestroy(gk.second.exec); }
    if (c.h_scratch) cudaFreeHost(c.h_scratch);
    if (c.h_pyrMirror) cudaFreeHost(c.h_pyrMirror);
    if (c.evFork) cudaEventDestroy(c.evFork);
    if (c.evJoin) cudaEventDestroy(c.evJoin);
    if (c.stream2) cudaStreamDestroy(c.stream2);
    if (c.evFork0) cudaEventDestroy(c.evFork0);
    if (c.evJoin0) cudaEventDestroy(c.evJoin0);
    if (c.stream3) cudaStreamDestroy(c.stream3);
    if (c.evBirdCarry) cudaEventDestroy(c.evBirdCarry);
    if (c.evFork4) cudaEventDestroy(c.evFork4);
    if (c.evJoin4) cudaEventDestroy(c.evJoin4);
    if (c.stream4) cudaStreamDestroy(c.stream4);
    if (c.evBirdFork) cudaEventDestroy(c.evBirdFork);
    if (c.evBirdJoin) cudaEventDestroy(c.evBirdJoin);
    if (c.streamBird) cudaStreamDestroy(c.streamBird);
    if (c.stream) cudaStreamDestroy(c.stream);
    delete ctx;
}

const char* orbb200_last_error(const orbb200_ctx* ctx) { return ctx ? ctx->c.err.c_str() : g_create_err.c_str(); }


int orbb200_sync(orbb200_ctx* ctx)
{
    CTX_ENTER(ctx);
    ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
    deliver_host_copies(c);
    return ORBB200_OK;
}

void* orbb200_stream(orbb200_ctx* ctx) { return ctx ? (void*)ctx->c.stream : nullptr; }
int orbb200_get_levels(const orbb200_ctx* ctx) { return ctx ? ctx->c.nlevels : ORBB200_ERR_ARG; }

int orbb200_get_scale_table(const orbb200_ctx* ctx, int which, float* out)
{
    if (!ctx || !out || which < 0 || which > 3) return ORBB200_ERR_ARG;
    const Ctx& c = ctx->c;
    const std::vector<float>& v = which == 0 ? c.scale : which == 1 ? c.invScale : which == 2 ? c.sigma2 : c.invSigma2;
    std::copy(v.begin(), v.end(), out);
    return ORBB200_OK;
}

int orbb200_get_features_per_level(const orbb200_ctx* ctx, int32_t* out)
{
    if (!ctx || !out) return ORBB200_ERR_ARG;
    std::copy(ctx->c.quota.begin(), ctx->c.quota.end(), out);
    return ORBB200_OK;
}

int orbb200_max_keypoints(const orbb200_ctx* ctx) { return ctx ? ctx->c.gmax.kpPerImg : ORBB200_ERR_ARG; }
long long orbb200_launch_count(const orbb200_ctx* ctx) { return ctx ? ctx->c.launches : 0; }

// ---- extraction -------------------------------------------------------------------------------------
int orbb200_extract_device(orbb200_ctx* ctx, const uint8_t* d_imgs, size_t img_bytes, int n, int w, int h, size_t stride)
{
    CTX_ENTER(ctx);
    if (!d_imgs || n <= 0 || n > c.maxBatch || w <= 0 || h <= 0 || stride < (size_t)w) { c.err = "extract: bad argument"; return ORBB200_ERR_ARG; }
    const ShapeTables* st = get_shape(c, w, h);
    if (!st) return c.err.find("exceeds") != std::string::npos ? ORBB200_ERR_ARG : ORBB200_ERR_UNSUPPORTED;
    c.cur = st; c.curN = n;
    { StageTimer t(c, 0); launch_import(c, d_imgs, img_bytes, stride, n); }
    return run_extract(c, n);
}

// Results of small batches (the one-frame-per-call pattern of Frame::ExtractORB) come back through a pinned staging block: the
// status word, the counts and the keypoint / descriptor rows are four asynchronous copies behind ONE synchronisation, and only the
// n_out[i] valid records of each image are then copied into the caller's (pageable) buffers.  Copies straight into pageable memory
// are four serialised round trips of the copy engine's bounce buffer.

int orbb200_download_results(orbb200_ctx* ctx, int n, orbb200_kp_t* kps, uint8_t* desc, int cap_per_img, int* n_out)
{
    CTX_ENTER(ctx);
    if (!c.cur || n <= 0 || n > c.curN || !kps || !desc || !n_out || cap_per_img <= 0) { c.err = "download: bad argument"; return ORBB200_ERR_ARG; }
    if (!c.hostCopies.empty()) { ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream)); deliver_host_copies(c); }   // staged results of an earlier frame step
    const int kpi = c.cur->g.kpPerImg;
    const int take = std::min(cap_per_img, kpi);
    const size_t rowK = (size_t)take * sizeof(orbb200_kp_t), rowD = (size_t)take * 32, cntB = align_up(sizeof(int32_t) * (size_t)n, 64);
    const size_t stageBytes = 64 + cntB + (size_t)n * (rowK + rowD);
    if (stageBytes <= STAGE_LIMIT - STAGE_D2H_OFF && ensure_scratch(c, 0, STAGE_LIMIT)) {
        uint8_t* hs = c.h_scratch + STAGE_D2H_OFF;
        int32_t* hStatus = reinterpret_cast<int32_t*>(hs);
        int32_t* hCnt = reinterpret_cast<int32_t*>(hs + 64);
        uint8_t* hK = hs + 64 + cntB;
        uint8_t* hD = hK + (size_t)n * rowK;
        ORBB200_CUDA_OK(c, cudaMemcpyAsync(hStatus, c.d_status, sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
        ORBB200_CUDA_OK(c, cudaMemcpyAsync(hCnt, c.d_counts, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c.stream));
        ORBB200_CUDA_OK(c, cudaMemcpy2DAsync(hK, rowK, c.d_kps, (size_t)kpi * sizeof(orbb200_kp_t), rowK, n, cudaMemcpyDeviceToHost, c.stream));
        ORBB200_CUDA_OK(c, cudaMemcpy2DAsync(hD, rowD, c.d_desc, (size_t)kpi * 32, rowD, n, cudaMemcpyDeviceToHost, c.stream));
        ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
        if (*hStatus != 0) {
            cudaMemsetAsync(c.d_status, 0, sizeof(int32_t), c.stream);
            c.err = "device-side overflow in octree distribution (status " + std::to_string(*hStatus) + ")";
            return ORBB200_ERR_UNSUPPORTED;
        }
        for (int i = 0; i < n; i++) {
            n_out[i] = hCnt[i];
            if (n_out[i] > cap_per_img) { c.err = "extract: caller capacity too small"; return ORBB200_ERR_CAPACITY; }
            const size_t m = (size_t)std::max(std::min(n_out[i], take), 0);
            memcpy(kps + (size_t)i * cap_per_img, hK + (size_t)i * rowK, m * sizeof(orbb200_kp_t));
            memcpy(desc + (size_t)i * cap_per_img * 32, hD + (size_t)i * rowD, m * 32);
        }
        return ORBB200_OK;
    }
    ORBB200_CUDA_OK(c, cudaMemcpy2DAsync(kps, (size_t)cap_per_img * sizeof(orbb200_kp_t), c.d_kps, (size_t)kpi * sizeof(orbb200_kp_t),
                                         (size_t)take * sizeof(orbb200_kp_t), n, cudaMemcpyDeviceToHost, c.stream));
    ORBB200_CUDA_OK(c, cudaMemcpy2DAsync(desc, (size_t)cap_per_img * 32, c.d_desc, (size_t)kpi * 32, (size_t)take * 32, n,
                                         cudaMemcpyDeviceToHost, c.stream));
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(n_out, c.d_counts, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c.stream));
    int rc = check_status(c);   // synchronises
    if (rc != ORBB200_OK) return rc;
    for (int i = 0; i < n; i++)
        if (n_out[i] > cap_per_img) { c.err = "extract: caller capacity too small"; return ORBB200_ERR_CAPACITY; }
    return ORBB200_OK;
}

int orbb200_extract_batch(orbb200_ctx* ctx, const uint8_t* const* imgs, int n, int w, int h, size_t stride,
                          orbb200_kp_t* kps, uint8_t* desc, int cap_per_img, int* n_out)
{
    CTX_ENTER(ctx);
    if (!imgs || n <= 0 || n > c.maxBatch || w <= 0 || h <= 0 || stride < (size_t)w) { c.err = "extract: bad argument"; return ORBB200_ERR_ARG; }
    if (!c.hostCopies.empty()) { ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream)); deliver_host_copies(c); }
    const ShapeTables* st = get_shape(c, w, h);
    if (!st) return c.err.find("exceeds") != std::string::npos ? ORBB200_ERR_ARG : ORBB200_ERR_UNSUPPORTED;
    c.cur = st; c.curN = n;
    const Geom& g = st->g;
    // rows go straight into level 0 of the pyramid pool (row pitch conversion by the copy engine); small batches pass through the
    // pinned staging block (a host memcpy + a true DMA beats the copy engine's chunked bounce of pageable memory)
    for (int i = 0; i < n; i++)
        if (!imgs[i]) { c.err = "extract: null image"; return ORBB200_ERR_ARG; }
    const size_t imgB = (size_t)w * h;
    // measured per call: 752x480 0.241 -> 0.223 ms, 1241x376 0.370 -> 0.252 ms staged; 1920x1080 0.458 -> 0.497 ms (the host memcpy of
    // 2 MB costs more than the bounce saves): staged up to 1.5 MB per call
    const bool staged = c.stageUploads && imgB * n <= c.stageMaxBytes && ensure_scratch(c, 0, STAGE_LIMIT);
    // One or two staged images whose results fit the mirror half of the block: ONE graph launch does the upload (the first kernel
    // reads the pinned rows over PCIe), the extraction and the delivery (the last kernel also stores into the pinned mirror) --
    // no copy-engine operation on either side.  Measured per call (752x480 / 1241x376): 0.195 / 0.199 -> see DESIGN.md section 5.
    const int rowB = (w + 15) & ~15;
    const int kpi = g.kpPerImg;
    const size_t cntB = align_up(sizeof(int32_t) * (size_t)n, 64);
    if (staged && c.hostGraph && n <= 2 && (size_t)rowB * h * n <= STAGE_D2H_OFF && 64 + cntB + (size_t)n * kpi * 60 <= STAGE_LIMIT - STAGE_D2H_OFF && kps &&
        desc && n_out && cap_per_img > 0) {
        HostStage hs;
        hs.imgs = c.h_scratch + STAGE_H2D_OFF; hs.imgBytes = (size_t)rowB * h; hs.rowBytes = rowB;
        uint8_t* hm = c.h_scratch + STAGE_D2H_OFF;
        hs.mirror.status = reinterpret_cast<int32_t*>(hm);
        hs.mirror.counts = reinterpret_cast<int32_t*>(hm + 64);
        hs.mirror.kps = reinterpret_cast<orbb200_kp_t*>(hm + 64 + cntB);
        hs.mirror.desc = hm + 64 + cntB + (size_t)n * kpi * sizeof(orbb200_kp_t);
        hs.mirror.d_status = c.d_status;
        hs.hostPyr = nullptr;
        if (c.mirrorPyramid) {                      // the drop-in's mvImagePyramid: image 0's levels straight into the pinned mirror
            if (c.h_pyrMirrorBytes < c.gmax.pyrBytes) {
                c.allocEpoch++;
                ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
                if (c.h_pyrMirror) cudaFreeHost(c.h_pyrMirror);
                c.h_pyrMirror = nullptr; c.h_pyrMirrorBytes = 0;
                ORBB200_CUDA_OK(c, cudaMallocHost((void**)&c.h_pyrMirror, c.gmax.pyrBytes));
                c.h_pyrMirrorBytes = c.gmax.pyrBytes;
                for (auto& kv : c.shapes)           // recordings that point at the old mirror
                    for (auto it = kv.second.graphs.begin(); it != kv.second.graphs.end();)
                        if (it->first >= 65536 + 131072) { cudaGraphExecDestroy(it->second.exec); it = kv.second.graphs.erase(it); } else ++it;
            }
            hs.hostPyr = c.h_pyrMirror;
        }
        for (int i = 0; i < n; i++) {
            uint8_t* dst = c.h_scratch + STAGE_H2D_OFF + (size_t)i * hs.imgBytes;
            if (stride == (size_t)w && rowB == w) memcpy(dst, imgs[i], imgB);
            else for (int y = 0; y < h; y++) memcpy(dst + (size_t)y * rowB, imgs[i] + (size_t)y * stride, (size_t)w);
        }
        int rc = run_extract(c, n, &hs);
        if (rc != ORBB200_OK) return rc;
        ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
        if (hs.hostPyr) c.pyrMirrorFresh = c.cur;
        if (*hs.mirror.status != 0) {
            cudaMemsetAsync(c.d_status, 0, sizeof(int32_t), c.stream);
            c.err = "device-side overflow in octree distribution (status " + std::to_string(*hs.mirror.status) + ")";
            return ORBB200_ERR_UNSUPPORTED;
        }
        for (int i = 0; i < n; i++) {
            n_out[i] = hs.mirror.counts[i];
            if (n_out[i] > cap_per_img) { c.err = "extract: caller capacity too small"; return ORBB200_ERR_CAPACITY; }
            const size_t m = (size_t)std::max(n_out[i], 0);
            memcpy(kps + (size_t)i * cap_per_img, hs.mirror.kps + (size_t)i * kpi, m * sizeof(orbb200_kp_t));
            memcpy(desc + (size_t)i * cap_per_img * 32, hs.mirror.desc + (size_t)i * kpi * 32, m * 32);
        }
        return ORBB200_OK;
    }
    for (int i = 0; i < n; i++) {
        const uint8_t* src = imgs[i];
        size_t srcStride = stride;
        if (staged) {
            uint8_t* hs = c.h_scratch + STAGE_H2D_OFF + (size_t)i * imgB;
            if (stride == (size_t)w) memcpy(hs, imgs[i], imgB);
            else for (int y = 0; y < h; y++) memcpy(hs + (size_t)y * w, imgs[i] + (size_t)y * stride, (size_t)w);
            src = hs; srcStride = (size_t)w;
        }
        ORBB200_CUDA_OK(c, cudaMemcpy2DAsync(c.d_pyr + (size_t)i * g.pyrBytes + g.lv[0].off, g.lv[0].pitch, src, srcStride, (size_t)w, (size_t)h,
                                             cudaMemcpyHostToDevice, c.stream));
    }
    int rc = run_extract(c, n);
    if (rc != ORBB200_OK) return rc;
    return orbb200_download_results(ctx, n, kps, desc, cap_per_img, n_out);
}

int orbb200_extract(orbb200_ctx* ctx, const uint8_t* img, int w, int h, size_t stride,
                    orbb200_kp_t* kps, uint8_t* desc, int cap, int* n_out)
{
    if (ctx && (!img || w <= 0 || h <= 0)) {   // ORBextractor::operator(): empty image -> return (:1046-1047)
        if (n_out) *n_out = 0;
        return ORBB200_OK;
    }
    const uint8_t* imgs[1] = {img};
    return orbb200_extract_batch(ctx, imgs, 1, w, h, stride, kps, desc, cap, n_out);
}

int orbb200_results_device(orbb200_ctx* ctx, const orbb200_kp_t** d_kps, const uint8_t** d_desc, const int32_t** d_counts, int* cap_per_img)
{
    if (!ctx || !ctx->c.cur) return ORBB200_ERR_ARG;
    Ctx& c = ctx->c;
    if (d_kps) *d_kps = c.d_kps;
    if (d_desc) *d_desc = c.d_desc;
    if (d_counts) *d_counts = c.d_counts;
    if (cap_per_img) *cap_per_img = c.cur->g.kpPerImg;
    return ORBB200_OK;
}

int orbb200_pyramid_level(orbb200_ctx* ctx, int img_index, int level, int blurred, uint8_t* dst, size_t dst_stride, int* w, int* h)
{
    CTX_ENTER(ctx);
    if (!c.cur || img_index < 0 || img_index >= c.curN || level < 0 || level >= c.nlevels) { c.err = "pyramid_level: bad argument"; return ORBB200_ERR_ARG; }
    const Geom& g = c.cur->g;
    const LevelGeom& L = g.lv[level];
    if (w) *w = L.w;
    if (h) *h = L.h;
    if (!dst) return ORBB200_OK;
    if (dst_stride < (size_t)L.w) { c.err = "pyramid_level: stride too small"; return ORBB200_ERR_ARG; }
    if (L.w <= 0 || L.h <= 0) return ORBB200_OK;
    const uint8_t* src = (blurred ? c.d_blur : c.d_pyr) + (size_t)img_index * g.pyrBytes + L.off;
    ORBB200_CUDA_OK(c, cudaMemcpy2DAsync(dst, dst_stride, src, L.pitch, (size_t)L.w, (size_t)L.h, cudaMemcpyDeviceToHost, c.stream));
    ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
    return ORBB200_OK;
}

int orbb200_pyramid_mirror(orbb200_ctx* ctx, int img_index, int blurred, const uint8_t** level_ptr, size_t* level_pitch, int* level_w, int* level_h)
{
    CTX_ENTER(ctx);
    if (!c.cur || img_index < 0 || img_index >= c.curN || !level_ptr || !level_pitch) { c.err = "pyramid_mirror: bad argument"; return ORBB200_ERR_ARG; }
    const Geom& g = c.cur->g;
    if (c.h_pyrMirrorBytes < g.pyrBytes) {
        c.allocEpoch++;
        if (c.h_pyrMirror) cudaFreeHost(c.h_pyrMirror);
        c.h_pyrMirror = nullptr; c.h_pyrMirrorBytes = 0;
        c.pyrMirrorFresh = nullptr;
        ORBB200_CUDA_OK(c, cudaMallocHost((void**)&c.h_pyrMirror, c.gmax.pyrBytes));
        c.h_pyrMirrorBytes = c.gmax.pyrBytes;
    }
    if (!(c.pyrMirrorFresh == c.cur && img_index == 0 && !blurred)) {       // (else the extraction's kernels have already written it: orbb200_set_pyramid_mirror)
        const uint8_t* src = (blurred ? c.d_blur : c.d_pyr) + (size_t)img_index * g.pyrBytes;
        ORBB200_CUDA_OK(c, cudaMemcpyAsync(c.h_pyrMirror, src, g.pyrBytes, cudaMemcpyDeviceToHost, c.stream));
        ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
        c.pyrMirrorFresh = nullptr;
    }
    for (int l = 0; l < c.nlevels; l++) {
        const LevelGeom& L = g.lv[l];
        level_ptr[l] = (L.w > 0 && L.h > 0) ? c.h_pyrMirror + L.off : nullptr;
        level_pitch[l] = (size_t)L.pitch;
        if (level_w) level_w[l] = L.w;
        if (level_h) level_h[l] = L.h;
    }
    return ORBB200_OK;
}

int orbb200_set_pyramid_mirror(orbb200_ctx* ctx, int enable)
{
    CTX_ENTER(ctx);
    c.mirrorPyramid = enable != 0;
    return ORBB200_OK;
}

int orbb200_level_candidates(orbb200_ctx* ctx, int img_index, int level, int32_t* xyr, int cap)
{
    CTX_ENTER(ctx);
    if (!c.cur || img_index < 0 || img_index >= c.curN || level < 0 || level >= c.nlevels) { c.err = "level_candidates: bad argument"; return ORBB200_ERR_ARG; }
    const Geom& g = c.cur->g;
    const LevelGeom& L = g.lv[level];
    int32_t cnt = 0;
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(&cnt, c.d_candCount + img_index * MAX_LEVELS + level, sizeof(cnt), cudaMemcpyDeviceToHost, c.stream));
    ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
    const int take = std::min(std::min(cnt, L.candCap), cap);
    if (take > 0 && xyr) {
        std::vector<uint32_t> tmp(take);
        ORBB200_CUDA_OK(c, cudaMemcpyAsync(tmp.data(), c.d_cand + (size_t)img_index * g.candPerImg + L.candOff, sizeof(uint32_t) * take,
                                           cudaMemcpyDeviceToHost, c.stream));
        ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
        for (int i = 0; i < take; i++) { xyr[3 * i] = tmp[i] & 0xfff; xyr[3 * i + 1] = (tmp[i] >> 12) & 0xfff; xyr[3 * i + 2] = tmp[i] >> 24; }
    }
    return cnt;
}

// ---- brute-force Hamming ----------------------------------------------------------------------------
static int knn2_splits(int nq, int nm)
{
    // enough CTAs for ~2 waves of 148 SMs x 4 resident CTAs, at least 64 map descriptors per split
    const int qtiles = (nq + 128 * knn2_queries_per_thread(nq, nm) - 1) / (128 * knn2_queries_per_thread(nq, nm));
    int s = std::max(1, (148 * 8) / std::max(qtiles, 1));
    s = std::min(s, std::max(1, nm / 64));
    return s;
}

int orbb200_hamming_knn2_device(orbb200_ctx* ctx, const uint8_t* d_q, int nq, const uint8_t* d_m, int nm,
                                int32_t* d_best_idx, int32_t* d_best_d, int32_t* d_second_d)
{
    CTX_ENTER(ctx);
    if (nq <= 0) return ORBB200_OK;
    if (!d_q || (!d_m && nm > 0) || nm < 0 || !d_best_idx || !d_best_d || !d_second_d) { c.err = "knn2: bad argument"; return ORBB200_ERR_ARG; }
    const int ns = knn2_splits(nq, nm);
    static_assert(sizeof(int4) == 16, "");
    // partials live at the end of the scratch arena (host wrapper uses the front)
    const size_t need = (size_t)ns * nq * sizeof(int4);
    if (!ensure_scratch(c, align_up(need, 256) + 256, 0)) return ORBB200_ERR_CUDA;
    int4* part = reinterpret_cast<int4*>(c.d_scratch + c.d_scratch_bytes - align_up(need, 256));
    launch_knn2(c, d_q, nq, d_m, nm, ns, part, d_best_idx, d_best_d, d_second_d);
    ORBB200_CUDA_OK(c, cudaGetLastError());
    return ORBB200_OK;
}

int orbb200_hamming_knn2(orbb200_ctx* ctx, const uint8_t* q, int nq, const uint8_t* m, int nm,
                         int32_t* best_idx, int32_t* best_d, int32_t* second_d)
{
    CTX_ENTER(ctx);
    if (nq <= 0) return ORBB200_OK;
    if (!q || (!m && nm > 0) || nm < 0 || !best_idx || !best_d || !second_d) { c.err = "knn2: bad argument"; return ORBB200_ERR_ARG; }
    const int ns = knn2_splits(nq, nm);
    const size_t qB = align_up((size_t)nq * 32, 256), mB = align_up((size_t)std::max(nm, 1) * 32, 256), oB = align_up((size_t)nq * 4, 256);
    const size_t need = qB + mB + 3 * oB + align_up((size_t)ns * nq * sizeof(int4), 256) + 4096;
    if (!ensure_scratch(c, need, 0)) return ORBB200_ERR_CUDA;
    uint8_t* dq = c.d_scratch;
    uint8_t* dm = dq + qB;
    int32_t* bi = reinterpret_cast<int32_t*>(dm + mB);
    int32_t* bd = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(bi) + oB);
    int32_t* sd = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(bd) + oB);
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(dq, q, (size_t)nq * 32, cudaMemcpyHostToDevice, c.stream));
    if (nm > 0) ORBB200_CUDA_OK(c, cudaMemcpyAsync(dm, m, (size_t)nm * 32, cudaMemcpyHostToDevice, c.stream));
    int rc = orbb200_hamming_knn2_device(ctx, dq, nq, dm, nm, bi, bd, sd);
    if (rc != ORBB200_OK) return rc;
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(best_idx, bi, (size_t)nq * 4, cudaMemcpyDeviceToHost, c.stream));
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(best_d, bd, (size_t)nq * 4, cudaMemcpyDeviceToHost, c.stream));
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(second_d, sd, (size_t)nq * 4, cudaMemcpyDeviceToHost, c.stream));
    ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
    return ORBB200_OK;
}

int orbb200_distinctive_descriptors(orbb200_ctx* ctx, const uint8_t* desc, const int32_t* group_ptr, int n_groups,
                                    int32_t* best_idx, int32_t* best_median)
{
    CTX_ENTER(ctx);
    if (n_groups <= 0) return ORBB200_OK;
    if (!group_ptr || !best_idx || !best_median) { c.err = "distinctive_descriptors: bad argument"; return ORBB200_ERR_ARG; }
    const int total = group_ptr[n_groups];
    if (group_ptr[0] != 0 || total < 0 || (total > 0 && !desc)) { c.err = "distinctive_descriptors: bad group_ptr"; return ORBB200_ERR_ARG; }
    for (int g = 0; g < n_groups; g++) {
        const int n = group_ptr[g + 1] - group_ptr[g];
        if (n < 0 || n >= (1 << 20)) { c.err = "distinctive_descriptors: group sizes must be in [0, 2^20)"; return ORBB200_ERR_ARG; }
    }
    const size_t dB = align_up((size_t)std::max(total, 1) * 32, 256), pB = align_up((size_t)(n_groups + 1) * 4, 256), oB = align_up((size_t)n_groups * 4, 256);
    if (!ensure_scratch(c, dB + pB + 2 * oB + 4096, 0)) return ORBB200_ERR_CUDA;
    uint8_t* dDesc = c.d_scratch;
    int32_t* dPtr = reinterpret_cast<int32_t*>(dDesc + dB);
    int32_t* dBi = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(dPtr) + pB);
    int32_t* dBm = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(dBi) + oB);
    if (total > 0) ORBB200_CUDA_OK(c, cudaMemcpyAsync(dDesc, desc, (size_t)total * 32, cudaMemcpyHostToDevice, c.stream));
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(dPtr, group_ptr, (size_t)(n_groups + 1) * 4, cudaMemcpyHostToDevice, c.stream));
    launch_distinctive(c, dDesc, dPtr, n_groups, dBi, dBm);
    ORBB200_CUDA_OK(c, cudaGetLastError());
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(best_idx, dBi, (size_t)n_groups * 4, cudaMemcpyDeviceToHost, c.stream));
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(best_median, dBm, (size_t)n_groups * 4, cudaMemcpyDeviceToHost, c.stream));
    ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
    return ORBB200_OK;
}

// POPC issue peak in G popc/s measured on this device (roofline denominator for matching)
double orbb200_measure_popc_peak(orbb200_ctx* ctx)
{
    if (!ctx) return -1.0;
    Ctx& c = ctx->c;
    if (cudaSetDevice(c.device) != cudaSuccess) return -1.0;
    if (!ensure_scratch(c, 4096, 0)) return -1.0;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, c.device);
    const int blocks = prop.multiProcessorCount * 8, iters = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch_popc_peak(c, reinterpret_cast<uint32_t*>(c.d_scratch), blocks, 100);
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0, c.stream);
        launch_popc_peak(c, reinterpret_cast<uint32_t*>(c.d_scratch), blocks, iters);
        cudaEventRecord(e1, c.stream);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    const double popc = (double)blocks * 256 * iters * 32;
    return popc / (best * 1e-3) * 1e-9;
}

// ---- frames -----------------------------------------------------------------------------------------
static int frame_finish(Ctx& c, orbb200_frame* f)
{
    ORBB200_CUDA_OK(c, cudaMalloc(&f->owned[3], sizeof(int32_t) * (GRID_CELLS + 1)));
    ORBB200_CUDA_OK(c, cudaMalloc(&f->owned[4], sizeof(int32_t) * std::max(f->cap, 1)));
    f->h.cellStart = (int32_t*)f->owned[3];
    f->h.cellItems = (int32_t*)f->owned[4];
    ORBB200_CUDA_OK(c, cudaMalloc(&f->owned[5], sizeof(int4) * std::max(f->cap, 1)));
    f->h.cellKp = (int4*)f->owned[5];
    ORBB200_CUDA_OK(c, cudaMalloc((void**)&f->d_self, sizeof(FrameDev)));
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(f->d_self, &f->h, sizeof(FrameDev), cudaMemcpyHostToDevice, c.stream));
    launch_grid_build(c, f->d_self, 1);
    ORBB200_CUDA_OK(c, cudaGetLastError());
    ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));   // f->h is host memory owned by f: safe, but keep semantics simple
    return ORBB200_OK;
}

int orbb200_frame_upload(orbb200_ctx* ctx, orbb200_frame** out, const orbb200_kp_t* kps, const uint8_t* desc,
                         const float* u_right, int n, float min_x, float min_y, float inv_w, float inv_h)
{
    CTX_ENTER(ctx);
    if (!out || n < 0 || (n > 0 && (!kps || !desc))) { c.err = "frame_upload: bad argument"; return ORBB200_ERR_ARG; }
    orbb200_frame* f = new orbb200_frame;
    f->ctx = &c; f->cap = n;
    auto bail = [&](int rc) { orbb200_frame_free(f); return rc; };
    if (cudaMalloc(&f->owned[0], sizeof(orbb200_kp_t) * std::max(n, 1)) != cudaSuccess || cudaMalloc(&f->owned[1], 32 * (size_t)std::max(n, 1)) != cudaSuccess) {
        c.err = "frame_upload: cudaMalloc failed"; return bail(ORBB200_ERR_CUDA);
    }
    if (n > 0) {
        cudaMemcpyAsync(f->owned[0], kps, sizeof(orbb200_kp_t) * n, cudaMemcpyHostToDevice, c.stream);
        cudaMemcpyAsync(f->owned[1], desc, 32 * (size_t)n, cudaMemcpyHostToDevice, c.stream);
    }
    if (u_right && n > 0) {
        if (cudaMalloc(&f->owned[2], sizeof(float) * n) != cudaSuccess) { c.err = "frame_upload: cudaMalloc failed"; return bail(ORBB200_ERR_CUDA); }
        cudaMemcpyAsync(f->owned[2], u_right, sizeof(float) * n, cudaMemcpyHostToDevice, c.stream);
    }
    f->h.kps = (const orbb200_kp_t*)f->owned[0]; f->h.desc = (const uint8_t*)f->owned[1]; f->h.uRight = (const float*)f->owned[2];
    f->h.n_ptr = nullptr; f->h.n = n;
    f->h.minX = min_x; f->h.minY = min_y; f->h.invW = inv_w; f->h.invH = inv_h;
    int rc = frame_finish(c, f);
    if (rc != ORBB200_OK) return bail(rc);
    *out = f;
    return ORBB200_OK;
}

int orbb200_frame_from_extract(orbb200_ctx* ctx, orbb200_frame** out, int img_index, float min_x, float min_y, float inv_w, float inv_h)
{
    CTX_ENTER(ctx);
    if (!out || !c.cur || img_index < 0 || img_index >= c.curN) { c.err = "frame_from_extract: bad argument"; return ORBB200_ERR_ARG; }
    const int kpi = c.cur->g.kpPerImg;
    orbb200_frame* f = new orbb200_frame;
    f->ctx = &c; f->cap = kpi;
    f->h.kps = c.d_kps + (size_t)img_index * kpi;
    f->h.desc = c.d_desc + (size_t)img_index * kpi * 32;
    f->h.uRight = nullptr;
    f->h.n_ptr = c.d_counts + img_index; f->h.n = kpi;
    f->h.minX = min_x; f->h.minY = min_y; f->h.invW = inv_w; f->h.invH = inv_h;
    int rc = frame_finish(c, f);
    if (rc != ORBB200_OK) { orbb200_frame_free(f); return rc; }
    *out = f;
    return ORBB200_OK;
}

int orbb200_frame_from_extract_stereo(orbb200_ctx* ctx, orbb200_frame** out, int img_left, float min_x, float min_y, float inv_w, float inv_h)
{
    CTX_ENTER(ctx);
    if (!c.stereoValid) { c.err = "frame_from_extract_stereo: no stereo matching since the last extraction"; return ORBB200_ERR_ARG; }
    int rc = orbb200_frame_from_extract(ctx, out, img_left, min_x, min_y, inv_w, inv_h);
    if (rc != ORBB200_OK) return rc;
    (*out)->h.uRight = c.d_uRight + (size_t)img_left * c.cur->g.kpPerImg;
    ORBB200_CUDA_OK(c, cudaMemcpyAsync((*out)->d_self, &(*out)->h, sizeof(FrameDev), cudaMemcpyHostToDevice, c.stream));
    ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
    return ORBB200_OK;
}

void orbb200_frame_free(orbb200_frame* f)
{
    if (!f) return;
    if (f->ctx) { cudaSetDevice(f->ctx->device); cudaStreamSynchronize(f->ctx->stream); }
    for (void* p : f->owned) if (p) cudaFree(p);
    if (f->d_self) cudaFree(f->d_self);
    delete f;
}

int orbb200_frame_features_in_area(orbb200_ctx* ctx, const orbb200_frame* f, float x, float y, float r,
                                   int min_level, int max_level, int32_t* out, int cap)
{
    CTX_ENTER(ctx);
    if (!f || cap < 0) { c.err = "features_in_area: bad argument"; return ORBB200_ERR_ARG; }
    if (!ensure_scratch(c, sizeof(int32_t) * (size_t)(cap + 64), 0)) return ORBB200_ERR_CUDA;
    int32_t* d_cnt = reinterpret_cast<int32_t*>(c.d_scratch);
    int32_t* d_out = d_cnt + 64;
    launch_features_in_area(c, f->d_self, x, y, r, min_level, max_level, d_out, cap, d_cnt);
    int32_t cnt = 0;
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(&cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, c.stream));
    ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
    if (out && std::min(cnt, cap) > 0) {
        ORBB200_CUDA_OK(c, cudaMemcpyAsync(out, d_out, sizeof(int32_t) * std::min(cnt, cap), cudaMemcpyDeviceToHost, c.stream));
        ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
    }
    return cnt;
}

}  // extern "C"

// ---- windowed searches: host wrappers ------------------------------------------------------------------
namespace {

struct QueryHost {
    int nq = 0;
    const uint8_t* valid = nullptr; const float* x = nullptr; const float* y = nullptr; const float* aux = nullptr;
    const int32_t* level = nullptr; const float* viewcos = nullptr; const float* angle = nullptr;
    const uint8_t* desc = nullptr; const uint8_t* obs_pos = nullptr; const uint8_t* kp_blocked = nullptr;
    const float* r = nullptr; const int32_t* maxlevel = nullptr; const int32_t* cand_idx = nullptr; int n_cand = 0;
    const float* inv_sigma2 = nullptr; int acc_th = 0, flags = 0;
    bool dev_queries = false;      // valid/x/y/aux/level/viewcos/desc already are device arrays (no upload)
};

// Upload one job's queries, run it, return device pointers of the outputs inside the scratch arena.
int run_window_job(Ctx& c, const orbb200_frame* F, const QueryHost& Q, int mode, int levelMode, int checkOri,
                   float th, float nnratio, float mbf,
                   int32_t* h_best_idx, int32_t* h_best_dist, int32_t* h_per_kp, int32_t* h_per_query, int* h_nmatches)
{
    const int nq = Q.nq, kpCap = std::max(F->cap, 1);
    const size_t need = (size_t)nq * (32 + 4 * 8 + 2 + 4 * 3) + (size_t)kpCap * (1 + 4) + (size_t)Q.n_cand * 4 + win_scratch_ints(kpCap, nq) * 4 +
                        80 * 256 + sizeof(WinJob);
    if (!c.hostCopies.empty()) { ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream)); deliver_host_copies(c); }   // staged results of an earlier frame step
    if (!ensure_scratch(c, need, 0)) return ORBB200_ERR_CUDA;
    Arena A(c.d_scratch, c.d_scratch_bytes);
    // Every host array of the call is first allocated back to back at the head of the device arena and mirrored at the same offset
    // of the pinned staging block: ONE copy uploads them all (descriptor, the query arrays, the per-keypoint flags), and the
    // outputs -- also back to back -- come home in ONE copy.  A matcher call used to be a dozen pageable-memory copies, each a
    // synchronous bounce through the driver, around 70 us of kernels.
    const size_t upBound = (size_t)nq * (32 + 4 * 8 + 2) + (size_t)kpCap + (size_t)Q.n_cand * 4 + 24 * 256 + sizeof(WinJob);
    const size_t outBound = (size_t)nq * 12 + (size_t)kpCap * 4 + 6 * 256;
    const bool staged = c.stageMatch && upBound <= STAGE_D2H_OFF - STAGE_H2D_OFF && outBound <= STAGE_LIMIT - STAGE_D2H_OFF && ensure_scratch(c, 0, STAGE_LIMIT);
    uint8_t* const hUp = staged ? c.h_scratch + STAGE_H2D_OFF : nullptr;
    auto up = [&](const void* src, size_t bytes) -> void* {
        uint8_t* d = A.take<uint8_t>(bytes);
        if (staged) memcpy(hUp + (d - c.d_scratch), src, bytes);
        else cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, c.stream);
        return d;
    };
    WinJob J{};
    auto upF = [&](const float* src) -> const float* { return (!src || Q.dev_queries) ? src : static_cast<const float*>(up(src, sizeof(float) * nq)); };
    auto upB = [&](const uint8_t* src, size_t n) -> const uint8_t* { return src ? static_cast<const uint8_t*>(up(src, n)) : nullptr; };
    auto upI = [&](const int32_t* src, size_t n) -> const int32_t* { return src ? static_cast<const int32_t*>(up(src, 4 * n)) : nullptr; };
    J.frame = F->d_self; J.nq = nq; J.mode = mode; J.levelMode = levelMode; J.checkOri = checkOri; J.kpCap = kpCap;
    J.th = th; J.nnratio = nnratio; J.mbf = mbf;
    WinJob* dJ = A.take<WinJob>(1);                    // filled last (it holds the output pointers), uploaded with the rest
    float sf[MAX_LEVELS] = {};
    std::copy(c.scale.begin(), c.scale.begin() + c.nlevels, sf);
    J.scaleFactors = static_cast<const float*>(up(sf, sizeof(sf))); J.nLevels = c.nlevels;
    J.q_valid = Q.dev_queries ? Q.valid : upB(Q.valid, nq); J.q_x = upF(Q.x); J.q_y = upF(Q.y); J.q_aux = upF(Q.aux);
    J.q_level = (Q.level && Q.dev_queries) ? Q.level : upI(Q.level, nq);
    J.q_viewcos = upF(Q.viewcos); J.q_angle = upF(Q.angle); J.q_r = upF(Q.r);
    J.q_maxlevel = upI(Q.maxlevel, nq);
    if (Q.cand_idx && Q.n_cand > 0) J.cand_idx = upI(Q.cand_idx, Q.n_cand);
    if (Q.inv_sigma2) {
        float is2[MAX_LEVELS] = {};
        std::copy(Q.inv_sigma2, Q.inv_sigma2 + c.nlevels, is2);
        J.invLevelSigma2 = static_cast<const float*>(up(is2, sizeof(is2)));
    }
    J.accTh = Q.acc_th; J.flags = Q.flags;
    J.q_desc = Q.dev_queries ? Q.desc : upB(Q.desc, (size_t)nq * 32); J.q_obs_pos = upB(Q.obs_pos, nq); J.kp_blocked = upB(Q.kp_blocked, F->cap);
    const size_t upEnd = A.off;
    J.out_best_idx = A.take<int32_t>(nq);              // outputs back to back: one download
    const size_t outBegin = reinterpret_cast<uint8_t*>(J.out_best_idx) - c.d_scratch;
    J.out_best_dist = A.take<int32_t>(nq);
    J.out_per_kp = A.take<int32_t>(kpCap); J.out_per_query = A.take<int32_t>(nq);
    J.out_nmatches = A.take<int32_t>(1);
    const size_t outEnd = A.off;
    J.scratch = A.take<int>(win_scratch_ints(kpCap, nq));
    if (staged) {
        const size_t jOff = reinterpret_cast<uint8_t*>(dJ) - c.d_scratch;
        memcpy(hUp + jOff, &J, sizeof(J));
        ORBB200_CUDA_OK(c, cudaMemcpyAsync(c.d_scratch + jOff, hUp + jOff, upEnd - jOff, cudaMemcpyHostToDevice, c.stream));
    } else {
        ORBB200_CUDA_OK(c, cudaMemcpyAsync(dJ, &J, sizeof(J), cudaMemcpyHostToDevice, c.stream));
    }
    launch_window_match(c, dJ, 1, nq, kpCap);
    ORBB200_CUDA_OK(c, cudaGetLastError());
    int32_t nm = 0;
    if (staged) {
        uint8_t* hOut = c.h_scratch + STAGE_D2H_OFF;
        ORBB200_CUDA_OK(c, cudaMemcpyAsync(hOut, c.d_scratch + outBegin, outEnd - outBegin, cudaMemcpyDeviceToHost, c.stream));
        ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
        auto home = [&](void* dst, const void* dsrc, size_t bytes) { if (dst && bytes) memcpy(dst, hOut + (static_cast<const uint8_t*>(dsrc) - (c.d_scratch + outBegin)), bytes); };
        home(h_best_idx, J.out_best_idx, 4 * (size_t)nq); home(h_best_dist, J.out_best_dist, 4 * (size_t)nq);
        home(h_per_kp, J.out_per_kp, 4 * (size_t)std::max(F->cap, 0)); home(h_per_query, J.out_per_query, 4 * (size_t)nq);
        home(&nm, J.out_nmatches, 4);
    } else {
        if (h_best_idx) cudaMemcpyAsync(h_best_idx, J.out_best_idx, 4 * (size_t)nq, cudaMemcpyDeviceToHost, c.stream);
        if (h_best_dist) cudaMemcpyAsync(h_best_dist, J.out_best_dist, 4 * (size_t)nq, cudaMemcpyDeviceToHost, c.stream);
        if (h_per_kp && F->cap > 0) cudaMemcpyAsync(h_per_kp, J.out_per_kp, 4 * (size_t)F->cap, cudaMemcpyDeviceToHost, c.stream);
        if (h_per_query) cudaMemcpyAsync(h_per_query, J.out_per_query, 4 * (size_t)nq, cudaMemcpyDeviceToHost, c.stream);
        cudaMemcpyAsync(&nm, J.out_nmatches, 4, cudaMemcpyDeviceToHost, c.stream);
        ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
    }
    if (h_nmatches) *h_nmatches = nm;
    return ORBB200_OK;
}

}  // namespace

extern "C" {

int orbb200_search_by_projection(orbb200_ctx* ctx, const orbb200_frame* F, int nq,
                                 const uint8_t* q_valid, const float* q_u, const float* q_v, const float* q_uR,
                                 const int32_t* q_level, const float* q_viewcos, const uint8_t* q_desc,
                                 const uint8_t* q_obs_pos, const uint8_t* kp_blocked, float th, float nnratio,
                                 int32_t* out_best_idx, int32_t* out_best_dist, int32_t* out_query_of_kp, int* nmatches)
{
    CTX_ENTER(ctx);
    if (!F || nq < 0 || (nq > 0 && (!q_u || !q_v || !q_uR || !q_level || !q_viewcos || !q_desc))) { c.err = "search_by_projection: bad argument"; return ORBB200_ERR_ARG; }
    QueryHost Q; Q.nq = nq; Q.valid = q_valid; Q.x = q_u; Q.y = q_v; Q.aux = q_uR; Q.level = q_level; Q.viewcos = q_viewcos;
    Q.desc = q_desc; Q.obs_pos = q_obs_pos; Q.kp_blocked = kp_blocked;
    return run_window_job(c, F, Q, WM_PROJ, 0, 0, th, nnratio, 0.f, out_best_idx, out_best_dist, out_query_of_kp, nullptr, nmatches);
}

int orbb200_map_upload(orbb200_ctx* ctx, orbb200_map** map, int n, const float* pos, const float* normal,
                       const float* max_distance, const float* min_distance, const uint8_t* desc)
{
    CTX_ENTER(ctx);
    if (!map || n < 0 || (n > 0 && (!pos || !normal || !max_distance || !min_distance || !desc))) { c.err = "map_upload: bad argument"; return ORBB200_ERR_ARG; }
    orbb200_map* m = new orbb200_map();
    m->ctx = &c; m->n = n;
    c.allocEpoch++;                                // a new map may reuse the address of a freed one
    const size_t k = (size_t)std::max(n, 1);
    bool ok = cudaMalloc((void**)&m->d_pos, k * 12) == cudaSuccess && cudaMalloc((void**)&m->d_normal, k * 12) == cudaSuccess &&
              cudaMalloc((void**)&m->d_maxDist, k * 4) == cudaSuccess && cudaMalloc((void**)&m->d_minDist, k * 4) == cudaSuccess &&
              cudaMalloc((void**)&m->d_desc, k * 32) == cudaSuccess && cudaMalloc((void**)&m->d_candidate, k) == cudaSuccess &&
              cudaMalloc((void**)&m->d_inView, k) == cudaSuccess && cudaMalloc((void**)&m->d_u, k * 4) == cudaSuccess &&
              cudaMalloc((void**)&m->d_v, k * 4) == cudaSuccess && cudaMalloc((void**)&m->d_uR, k * 4) == cudaSuccess &&
              cudaMalloc((void**)&m->d_viewcos, k * 4) == cudaSuccess && cudaMalloc((void**)&m->d_level, k * 4) == cudaSuccess &&
              cudaMalloc((void**)&m->d_count, 4) == cudaSuccess;
    if (!ok) { orbb200_map_free(m); cudaGetLastError(); c.err = "map_upload: cudaMalloc failed"; return ORBB200_ERR_CUDA; }
    if (n > 0) {
        cudaMemcpyAsync(m->d_pos, pos, (size_t)n * 12, cudaMemcpyHostToDevice, c.stream);
        cudaMemcpyAsync(m->d_normal, normal, (size_t)n * 12, cudaMemcpyHostToDevice, c.stream);
        cudaMemcpyAsync(m->d_maxDist, max_distance, (size_t)n * 4, cudaMemcpyHostToDevice, c.stream);
        cudaMemcpyAsync(m->d_minDist, min_distance, (size_t)n * 4, cudaMemcpyHostToDevice, c.stream);
        cudaMemcpyAsync(m->d_desc, desc, (size_t)n * 32, cudaMemcpyHostToDevice, c.stream);
    }
    ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
    *map = m;
    return ORBB200_OK;
}

void orbb200_map_free(orbb200_map* m)
{
    if (!m) return;
    if (m->ctx) { cudaSetDevice(m->ctx->device); cudaStreamSynchronize(m->ctx->stream); }
    void* ptrs[] = {m->d_pos, m->d_normal, m->d_maxDist, m->d_minDist, m->d_desc, m->d_candidate, m->d_inView, m->d_u, m->d_v,
                    m->d_uR, m->d_viewcos, m->d_level, m->d_count};
    for (void* p : ptrs) if (p) cudaFree(p);
    delete m;
}

// enqueue the projection of the whole map; results stay in the map's device arrays
static int enqueue_frustum(Ctx& c, const orbb200_map* m, const orbb200_camera_pose* pose, float cosLimit, const uint8_t* candidate)
{
    if (candidate && m->n > 0) ORBB200_CUDA_OK(c, cudaMemcpyAsync(m->d_candidate, candidate, (size_t)m->n, cudaMemcpyHostToDevice, c.stream));
    FrustumJob J{};
    J.pose = *pose; J.cosLimit = cosLimit; J.n = m->n;
    J.pos = m->d_pos; J.normal = m->d_normal; J.maxDist = m->d_maxDist; J.minDist = m->d_minDist;
    J.candidate = candidate ? m->d_candidate : nullptr;
    J.inView = m->d_inView; J.u = m->d_u; J.v = m->d_v; J.uR = m->d_uR; J.level = m->d_level; J.viewcos = m->d_viewcos;
    J.count = m->d_count;
    launch_frustum(c, J);
    ORBB200_CUDA_OK(c, cudaGetLastError());
    return ORBB200_OK;
}

static void download_frustum(Ctx& c, const orbb200_map* m, uint8_t* in_view, float* u, float* v, float* uR, int32_t* level,
                             float* viewcos, int32_t* h_count)
{
    const size_t n = (size_t)m->n;
    if (n > 0) {
        if (in_view) cudaMemcpyAsync(in_view, m->d_inView, n, cudaMemcpyDeviceToHost, c.stream);
        if (u) cudaMemcpyAsync(u, m->d_u, n * 4, cudaMemcpyDeviceToHost, c.stream);
        if (v) cudaMemcpyAsync(v, m->d_v, n * 4, cudaMemcpyDeviceToHost, c.stream);
        if (uR) cudaMemcpyAsync(uR, m->d_uR, n * 4, cudaMemcpyDeviceToHost, c.stream);
        if (level) cudaMemcpyAsync(level, m->d_level, n * 4, cudaMemcpyDeviceToHost, c.stream);
        if (viewcos) cudaMemcpyAsync(viewcos, m->d_viewcos, n * 4, cudaMemcpyDeviceToHost, c.stream);
    }
    cudaMemcpyAsync(h_count, m->d_count, 4, cudaMemcpyDeviceToHost, c.stream);
}

int orbb200_is_in_frustum(orbb200_ctx* ctx, const orbb200_map* map, const orbb200_camera_pose* pose, float viewing_cos_limit,
                          const uint8_t* candidate, uint8_t* out_in_view, float* out_u, float* out_v, float* out_uR,
                          int32_t* out_level, float* out_viewcos, int* n_in_view)
{
    CTX_ENTER(ctx);
    if (!map || !pose || map->ctx != &c) { c.err = "is_in_frustum: bad argument"; return ORBB200_ERR_ARG; }
    const int rc = enqueue_frustum(c, map, pose, viewing_cos_limit, candidate);
    if (rc != ORBB200_OK) return rc;
    int32_t cnt = 0;
    download_frustum(c, map, out_in_view, out_u, out_v, out_uR, out_level, out_viewcos, &cnt);
    ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
    if (n_in_view) *n_in_view = cnt;
    return ORBB200_OK;
}

int orbb200_search_local_points(orbb200_ctx* ctx, const orbb200_frame* F, const orbb200_map* map,
                                const orbb200_camera_pose* pose, float viewing_cos_limit, const uint8_t* candidate,
                                const uint8_t* obs_pos, const uint8_t* kp_blocked, float th, float nnratio,
                                uint8_t* out_in_view, float* out_u, float* out_v, float* out_uR, int32_t* out_level,
                                float* out_viewcos, int* n_in_view,
                                int32_t* out_best_idx, int32_t* out_best_dist, int32_t* out_query_of_kp, int* nmatches)
{
    CTX_ENTER(ctx);
    if (!F || !map || !pose || map->ctx != &c) { c.err = "search_local_points: bad argument"; return ORBB200_ERR_ARG; }
    int rc = enqueue_frustum(c, map, pose, viewing_cos_limit, candidate);
    if (rc != ORBB200_OK) return rc;
    int32_t cnt = 0;
    download_frustum(c, map, out_in_view, out_u, out_v, out_uR, out_level, out_viewcos, &cnt);
    // the projected map points are the queries of SearchByProjection: mbTrackInView -> valid
    QueryHost Q; Q.nq = map->n; Q.dev_queries = true;
    Q.valid = map->d_inView; Q.x = map->d_u; Q.y = map->d_v; Q.aux = map->d_uR; Q.level = map->d_level; Q.viewcos = map->d_viewcos;
    Q.desc = map->d_desc; Q.obs_pos = obs_pos; Q.kp_blocked = kp_blocked;
    rc = run_window_job(c, F, Q, WM_PROJ, 0, 0, th, nnratio, 0.f, out_best_idx, out_best_dist, out_query_of_kp, nullptr, nmatches);
    if (n_in_view) *n_in_view = cnt;      // run_window_job synchronised the stream
    return rc;
}

int orbb200_search_by_projection_frame(orbb200_ctx* ctx, const orbb200_frame* Cur, int nq,
                                       const uint8_t* q_valid, const float* q_u, const float* q_v, const float* q_invz,
                                       const int32_t* q_octave, const float* q_angle, const uint8_t* q_desc,
                                       const uint8_t* q_obs_pos, const uint8_t* kp_blocked,
                                       float th, float mbf, int mode, int check_ori, int32_t* out_query_of_kp, int* nmatches)
{
    CTX_ENTER(ctx);
    if (!Cur || nq < 0 || mode < 0 || mode > 2 || (nq > 0 && (!q_u || !q_v || !q_invz || !q_octave || !q_desc || (check_ori && !q_angle)))) {
        c.err = "search_by_projection_frame: bad argument"; return ORBB200_ERR_ARG;
    }
    QueryHost Q; Q.nq = nq; Q.valid = q_valid; Q.x = q_u; Q.y = q_v; Q.aux = q_invz; Q.level = q_octave; Q.angle = q_angle;
    Q.desc = q_desc; Q.obs_pos = q_obs_pos; Q.kp_blocked = kp_blocked;
    return run_window_job(c, Cur, Q, WM_PROJ_FRAME, mode, check_ori, th, 0.f, mbf, nullptr, nullptr, out_query_of_kp, nullptr, nmatches);
}

int orbb200_birdview_match(orbb200_ctx* ctx, const orbb200_kp_t* kps1, const uint8_t* desc1, int n1,
                           const orbb200_frame* F2, float* prev_xy, int window_size, float nnratio, int check_ori,
                           int32_t* matches12, int* nmatches)
{
    CTX_ENTER(ctx);
    if (!F2 || n1 < 0 || (n1 > 0 && (!kps1 || !desc1 || !matches12))) { c.err = "birdview_match: bad argument"; return ORBB200_ERR_ARG; }
    std::vector<float> x(n1), y(n1), ang(n1);
    std::vector<int32_t> lvl(n1);
    for (int i = 0; i < n1; i++) {
        x[i] = prev_xy ? prev_xy[2 * i] : kps1[i].x;
        y[i] = prev_xy ? prev_xy[2 * i + 1] : kps1[i].y;
        ang[i] = kps1[i].angle; lvl[i] = kps1[i].octave;
    }
    QueryHost Q; Q.nq = n1; Q.x = x.data(); Q.y = y.data(); Q.level = lvl.data(); Q.angle = ang.data(); Q.desc = desc1;
    int rc = run_window_job(c, F2, Q, WM_BIRD, prev_xy ? 1 : 0, check_ori, (float)window_size, nnratio, 0.f, nullptr, nullptr, nullptr, matches12, nmatches);
    if (rc != ORBB200_OK) return rc;
    if (prev_xy && n1 > 0) {
        // vPrevMatched[i1] = F2.mvKeysBird[vnMatches12[i1]].pt (:1777-1779)
        std::vector<orbb200_kp_t> k2(std::max(F2->cap, 1));
        ORBB200_CUDA_OK(c, cudaMemcpyAsync(k2.data(), F2->h.kps, sizeof(orbb200_kp_t) * F2->cap, cudaMemcpyDeviceToHost, c.stream));
        ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
        for (int i = 0; i < n1; i++)
            if (matches12[i] >= 0) { prev_xy[2 * i] = k2[matches12[i]].x; prev_xy[2 * i + 1] = k2[matches12[i]].y; }
    }
    return ORBB200_OK;
}

int orbb200_search_by_match_bird_kf(orbb200_ctx* ctx, const orbb200_kp_t* kf_kps, const uint8_t* has_mp,
                                    const uint8_t* mp_desc, int nk, const orbb200_frame* F, float r, float nnratio,
                                    int check_ori, int32_t* out_mp_of_kp, int* nmatches)
{
    CTX_ENTER(ctx);
    if (!F || nk < 0 || (nk > 0 && (!kf_kps || !has_mp || !mp_desc))) { c.err = "search_by_match_bird_kf: bad argument"; return ORBB200_ERR_ARG; }
    std::vector<float> x(nk), y(nk), ang(nk);
    for (int i = 0; i < nk; i++) { x[i] = kf_kps[i].x; y[i] = kf_kps[i].y; ang[i] = kf_kps[i].angle; }
    QueryHost Q; Q.nq = nk; Q.valid = has_mp; Q.x = x.data(); Q.y = y.data(); Q.angle = ang.data(); Q.desc = mp_desc;
    return run_window_job(c, F, Q, WM_BIRD_KF, 0, check_ori, r, nnratio, 0.f, nullptr, nullptr, out_mp_of_kp, nullptr, nmatches);
}

int orbb200_search_by_projection_bird(orbb200_ctx* ctx, const orbb200_frame* F, int nq, const uint8_t* q_valid,
                                      const float* q_x, const float* q_y, const uint8_t* q_desc,
                                      const uint8_t* q_obs_pos, const uint8_t* kp_blocked, float r, float nnratio,
                                      int32_t* out_query_of_kp, int* nmatches)
{
    CTX_ENTER(ctx);
    if (!F || nq < 0 || (nq > 0 && (!q_x || !q_y || !q_desc))) { c.err = "search_by_projection_bird: bad argument"; return ORBB200_ERR_ARG; }
    QueryHost Q; Q.nq = nq; Q.valid = q_valid; Q.x = q_x; Q.y = q_y; Q.desc = q_desc; Q.obs_pos = q_obs_pos; Q.kp_blocked = kp_blocked;
    return run_window_job(c, F, Q, WM_PROJ_BIRD, 0, 0, r, nnratio, 0.f, nullptr, nullptr, out_query_of_kp, nullptr, nmatches);
}

int orbb200_search_for_triangulation(orbb200_ctx* ctx,
                                     const orbb200_kp_t* kps1, const uint8_t* desc1, const float* uR1, const uint8_t* has_mp1, int n1,
                                     const orbb200_kp_t* kps2, const uint8_t* desc2, const float* uR2, const uint8_t* has_mp2, int n2,
                                     const int32_t* fv1_node, const int32_t* fv1_ptr, const int32_t* fv1_idx, int nn1,
                                     const int32_t* fv2_node, const int32_t* fv2_ptr, const int32_t* fv2_idx, int nn2,
                                     const float* F12, float ex, float ey,
                                     const float* scale_factors2, const float* level_sigma2_2,
                                     int only_stereo, int check_ori, int32_t* pairs, int* npairs)
{
    CTX_ENTER(ctx);
    if (n1 < 0 || n2 < 0 || !F12 || !scale_factors2 || !level_sigma2_2 || !pairs || !npairs ||
        (n1 > 0 && (!kps1 || !desc1 || !has_mp1)) || (n2 > 0 && (!kps2 || !desc2 || !has_mp2)) ||
        (nn1 > 0 && (!fv1_node || !fv1_ptr || !fv1_idx)) || (nn2 > 0 && (!fv2_node || !fv2_ptr || !fv2_idx))) {
        c.err = "search_for_triangulation: bad argument"; return ORBB200_ERR_ARG;
    }
    // shared vocabulary nodes: merge of the two ascending id lists (the f1it/f2it walk, :689-789)
    std::vector<int32_t> it1, ib2, ie2;
    int a = 0, b = 0;
    while (a < nn1 && b < nn2) {
        if (fv1_node[a] == fv2_node[b]) {
            for (int i = fv1_ptr[a]; i < fv1_ptr[a + 1]; i++) { it1.push_back(fv1_idx[i]); ib2.push_back(fv2_ptr[b]); ie2.push_back(fv2_ptr[b + 1]); }
            a++; b++;
        } else if (fv1_node[a] < fv2_node[b]) a++;
        else b++;
    }
    const int nItems = (int)it1.size();
    const int nidx2 = nn2 > 0 ? fv2_ptr[nn2] : 0;
    const size_t need = (size_t)(n1 + n2) * (28 + 32 + 4 + 1) + (size_t)nidx2 * 4 + (size_t)nItems * 12 + (size_t)n1 * 12 + 64 * 256;
    if (!ensure_scratch(c, need, 0)) return ORBB200_ERR_CUDA;
    Arena A(c.d_scratch, c.d_scratch_bytes);
    auto up = [&](const void* src, size_t bytes) -> void* {
        if (!src || bytes == 0) return nullptr;
        uint8_t* d = A.take<uint8_t>(bytes);
        cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, c.stream);
        return d;
    };
    TriJob J{};
    J.kps1 = (const orbb200_kp_t*)up(kps1, sizeof(orbb200_kp_t) * (size_t)n1); J.desc1 = (const uint8_t*)up(desc1, 32 * (size_t)n1);
    J.uR1 = (const float*)up(uR1, 4 * (size_t)n1); J.has_mp1 = (const uint8_t*)up(has_mp1, n1); J.n1 = n1;
    J.kps2 = (const orbb200_kp_t*)up(kps2, sizeof(orbb200_kp_t) * (size_t)n2); J.desc2 = (const uint8_t*)up(desc2, 32 * (size_t)n2);
    J.uR2 = (const float*)up(uR2, 4 * (size_t)n2); J.has_mp2 = (const uint8_t*)up(has_mp2, n2); J.n2 = n2;
    J.fv2_idx = (const int32_t*)up(fv2_idx, 4 * (size_t)nidx2);
    J.item_idx1 = (const int32_t*)up(it1.data(), 4 * (size_t)nItems); J.item_b2 = (const int32_t*)up(ib2.data(), 4 * (size_t)nItems);
    J.item_e2 = (const int32_t*)up(ie2.data(), 4 * (size_t)nItems); J.nItems = nItems;
    J.F12 = (const float*)up(F12, 36); J.ex = ex; J.ey = ey;
    J.scaleFactors2 = (const float*)up(scale_factors2, 4 * (size_t)c.nlevels); J.levelSigma2_2 = (const float*)up(level_sigma2_2, 4 * (size_t)c.nlevels);
    J.onlyStereo = only_stereo; J.checkOri = check_ori; J.nLevels2 = c.nlevels;
    J.match12 = A.take<int32_t>(std::max(n1, 1)); J.pairs = A.take<int32_t>(2 * (size_t)std::max(n1, 1)); J.npairs = A.take<int32_t>(1);
    launch_triangulation(c, J);
    ORBB200_CUDA_OK(c, cudaGetLastError());
    int32_t np = 0;
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(&np, J.npairs, 4, cudaMemcpyDeviceToHost, c.stream));
    ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));   // also keeps it1/ib2/ie2 alive until uploaded
    if (np > 0) {
        ORBB200_CUDA_OK(c, cudaMemcpyAsync(pairs, J.pairs, 8 * (size_t)np, cudaMemcpyDeviceToHost, c.stream));
        ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
    }
    *npairs = np;
    return ORBB200_OK;
}

int orbb200_search_for_initialization(orbb200_ctx* ctx, const orbb200_kp_t* kps1, const uint8_t* desc1, int n1,
                                      const orbb200_frame* F2, float* prev_xy, int window_size, float nnratio, int check_ori,
                                      int32_t* matches12, int* nmatches)
{
    // SearchForInitialization (:405-520) is the loop BirdviewMatch(F1,F2,..,vPrevMatched,..) (:1667-1786) was copied
    // from: octave-0 keypoints of F1, window around vbPrevMatched in F2's grid, same acceptance and stealing
    if (ctx && !prev_xy) { ctx->c.err = "search_for_initialization: prev_xy required"; return ORBB200_ERR_ARG; }
    return orbb200_birdview_match(ctx, kps1, desc1, n1, F2, prev_xy, window_size, nnratio, check_ori, matches12, nmatches);
}

int orbb200_search_window_best(orbb200_ctx* ctx, const orbb200_frame* F, int nq, const uint8_t* q_valid,
                               const float* q_x, const float* q_y, const float* q_r, const int32_t* q_min_level,
                               const int32_t* q_max_level, const uint8_t* q_desc, const float* q_aux, const float* q_angle,
                               const uint8_t* q_obs_pos, const uint8_t* kp_blocked, const float* inv_level_sigma2,
                               int acc_th, int flags, int32_t* out_best_idx, int32_t* out_best_dist,
                               int32_t* out_query_of_kp, int* nmatches)
{
    CTX_ENTER(ctx);
    const bool ori = (flags & ORBB200_WB_ORI) != 0;
    if (!F || nq < 0 || (nq > 0 && (!q_x || !q_y || !q_r || !q_min_level || !q_max_level || !q_desc)) ||
        ((flags & (ORBB200_WB_URCHECK | ORBB200_WB_CHI2)) && nq > 0 && !q_aux) || ((flags & ORBB200_WB_CHI2) && !inv_level_sigma2) ||
        (ori && nq > 0 && !q_angle)) {
        c.err = "search_window_best: bad argument"; return ORBB200_ERR_ARG;
    }
    QueryHost Q; Q.nq = nq; Q.valid = q_valid; Q.x = q_x; Q.y = q_y; Q.r = q_r; Q.level = q_min_level; Q.maxlevel = q_max_level;
    Q.desc = q_desc; Q.aux = q_aux; Q.angle = q_angle; Q.obs_pos = q_obs_pos; Q.kp_blocked = kp_blocked; Q.inv_sigma2 = inv_level_sigma2;
    Q.acc_th = acc_th;
    Q.flags = ((flags & ORBB200_WB_BLOCK) ? WF_BLOCK : 0) | ((flags & ORBB200_WB_URCHECK) ? WF_URCHECK : 0) | ((flags & ORBB200_WB_CHI2) ? WF_CHI2 : 0);
    return run_window_job(c, F, Q, WM_BEST, 0, ori ? 1 : 0, 0.f, 0.f, 0.f, out_best_idx, out_best_dist, out_query_of_kp, nullptr, nmatches);
}

int orbb200_search_by_bow(orbb200_ctx* ctx, const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int n1,
                          const orbb200_frame* F2, const uint8_t* valid2,
                          const int32_t* fv1_node, const int32_t* fv1_ptr, const int32_t* fv1_idx, int nn1,
                          const int32_t* fv2_node, const int32_t* fv2_ptr, const int32_t* fv2_idx, int nn2,
                          float nnratio, int check_ori, int kf_kf, int32_t* out, int* nmatches)
{
    CTX_ENTER(ctx);
    if (!F2 || n1 < 0 || !out || !nmatches || (n1 > 0 && (!desc1 || !valid1 || (check_ori && !angle1))) ||
        (nn1 > 0 && (!fv1_node || !fv1_ptr || !fv1_idx)) || (nn2 > 0 && (!fv2_node || !fv2_ptr || !fv2_idx))) {
        c.err = "search_by_bow: bad argument"; return ORBB200_ERR_ARG;
    }
    // queries in the reference's visiting order: shared nodes ascending, KF1 indices in node order (:178-255)
    std::vector<int32_t> qidx1, qb, qe;
    int a = 0, b = 0;
    while (a < nn1 && b < nn2) {
        if (fv1_node[a] == fv2_node[b]) {
            for (int i = fv1_ptr[a]; i < fv1_ptr[a + 1]; i++) {
                const int idx1 = fv1_idx[i];
                if (!valid1[idx1]) continue;                      // no MapPoint / bad MapPoint (:189-195)
                qidx1.push_back(idx1); qb.push_back(fv2_ptr[b]); qe.push_back(fv2_ptr[b + 1]);
            }
            a++; b++;
        } else if (fv1_node[a] < fv2_node[b]) a++;
        else b++;
    }
    const int nq = (int)qidx1.size();
    std::vector<uint8_t> qdesc((size_t)std::max(nq, 1) * 32);
    std::vector<float> qang(std::max(nq, 1), 0.f);
    for (int q = 0; q < nq; q++) { memcpy(&qdesc[(size_t)q * 32], desc1 + (size_t)qidx1[q] * 32, 32); if (angle1) qang[q] = angle1[qidx1[q]]; }
    QueryHost Q; Q.nq = nq; Q.level = qb.data(); Q.maxlevel = qe.data(); Q.desc = qdesc.data(); Q.angle = qang.data();
    Q.kp_blocked = kf_kf ? valid2 : nullptr;   // inverted below: kp_blocked means "not a candidate"
    std::vector<uint8_t> inv;
    if (kf_kf && valid2) { inv.resize(std::max(F2->cap, 1)); for (int i = 0; i < F2->cap; i++) inv[i] = !valid2[i]; Q.kp_blocked = inv.data(); }
    Q.cand_idx = fv2_idx; Q.n_cand = nn2 > 0 ? fv2_ptr[nn2] : 0;
    std::vector<int32_t> perKp(std::max(F2->cap, 1)), perQ(std::max(nq, 1));
    int nm = 0;
    int rc = ORBB200_OK;
    if (nq > 0) rc = run_window_job(c, F2, Q, kf_kf ? WM_BOW_KF_KF : WM_BOW_KF_F, 0, check_ori, 0.f, nnratio, 0.f, nullptr, nullptr, perKp.data(), perQ.data(), &nm);
    if (rc != ORBB200_OK) return rc;
    if (kf_kf) {
        for (int i = 0; i < n1; i++) out[i] = -1;                                   // vpMatches12 as KF2 keypoint indices
        for (int q = 0; q < nq; q++) if (perQ[q] >= 0) out[qidx1[q]] = perQ[q];
    } else {
        for (int i = 0; i < F2->cap; i++) out[i] = (nq > 0 && perKp[i] >= 0) ? qidx1[perKp[i]] : -1;   // vpMapPointMatches as KF keypoint indices
    }
    *nmatches = nm;
    return ORBB200_OK;
}

// ---- stereo matching ----------------------------------------------------------------
int orbb200_stereo_matches_device(orbb200_ctx* ctx, int n_frames, int left0, int right0, int stride_imgs, float mb, float mbf)
{
    CTX_ENTER(ctx);
    if (!c.cur || n_frames <= 0 || left0 < 0 || right0 < 0 || stride_imgs < 0 ||
        left0 + (n_frames - 1) * stride_imgs >= c.curN || right0 + (n_frames - 1) * stride_imgs >= c.curN || !(mb > 0.f) ||
        n_frames > c.maxBatch / 2 + 1) {
        c.err = "stereo_matches: bad argument (needs a previous extraction holding both images)"; return ORBB200_ERR_ARG;
    }
    launch_stereo(c, n_frames, left0, right0, stride_imgs, mb, mbf, c.d_invScale, c.d_nKept);
    ORBB200_CUDA_OK(c, cudaGetLastError());
    c.stereoValid = true;
    return ORBB200_OK;
}

int orbb200_stereo_results_device(orbb200_ctx* ctx, const float** d_uright, const float** d_depth, const int32_t** d_nkept)
{
    if (!ctx || !ctx->c.cur) return ORBB200_ERR_ARG;
    if (d_uright) *d_uright = ctx->c.d_uRight;
    if (d_depth) *d_depth = ctx->c.d_depth;
    if (d_nkept) *d_nkept = ctx->c.d_nKept;
    return ORBB200_OK;
}

int orbb200_compute_stereo_matches(orbb200_ctx* ctx, int img_left, int img_right, float mb, float mbf,
                                   float* u_right, float* depth, int cap, int* n_matches)
{
    CTX_ENTER(ctx);
    if (!u_right || !depth || cap < 0) { c.err = "compute_stereo_matches: bad argument"; return ORBB200_ERR_ARG; }
    int rc = orbb200_stereo_matches_device(ctx, 1, img_left, img_right, 0, mb, mbf);
    if (rc != ORBB200_OK) return rc;
    const int kpi = c.cur->g.kpPerImg, take = std::min(cap, kpi);
    int32_t nk = 0;
    if (take > 0) {
        ORBB200_CUDA_OK(c, cudaMemcpyAsync(u_right, c.d_uRight + (size_t)img_left * kpi, sizeof(float) * take, cudaMemcpyDeviceToHost, c.stream));
        ORBB200_CUDA_OK(c, cudaMemcpyAsync(depth, c.d_depth + (size_t)img_left * kpi, sizeof(float) * take, cudaMemcpyDeviceToHost, c.stream));
    }
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(&nk, c.d_nKept, sizeof(nk), cudaMemcpyDeviceToHost, c.stream));
    ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
    if (n_matches) *n_matches = nk;
    return ORBB200_OK;
}

// ---- batched C2 step ----------------------------------------------------------------------------------
int orbb200_stereo_step_device(orbb200_ctx* ctx, const uint8_t* d_imgs, size_t img_bytes, int n_frames, int w, int h,
                               size_t stride, const orbb200_proj_queries* d_queries, int nq_per_frame,
                               float th, float nnratio, float min_x, float min_y, float inv_w, float inv_h,
                               int32_t* d_out_best_idx, int32_t* d_out_best_dist, int32_t* d_nmatches)
{
    CTX_ENTER(ctx);
    if (n_frames <= 0 || 2 * n_frames > c.maxBatch || !d_queries || nq_per_frame < 0 || !d_out_best_idx || !d_out_best_dist || !d_nmatches) {
        c.err = "stereo_step: bad argument"; return ORBB200_ERR_ARG;
    }
    int rc = orbb200_extract_device(ctx, d_imgs, img_bytes, 2 * n_frames, w, h, stride);
    if (rc != ORBB200_OK) return rc;
    const int kpi = c.cur->g.kpPerImg, nq = nq_per_frame;
    // The frame/job descriptors only hold pointers and parameters: build them once per distinct argument
    // set and keep them on the device, so that the steady-state step never synchronises with the host.
    StepPlan key{};
    key.q = *d_queries; key.nq = nq; key.n_frames = n_frames; key.kpi = kpi; key.th = th; key.nnratio = nnratio;
    key.minX = min_x; key.minY = min_y; key.invW = inv_w; key.invH = inv_h;
    key.o0 = d_out_best_idx; key.o1 = d_out_best_dist; key.o2 = d_nmatches; key.stereo = c.stepStereo ? 1 : 0;
    StepPlan* plan = nullptr;
    for (auto& p : c.plans)
        if (memcmp(&p, &key, offsetof(StepPlan, dF)) == 0) { plan = &p; break; }
    if (!plan) {
        if (c.plans.size() >= 16) {   // bounded cache: drop everything (rare: callers reuse a few buffer sets)
            cudaStreamSynchronize(c.stream);
            for (auto& p : c.plans) cudaFree(p.block);
            c.plans.clear();
        }
        const size_t perFrame = align_up(4 * (size_t)(GRID_CELLS + 1), 256) + align_up(4 * (size_t)kpi, 256) * 2 + align_up(16 * (size_t)kpi, 256) +
                                align_up(4 * win_scratch_ints(kpi, nq), 256);
        const size_t bytes = align_up(sizeof(FrameDev) * n_frames, 256) + align_up(sizeof(WinJob) * n_frames, 256) + 256 + perFrame * n_frames + 4096;
        ORBB200_CUDA_OK(c, cudaMalloc(&key.block, bytes));
        Arena A((uint8_t*)key.block, bytes);
        key.dF = A.take<FrameDev>(n_frames);
        key.dJ = A.take<WinJob>(n_frames);
        float* dsf = A.take<float>(MAX_LEVELS);
        std::vector<FrameDev> hF(n_frames);
        std::vector<WinJob> hJ(n_frames);
        for (int i = 0; i < n_frames; i++) {
            FrameDev& f = hF[i];
            const int img = 2 * i;
            f.kps = c.d_kps + (size_t)img * kpi; f.desc = c.d_desc + (size_t)img * kpi * 32;
            f.uRight = c.stepStereo ? c.d_uRight + (size_t)img * kpi : nullptr;
            f.n_ptr = c.d_counts + img; f.n = kpi;
            f.cellStart = A.take<int32_t>(GRID_CELLS + 1); f.cellItems = A.take<int32_t>(kpi); f.cellKp = A.take<int4>(kpi);
            f.minX = min_x; f.minY = min_y; f.invW = inv_w; f.invH = inv_h;
            WinJob& J = hJ[i];
            memset(&J, 0, sizeof(J));
            J.frame = key.dF + i; J.nq = nq; J.mode = WM_PROJ; J.kpCap = kpi; J.th = th; J.nnratio = nnratio; J.scaleFactors = dsf; J.nLevels = c.nlevels;
            const size_t o = (size_t)i * nq;
            J.q_valid = d_queries->q_valid ? d_queries->q_valid + o : nullptr;
            J.q_x = d_queries->q_u + o; J.q_y = d_queries->q_v + o; J.q_aux = d_queries->q_uR + o; J.q_level = d_queries->q_level + o;
            J.q_viewcos = d_queries->q_viewcos + o; J.q_desc = d_queries->q_desc + o * 32;
            J.q_obs_pos = d_queries->q_obs_pos ? d_queries->q_obs_pos + o : nullptr;
            J.scratch = A.take<int>(win_scratch_ints(kpi, nq));
            J.out_best_idx = d_out_best_idx + o; J.out_best_dist = d_out_best_dist + o;
            J.out_per_kp = A.take<int32_t>(kpi);
            J.out_nmatches = d_nmatches + i;
        }
        ORBB200_CUDA_OK(c, cudaMemcpyAsync(key.dF, hF.data(), sizeof(FrameDev) * n_frames, cudaMemcpyHostToDevice, c.stream));
        ORBB200_CUDA_OK(c, cudaMemcpyAsync(key.dJ, hJ.data(), sizeof(WinJob) * n_frames, cudaMemcpyHostToDevice, c.stream));
        ORBB200_CUDA_OK(c, cudaMemcpyAsync(dsf, c.scale.data(), sizeof(float) * c.nlevels, cudaMemcpyHostToDevice, c.stream));
        ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));   // one-time: host vectors go out of scope
        c.plans.push_back(key);
        plan = &c.plans.back();
    }
    FrameDev* dF = plan->dF;
    WinJob* dJ = plan->dJ;
    if (c.stepStereo) {
        StageTimer t(c, 8);
        launch_stereo(c, n_frames, 0, 1, 2, c.stepMb, c.stepMbf, c.d_invScale, c.d_nKept);
        c.stereoValid = true;
    }
    { StageTimer t(c, 6); launch_grid_build(c, dF, n_frames); }
    { StageTimer t(c, 7); launch_window_match(c, dJ, n_frames, nq, c.cur->g.kpPerImg); }
    ORBB200_CUDA_OK(c, cudaGetLastError());
    return ORBB200_OK;
}

int orbb200_stereo_step_host(orbb200_ctx* ctx, const uint8_t* h_imgs, int n_frames, int w, int h, size_t stride,
                             const orbb200_proj_queries* hq, int nq,
                             float th, float nnratio, float min_x, float min_y, float inv_w, float inv_h,
                             orbb200_kp_t* h_kps, uint8_t* h_desc, int cap_per_img,
                             int32_t* h_counts, int32_t* h_best_idx, int32_t* h_best_dist, int32_t* h_nmatches)
{
    CTX_ENTER(ctx);
    if (!h_imgs || n_frames <= 0 || 2 * n_frames > c.maxBatch || !hq || nq < 0 || !h_kps || !h_desc || !h_counts || !h_best_idx || !h_best_dist ||
        !h_nmatches || stride < (size_t)w) { c.err = "stereo_step_host: bad argument"; return ORBB200_ERR_ARG; }
    const int ni = 2 * n_frames;
    const size_t imgBytes = (size_t)h * stride, Q = (size_t)n_frames * nq;
    const size_t need = align_up(imgBytes * ni, 256) + align_up(Q, 256) * 2 + align_up(Q * 4, 256) * 7 + align_up(Q * 32, 256) + align_up(4 * (size_t)n_frames, 256) + 4096;
    if (need > c.d_step_bytes) {
        cudaStreamSynchronize(c.stream);
        if (c.d_step) cudaFree(c.d_step);
        c.d_step = nullptr; c.d_step_bytes = 0;
        ORBB200_CUDA_OK(c, cudaMalloc(&c.d_step, need));
        c.d_step_bytes = need;
    }
    Arena A(c.d_step, c.d_step_bytes);
    uint8_t* dImgs = A.take<uint8_t>(imgBytes * ni);
    orbb200_proj_queries dq{};
    auto up = [&](const void* src, size_t bytes) -> void* {
        if (!src) return nullptr;
        uint8_t* d = A.take<uint8_t>(bytes);
        cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, c.stream);
        return d;
    };
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(dImgs, h_imgs, imgBytes * ni, cudaMemcpyHostToDevice, c.stream));
    dq.q_valid = (const uint8_t*)up(hq->q_valid, Q); dq.q_u = (const float*)up(hq->q_u, Q * 4); dq.q_v = (const float*)up(hq->q_v, Q * 4);
    dq.q_uR = (const float*)up(hq->q_uR, Q * 4); dq.q_level = (const int32_t*)up(hq->q_level, Q * 4);
    dq.q_viewcos = (const float*)up(hq->q_viewcos, Q * 4); dq.q_desc = (const uint8_t*)up(hq->q_desc, Q * 32);
    dq.q_obs_pos = (const uint8_t*)up(hq->q_obs_pos, Q);
    if (!dq.q_u || !dq.q_v || !dq.q_uR || !dq.q_level || !dq.q_viewcos || !dq.q_desc) { c.err = "stereo_step_host: missing query array"; return ORBB200_ERR_ARG; }
    int32_t* dBi = A.take<int32_t>(Q);
    int32_t* dBd = A.take<int32_t>(Q);
    int32_t* dNm = A.take<int32_t>(n_frames);
    int rc = orbb200_stereo_step_device(ctx, dImgs, imgBytes, n_frames, w, h, stride, &dq, nq, th, nnratio, min_x, min_y, inv_w, inv_h, dBi, dBd, dNm);
    if (rc != ORBB200_OK) return rc;
    const int kpi = c.cur->g.kpPerImg, take = std::min(cap_per_img, kpi);
    ORBB200_CUDA_OK(c, cudaMemcpy2DAsync(h_kps, (size_t)cap_per_img * sizeof(orbb200_kp_t), c.d_kps, (size_t)kpi * sizeof(orbb200_kp_t),
                                         (size_t)take * sizeof(orbb200_kp_t), ni, cudaMemcpyDeviceToHost, c.stream));
    ORBB200_CUDA_OK(c, cudaMemcpy2DAsync(h_desc, (size_t)cap_per_img * 32, c.d_desc, (size_t)kpi * 32, (size_t)take * 32, ni, cudaMemcpyDeviceToHost, c.stream));
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(h_counts, c.d_counts, 4 * (size_t)ni, cudaMemcpyDeviceToHost, c.stream));
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(h_best_idx, dBi, 4 * Q, cudaMemcpyDeviceToHost, c.stream));
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(h_best_dist, dBd, 4 * Q, cudaMemcpyDeviceToHost, c.stream));
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(h_nmatches, dNm, 4 * (size_t)n_frames, cudaMemcpyDeviceToHost, c.stream));
    return ORBB200_OK;
}

// ---- batched frame step: stereo front camera + birdview ---------------------------------------------------------------
static int frame_step_check(Ctx& c, const orbb200_frame_step_params* p)
{
    if (!p || p->n_frames <= 0 || 2 * p->n_frames > c.maxBatch || p->w <= 0 || p->h <= 0 || p->stride < (size_t)p->w) { c.err = "frame_step: bad argument"; return ORBB200_ERR_ARG; }
    if (p->bird_w < 0 || p->bird_h < 0 || (p->bird_w > 0 && (p->bird_h <= 0 || p->bird_stride < (size_t)p->bird_w || p->bird_nfeatures <= 0 || p->bird_window <= 0))) {
        c.err = "frame_step: bad birdview argument"; return ORBB200_ERR_ARG;
    }
    if (p->map && p->map->ctx != &c) { c.err = "frame_step: the map belongs to another context"; return ORBB200_ERR_ARG; }
    return ORBB200_OK;
}

int orbb200_frame_step_device(orbb200_ctx* ctx, const orbb200_frame_step_params* p, const orbb200_frame_step_inputs* in,
                              const orbb200_frame_step_outputs* out)
{
    CTX_ENTER(ctx);
    int rc = frame_step_check(c, p);
    if (rc != ORBB200_OK) return rc;
    const int n = p->n_frames;
    const bool hasMap = p->map && p->map->n > 0, hasBird = p->bird_w > 0, stereo = p->mb > 0.f;
    if (!in || !out || !in->imgs || (hasBird && !in->bird_imgs) || (hasMap && (!in->poses || !out->map_best_idx || !out->map_best_dist || !out->map_nmatches)) ||
        (hasBird && (!out->bird_matches12 || !out->bird_nmatches))) { c.err = "frame_step: missing input or output array"; return ORBB200_ERR_ARG; }
    // birdview front-end + the query arrays of BirdviewMatch(previous, current): independent of the front camera, FP64- and
    // latency-bound where the front extraction is integer-ALU-bound -> on a side stream beside it, TOGETHER WITH ITS OWN MATCHING
    // (grid of the birdview frame + the BirdviewMatch job): the two paths only meet at the end of the step.  (Per-stage timing keeps
    // everything on one stream so that a stage's events bracket its kernels alone.)
    BirdStepView bv{};
    const bool fork = hasBird && (c.forkBird || n <= 8) && !c.timing;      // a few frames leave most SMs idle: the two paths overlap
    if (hasBird) {                                  // the plan's pools (no launch yet): the frame plan below points into them
        rc = bird_step_view(c, p->bird_w, p->bird_h, p->bird_nfeatures, n, &bv);
        if (rc != ORBB200_OK) return rc;
    }
    const ShapeTables* shape = get_shape(c, p->w, p->h);
    if (!shape) return c.err.find("exceeds") != std::string::npos ? ORBB200_ERR_ARG : ORBB200_ERR_UNSUPPORTED;
    const int kpi = shape->g.kpPerImg, mapN = hasMap ? p->map->n : 0;
    FramePlan key{};
    key.n = n; key.kpi = kpi; key.birdKpi = hasBird ? bv.kpPerImg : 0; key.mapN = mapN; key.stereo = stereo ? 1 : 0; key.birdWindow = p->bird_window;
    key.birdOri = p->bird_check_ori; key.hasBird = hasBird ? 1 : 0; key.th = p->th; key.nnratio = p->nnratio; key.minX = p->min_x; key.minY = p->min_y;
    key.invW = p->inv_w; key.invH = p->inv_h;
    // Frame: mfGridElementWidthInvBirdview = FRAME_GRID_COLS / width, HeightInv = FRAME_GRID_ROWS / height (src/Frame.cc:350-351)
    key.birdInvW = hasBird ? (float)GRID_COLS / (float)p->bird_w : 0.f; key.birdInvH = hasBird ? (float)GRID_ROWS / (float)p->bird_h : 0.f;
    key.birdRatio = p->bird_nnratio; key.map = hasMap ? p->map : nullptr; key.birdKps = bv.d_kps; key.birdQx = bv.d_qx;
    key.oBi = out->map_best_idx; key.oBd = out->map_best_dist; key.oNm = out->map_nmatches; key.oM12 = out->bird_matches12; key.oBnm = out->bird_nmatches;
    FramePlan* plan = nullptr;
    for (auto& q : c.framePlans)
        if (memcmp(&q, &key, offsetof(FramePlan, dF)) == 0) { plan = &q; break; }
    if (!plan) {
        c.allocEpoch++;
        if (c.framePlans.size() >= 16) {
            cudaStreamSynchronize(c.stream);
            for (auto& q : c.framePlans) cudaFree(q.block);
            c.framePlans.clear();
        }
        const int bk = key.birdKpi;
        const size_t gridBytes = align_up(4 * (size_t)(GRID_CELLS + 1), 256);
        const size_t perFront = gridBytes + align_up(4 * (size_t)kpi, 256) * 2 + align_up(16 * (size_t)kpi, 256) + align_up(4 * win_scratch_ints(kpi, mapN), 256) +
                                align_up((size_t)mapN, 256) + align_up(4 * (size_t)mapN, 256) * 5;
        const size_t perBird = hasBird ? gridBytes + align_up(4 * (size_t)bk, 256) * 2 + align_up(16 * (size_t)bk, 256) + align_up(4 * win_scratch_ints(bk, bk), 256) : 0;
        const size_t bytes = align_up(sizeof(FrameDev) * 2 * n, 256) + align_up(sizeof(WinJob) * 2 * n, 256) + 1024 + (perFront + perBird) * n + align_up(4 * (size_t)n, 256) + 8192;
        ORBB200_CUDA_OK(c, cudaMalloc(&key.block, bytes));
        Arena A((uint8_t*)key.block, bytes);
        key.dF = A.take<FrameDev>(2 * n);
        key.dJ = A.take<WinJob>(2 * n);
        float* dsf = A.take<float>(MAX_LEVELS);
        key.inView = A.take<uint8_t>((size_t)n * std::max(mapN, 1));
        key.u = A.take<float>((size_t)n * std::max(mapN, 1)); key.v = A.take<float>((size_t)n * std::max(mapN, 1));
        key.uR = A.take<float>((size_t)n * std::max(mapN, 1)); key.viewcos = A.take<float>((size_t)n * std::max(mapN, 1));
        key.level = A.take<int32_t>((size_t)n * std::max(mapN, 1));
        key.count = A.take<int32_t>(n);
        std::vector<FrameDev> hF(2 * n);
        std::vector<WinJob> hJ;
        for (int i = 0; i < n; i++) {              // front frames: grid over the left image's keypoints
            FrameDev& f = hF[i];
            const int img = 2 * i;
            f.kps = c.d_kps + (size_t)img * kpi; f.desc = c.d_desc + (size_t)img * kpi * 32;
            f.uRight = stereo ? c.d_uRight + (size_t)img * kpi : nullptr;
            f.n_ptr = c.d_counts + img; f.n = kpi;
            f.cellStart = A.take<int32_t>(GRID_CELLS + 1); f.cellItems = A.take<int32_t>(kpi); f.cellKp = A.take<int4>(kpi);
            f.minX = p->min_x; f.minY = p->min_y; f.invW = p->inv_w; f.invH = p->inv_h;
            if (!hasMap) continue;
            WinJob J;
            memset(&J, 0, sizeof(J));
            const size_t o = (size_t)i * mapN;
            J.frame = key.dF + i; J.nq = mapN; J.mode = WM_PROJ; J.kpCap = kpi; J.th = p->th; J.nnratio = p->nnratio; J.scaleFactors = dsf; J.nLevels = c.nlevels;
            J.q_valid = key.inView + o; J.q_x = key.u + o; J.q_y = key.v + o; J.q_aux = key.uR + o; J.q_level = key.level + o; J.q_viewcos = key.viewcos + o;
            J.q_desc = p->map->d_desc;             // the map points' representative descriptors, shared by every frame
            J.scratch = A.take<int>(win_scratch_ints(kpi, mapN));
            J.out_best_idx = out->map_best_idx + o; J.out_best_dist = out->map_best_dist + o;
            J.out_per_kp = A.take<int32_t>(kpi);
            J.out_nmatches = out->map_nmatches + i;
            hJ.push_back(J);
        }
        for (int i = 0; hasBird && i < n; i++) {   // birdview frames: grid over frame i, queries = frame i-1 (slot i of the query arrays)
            FrameDev& f = hF[n + i];
            f.kps = bv.d_kps + (size_t)i * bk; f.desc = bv.d_desc + (size_t)i * bk * 32; f.uRight = nullptr;
            f.n_ptr = bv.d_counts + i; f.n = bk;
            f.cellStart = A.take<int32_t>(GRID_CELLS + 1); f.cellItems = A.take<int32_t>(bk); f.cellKp = A.take<int4>(bk);
            f.minX = 0.f; f.minY = 0.f; f.invW = key.birdInvW; f.invH = key.birdInvH;
            WinJob J;
            memset(&J, 0, sizeof(J));
            const size_t qo = (size_t)i * bk;
            J.frame = key.dF + n + i; J.nq = bk; J.mode = WM_BIRD; J.levelMode = 0; J.checkOri = p->bird_check_ori; J.kpCap = bk;
            J.th = (float)p->bird_window; J.nnratio = p->bird_nnratio; J.scaleFactors = dsf; J.nLevels = c.nlevels;
            J.q_valid = bv.d_qvalid + qo; J.q_x = bv.d_qx + qo; J.q_y = bv.d_qy + qo; J.q_level = bv.d_qlevel + qo; J.q_angle = bv.d_qangle + qo;
            J.q_desc = i == 0 ? bv.d_carryDesc : bv.d_desc + (size_t)(i - 1) * bk * 32;
            J.scratch = A.take<int>(win_scratch_ints(bk, bk));
            J.out_per_query = out->bird_matches12 + (size_t)i * bk;
            J.out_per_kp = A.take<int32_t>(bk);
            J.out_nmatches = out->bird_nmatches + i;
            hJ.push_back(J);
        }
        key.nFrames = hasBird ? 2 * n : n; key.nJobs = (int)hJ.size();
        key.maxNq = std::max(mapN, hasBird ? bk : 0); key.maxKpCap = std::max(kpi, hasBird ? bk : 0);
        ORBB200_CUDA_OK(c, cudaMemcpyAsync(key.dF, hF.data(), sizeof(FrameDev) * hF.size(), cudaMemcpyHostToDevice, c.stream));
        if (!hJ.empty()) ORBB200_CUDA_OK(c, cudaMemcpyAsync(key.dJ, hJ.data(), sizeof(WinJob) * hJ.size(), cudaMemcpyHostToDevice, c.stream));
        ORBB200_CUDA_OK(c, cudaMemcpyAsync(dsf, c.scale.data(), sizeof(float) * c.nlevels, cudaMemcpyHostToDevice, c.stream));
        ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));   // one-time: host vectors go out of scope
        c.framePlans.push_back(key);
        plan = &c.framePlans.back();
    }
    const int nFrontJobs = hasMap ? n : 0, nBirdJobs = plan->nJobs - nFrontJobs, birdKpi = plan->birdKpi;
    if (hasBird) {
        cudaStream_t main = c.stream;
        struct Restore { cudaStream_t& s; cudaStream_t v; ~Restore() { s = v; } } restore{c.stream, main};
        if (fork) {
            // (the host variant records the event itself, right behind the birdview upload: the birdview path -- the longer of the
            // two -- then starts while the front images are still on their way)
            if (!c.birdForkRecorded) ORBB200_CUDA_OK(c, cudaEventRecord(c.evBirdFork, main));
            ORBB200_CUDA_OK(c, cudaStreamWaitEvent(c.streamBird, c.evBirdFork, 0));
            c.stream = c.streamBird;
        }
        rc = bird_step_enqueue(c, p->bird_w, p->bird_h, p->bird_nfeatures, n, in->bird_imgs, (size_t)p->bird_h * p->bird_stride, p->bird_stride,
                               p->chain != 0, &bv);
        if (rc != ORBB200_OK) return rc;
        if (fork) {
            launch_grid_build(c, plan->dF + n, n);
            if (nBirdJobs > 0) launch_window_match(c, plan->dJ + nFrontJobs, nBirdJobs, birdKpi, birdKpi);
            ORBB200_CUDA_OK(c, cudaEventRecord(c.evBirdJoin, c.streamBird));        // the birdview results are complete here
            rc = bird_step_carry(c, bv, n);                                          // (only the next step needs these copies)
            if (rc != ORBB200_OK) return rc;
            ORBB200_CUDA_OK(c, cudaEventRecord(c.evBirdCarry, c.streamBird));
        }
    }
    // front camera: both images of every frame in one extraction, then the stereo matcher on its pools
    rc = orbb200_extract_device(ctx, in->imgs, (size_t)p->h * p->stride, 2 * n, p->w, p->h, p->stride);
    if (rc != ORBB200_OK) return rc;
    if (stereo) {
        StageTimer t(c, 8);
        launch_stereo(c, n, 0, 1, 2, p->mb, p->mbf, c.d_invScale, c.d_nKept);
        c.stereoValid = true;
    }
    if (hasMap) {                                  // Frame::isInFrustum for every (frame, map point)
        StageTimer t(c, 13);
        FrustumJob J{};
        J.poses = in->poses; J.nFrames = n; J.cosLimit = p->viewing_cos_limit; J.n = mapN;
        J.pos = p->map->d_pos; J.normal = p->map->d_normal; J.maxDist = p->map->d_maxDist; J.minDist = p->map->d_minDist; J.candidate = nullptr;
        J.inView = plan->inView; J.u = plan->u; J.v = plan->v; J.uR = plan->uR; J.level = plan->level; J.viewcos = plan->viewcos; J.count = plan->count;
        launch_frustum(c, J);
    }
    if (fork) {                                    // the birdview frames were matched on their own stream
        launch_grid_build(c, plan->dF, n);
        if (nFrontJobs > 0) launch_window_match(c, plan->dJ, nFrontJobs, std::max(mapN, 1), kpi);
        // The host variant joins later: its front-camera downloads go ahead of the wait for the (longer) birdview path, and the carry
        // copies are only waited for at the very end.  Everybody else gets a fully joined stream back.
        if (c.deferBirdJoin) c.birdJoinPending = true;
        else {
            ORBB200_CUDA_OK(c, cudaStreamWaitEvent(c.stream, c.evBirdJoin, 0));
            ORBB200_CUDA_OK(c, cudaStreamWaitEvent(c.stream, c.evBirdCarry, 0));
        }
    } else {
        { StageTimer t(c, 6); launch_grid_build(c, plan->dF, plan->nFrames); }
        if (plan->nJobs > 0) { StageTimer t(c, 7); launch_window_match(c, plan->dJ, plan->nJobs, plan->maxNq, plan->maxKpCap); }
        if (hasBird) {
            rc = bird_step_carry(c, bv, n);
            if (rc != ORBB200_OK) return rc;
        }
    }
    ORBB200_CUDA_OK(c, cudaGetLastError());
    return ORBB200_OK;
}

int orbb200_frame_step_host(orbb200_ctx* ctx, const orbb200_frame_step_params* p, const orbb200_frame_step_inputs* in,
                            const orbb200_frame_step_outputs* out)
{
    CTX_ENTER(ctx);
    int rc = frame_step_check(c, p);
    if (rc != ORBB200_OK) return rc;
    if (!in || !out || !in->imgs) { c.err = "frame_step_host: missing input"; return ORBB200_ERR_ARG; }
    const int n = p->n_frames, ni = 2 * n;
    const bool hasMap = p->map && p->map->n > 0, hasBird = p->bird_w > 0;
    const int mapN = hasMap ? p->map->n : 0;
    const int bk = hasBird ? orbb200_bird_max_keypoints(ctx, p->bird_w, p->bird_h, p->bird_nfeatures) : 0;
    if (hasBird && bk <= 0) return ORBB200_ERR_CUDA;
    if ((hasBird && !in->bird_imgs) || (hasMap && !in->poses)) { c.err = "frame_step_host: missing input"; return ORBB200_ERR_ARG; }
    const size_t imgBytes = (size_t)p->h * p->stride, birdBytes = hasBird ? (size_t)p->bird_h * p->bird_stride : 0;
    const size_t Qm = (size_t)n * mapN, Qb = (size_t)n * bk;
    const size_t need = align_up(imgBytes * ni, 256) + align_up(birdBytes * n, 256) + align_up(sizeof(orbb200_camera_pose) * n, 256) + align_up(4 * Qm, 256) * 2 +
                        align_up(4 * Qb, 256) + align_up(4 * (size_t)n, 256) * 2 + 4096;
    if (need > c.d_fstep_bytes) {
        c.allocEpoch++;
        cudaStreamSynchronize(c.stream);
        if (c.d_fstep) cudaFree(c.d_fstep);
        c.d_fstep = nullptr; c.d_fstep_bytes = 0;
        for (auto& q : c.framePlans) cudaFree(q.block);     // their output pointers refer to the old staging block
        c.framePlans.clear();
        ORBB200_CUDA_OK(c, cudaMalloc(&c.d_fstep, need));
        c.d_fstep_bytes = need;
    }
    Arena A(c.d_fstep, c.d_fstep_bytes);
    orbb200_frame_step_inputs din{};
    orbb200_frame_step_outputs dout{};
    // One or a few frames from pageable memory (the per-frame call pattern): inputs and results pass through the pinned staging
    // block -- a host memcpy + one DMA each way instead of a dozen serialised bounces of the copy engine; the results reach the
    // caller's buffers in orbb200_sync().  Pinned caller buffers (the throughput path) are used directly.
    bool staged = false;
    if (c.stageUploads && !c.hostCopies.empty()) { ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream)); deliver_host_copies(c); }
    {
        const int kpi0 = c.gmax.kpPerImg;
        const size_t inB = imgBytes * ni + birdBytes * n + sizeof(orbb200_camera_pose) * n + 1024;
        const size_t outB = (size_t)ni * kpi0 * 60 + (size_t)n * kpi0 * 8 + 8 * Qm + (size_t)n * bk * 64 + 64 * (size_t)n + 8192;
        if (c.stageUploads && n <= 8 && inB <= STAGE_D2H_OFF - STAGE_H2D_OFF && outB <= STAGE_LIMIT - STAGE_D2H_OFF) {
            cudaPointerAttributes at{};
            const bool pageable = cudaPointerGetAttributes(&at, in->imgs) != cudaSuccess || at.type == cudaMemoryTypeUnregistered;
            cudaGetLastError();
            staged = pageable && ensure_scratch(c, 0, STAGE_LIMIT);
        }
    }
    // The staged step as one graph (see Ctx::FrameGraph): state 0 -> run eagerly, 1 -> capture + launch, 2 -> replay.
    Ctx::FrameGraph* fg = nullptr;
    if (staged && c.useGraphs && c.frameGraph && !c.timing) {
        if (c.frameGraphEpoch != c.allocEpoch) {
            ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
            for (auto& q : c.frameGraphs) if (q.exec) cudaGraphExecDestroy(q.exec);
            c.frameGraphs.clear();
            c.frameGraphEpoch = c.allocEpoch;
        }
        std::vector<uint8_t> key(sizeof(*p) + 16 + 2 * sizeof(void*));
        memcpy(key.data(), p, sizeof(*p));
        const uint32_t want = (out->kps ? 1u : 0u) | (out->desc ? 2u : 0u) | (out->counts ? 4u : 0u) | (out->u_right ? 8u : 0u) | (out->depth ? 16u : 0u) |
                              (out->map_best_idx ? 32u : 0u) | (out->map_best_dist ? 64u : 0u) | (out->map_nmatches ? 128u : 0u) | (out->bird_kps ? 256u : 0u) |
                              (out->bird_desc ? 512u : 0u) | (out->bird_counts ? 1024u : 0u) | (out->bird_matches12 ? 2048u : 0u) | (out->bird_nmatches ? 4096u : 0u);
        const int32_t caps[3] = {(int32_t)want, out->cap, out->bird_cap};
        memcpy(key.data() + sizeof(*p), caps, 12);
        const void* ptrs[2] = {c.h_scratch, c.d_fstep};
        memcpy(key.data() + sizeof(*p) + 16, ptrs, sizeof(ptrs));
        for (auto& q : c.frameGraphs)
            if (q.key == key) { fg = &q; break; }
        if (!fg) {
            if (c.frameGraphs.size() >= 8) {            // more parameter sets than slots: start over rather than run the newcomers eagerly for ever
                ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
                for (auto& q : c.frameGraphs) if (q.exec) cudaGraphExecDestroy(q.exec);
                c.frameGraphs.clear();
            }
            c.frameGraphs.push_back(Ctx::FrameGraph{key, 0, nullptr, 0});
            fg = &c.frameGraphs.back();
        }
    }
    const bool replay = fg && fg->state == 2, capture = fg && fg->state == 1;
    const unsigned long long epoch0 = c.allocEpoch;
    const long long launches0 = c.launches;
    struct CaptureGuard {      // an error return between Begin and EndCapture must not leave the stream capturing
        cudaStream_t s; bool on;
        ~CaptureGuard() { if (on) { cudaGraph_t g = nullptr; cudaStreamEndCapture(s, &g); if (g) cudaGraphDestroy(g); cudaGetLastError(); } }
    } guard{c.stream, false};
    if (capture) {
        ORBB200_CUDA_OK(c, cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeRelaxed));
        guard.on = true;
    }
    size_t hIn = STAGE_H2D_OFF, hOut = STAGE_D2H_OFF;
    auto h2d = [&](void* dDst, const void* hSrc, size_t bytes) -> cudaError_t {
        if (staged) {
            uint8_t* hs = c.h_scratch + hIn;
            hIn += align_up(bytes, 64);
            memcpy(hs, hSrc, bytes);
            hSrc = hs;
        }
        return replay ? cudaSuccess : cudaMemcpyAsync(dDst, hSrc, bytes, cudaMemcpyHostToDevice, c.stream);
    };
    auto d2h = [&](void* hDst, size_t dpitch, const void* dSrc, size_t spitch, size_t width, size_t rows) -> cudaError_t {
        if (staged) {
            uint8_t* hs = c.h_scratch + hOut;
            hOut += align_up(width * rows, 64);
            c.hostCopies.push_back(Ctx::HostCopy{hDst, dpitch, hs, width, rows});
            return replay ? cudaSuccess : cudaMemcpy2DAsync(hs, width, dSrc, spitch, width, rows, cudaMemcpyDeviceToHost, c.stream);
        }
        return cudaMemcpy2DAsync(hDst, dpitch, dSrc, spitch, width, rows, cudaMemcpyDeviceToHost, c.stream);
    };
    uint8_t* dImgs = A.take<uint8_t>(imgBytes * ni);
    struct ForkFlag { bool& f; ~ForkFlag() { f = false; } } forkFlag{c.birdForkRecorded};
    if (hasBird) {                                  // the birdview image first: see orbb200_frame_step_device
        uint8_t* dBird = A.take<uint8_t>(birdBytes * n);
        ORBB200_CUDA_OK(c, h2d(dBird, in->bird_imgs, birdBytes * n));
        din.bird_imgs = dBird;
        if (!replay) { ORBB200_CUDA_OK(c, cudaEventRecord(c.evBirdFork, c.stream)); c.birdForkRecorded = true; }
    }
    ORBB200_CUDA_OK(c, h2d(dImgs, in->imgs, imgBytes * ni));
    din.imgs = dImgs;
    if (hasMap) {
        orbb200_camera_pose* dP = A.take<orbb200_camera_pose>(n);
        ORBB200_CUDA_OK(c, h2d(dP, in->poses, sizeof(orbb200_camera_pose) * n));
        din.poses = dP;
        dout.map_best_idx = A.take<int32_t>(Qm); dout.map_best_dist = A.take<int32_t>(Qm); dout.map_nmatches = A.take<int32_t>(n);
    }
    if (hasBird) { dout.bird_matches12 = A.take<int32_t>(Qb); dout.bird_nmatches = A.take<int32_t>(n); }
    struct JoinFlags { bool& d; bool& pnd; ~JoinFlags() { d = false; pnd = false; } } joinFlags{c.deferBirdJoin, c.birdJoinPending};
    if (replay) {                                   // host-side state the enqueueing path leaves behind
        const ShapeTables* st = get_shape(c, p->w, p->h);
        if (!st) { c.hostCopies.clear(); return ORBB200_ERR_UNSUPPORTED; }
        c.cur = st; c.curN = ni; c.stereoValid = p->mb > 0.f;
    } else {
        c.deferBirdJoin = true;
        rc = orbb200_frame_step_device(ctx, p, &din, &dout);
        c.deferBirdJoin = false;
        if (rc != ORBB200_OK) {
            if (c.birdJoinPending) { cudaStreamWaitEvent(c.stream, c.evBirdJoin, 0); cudaStreamWaitEvent(c.stream, c.evBirdCarry, 0); }
            c.hostCopies.clear();
            return rc;
        }
    }
    // join the birdview path (deferred by orbb200_frame_step_device on request): before the first birdview download, and the carry
    // copies before the call is over
    auto join_bird = [&]() -> cudaError_t {
        if (!c.birdJoinPending) return cudaSuccess;
        c.birdJoinPending = false;
        return cudaStreamWaitEvent(c.stream, c.evBirdJoin, 0);
    };
    const bool carryPending = c.birdJoinPending;
    const int kpi = c.cur->g.kpPerImg;
    if (out->kps || out->desc || out->u_right || out->depth) {
        if (out->cap <= 0) { c.err = "frame_step_host: cap"; c.hostCopies.clear(); return ORBB200_ERR_ARG; }
        const int take = std::min(out->cap, kpi);
        if (out->kps) ORBB200_CUDA_OK(c, d2h(out->kps, (size_t)out->cap * sizeof(orbb200_kp_t), c.d_kps, (size_t)kpi * sizeof(orbb200_kp_t), (size_t)take * sizeof(orbb200_kp_t), ni));
        if (out->desc) ORBB200_CUDA_OK(c, d2h(out->desc, (size_t)out->cap * 32, c.d_desc, (size_t)kpi * 32, (size_t)take * 32, ni));
        // mvuRight / mvDepth live in the left images' rows of the pools (image 2i)
        if (out->u_right && p->mb > 0.f) ORBB200_CUDA_OK(c, d2h(out->u_right, (size_t)out->cap * 4, c.d_uRight, (size_t)kpi * 8, (size_t)take * 4, n));
        if (out->depth && p->mb > 0.f) ORBB200_CUDA_OK(c, d2h(out->depth, (size_t)out->cap * 4, c.d_depth, (size_t)kpi * 8, (size_t)take * 4, n));
    }
    if (out->counts) ORBB200_CUDA_OK(c, d2h(out->counts, 4 * (size_t)ni, c.d_counts, 4 * (size_t)ni, 4 * (size_t)ni, 1));
    if (hasMap) {
        if (out->map_best_idx) ORBB200_CUDA_OK(c, d2h(out->map_best_idx, 4 * Qm, dout.map_best_idx, 4 * Qm, 4 * Qm, 1));
        if (out->map_best_dist) ORBB200_CUDA_OK(c, d2h(out->map_best_dist, 4 * Qm, dout.map_best_dist, 4 * Qm, 4 * Qm, 1));
        if (out->map_nmatches) ORBB200_CUDA_OK(c, d2h(out->map_nmatches, 4 * (size_t)n, dout.map_nmatches, 4 * (size_t)n, 4 * (size_t)n, 1));
    }
    if (hasBird) {
        ORBB200_CUDA_OK(c, join_bird());
        const orbb200_kp_t* bKps = nullptr; const uint8_t* bDesc = nullptr; const int32_t* bCnt = nullptr; int bcap = 0;
        rc = orbb200_bird_results_device(ctx, p->bird_w, p->bird_h, p->bird_nfeatures, &bKps, &bDesc, &bCnt, &bcap);
        if (rc != ORBB200_OK) { if (carryPending) cudaStreamWaitEvent(c.stream, c.evBirdCarry, 0); c.hostCopies.clear(); return rc; }
        if (out->bird_kps || out->bird_desc || out->bird_matches12) {
            if (out->bird_cap <= 0) { c.err = "frame_step_host: bird_cap"; c.hostCopies.clear(); return ORBB200_ERR_ARG; }
            const int take = std::min(out->bird_cap, bk);
            if (out->bird_kps) ORBB200_CUDA_OK(c, d2h(out->bird_kps, (size_t)out->bird_cap * sizeof(orbb200_kp_t), bKps, (size_t)bk * sizeof(orbb200_kp_t), (size_t)take * sizeof(orbb200_kp_t), n));
            if (out->bird_desc) ORBB200_CUDA_OK(c, d2h(out->bird_desc, (size_t)out->bird_cap * 32, bDesc, (size_t)bk * 32, (size_t)take * 32, n));
            if (out->bird_matches12) ORBB200_CUDA_OK(c, d2h(out->bird_matches12, (size_t)out->bird_cap * 4, dout.bird_matches12, (size_t)bk * 4, (size_t)take * 4, n));
        }
        if (out->bird_counts) ORBB200_CUDA_OK(c, d2h(out->bird_counts, 4 * (size_t)n, bCnt, 4 * (size_t)n, 4 * (size_t)n, 1));
        if (out->bird_nmatches) ORBB200_CUDA_OK(c, d2h(out->bird_nmatches, 4 * (size_t)n, dout.bird_nmatches, 4 * (size_t)n, 4 * (size_t)n, 1));
    }
    ORBB200_CUDA_OK(c, join_bird());
    if (carryPending) ORBB200_CUDA_OK(c, cudaStreamWaitEvent(c.stream, c.evBirdCarry, 0));
    if (capture) {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        guard.on = false;
        cudaError_t e = cudaStreamEndCapture(c.stream, &graph);
        if (e == cudaSuccess && std::getenv("ORBB200_TEST_CAPTURE_FAIL")) e = cudaErrorUnknown;      // (test hook: exercises the fall-back below)
        if (e == cudaSuccess && c.allocEpoch == epoch0) e = cudaGraphInstantiate(&exec, graph, 0);
        else if (e == cudaSuccess) e = cudaErrorUnknown;       // something was (re)allocated while recording: the recording is not trusted
        if (graph) cudaGraphDestroy(graph);
        if (e != cudaSuccess || !exec) {                       // nothing has run yet: do this call eagerly and never try this key again
            cudaGetLastError();
            fg->state = -1;
            c.launches = launches0;
            c.hostCopies.clear();
            return orbb200_frame_step_host(ctx, p, in, out);
        }
        fg->exec = exec; fg->launches = c.launches - launches0; fg->state = 2;
        ORBB200_CUDA_OK(c, cudaGraphLaunch(exec, c.stream));
    } else if (replay) {
        ORBB200_CUDA_OK(c, cudaGraphLaunch(fg->exec, c.stream));
        c.launches += fg->launches;
    } else if (fg && fg->state == 0) {
        // plans, pools and kernel attributes exist now; record the next call -- unless this one had to allocate (then once more)
        if (c.allocEpoch == epoch0) fg->state = 1;
    }
    return ORBB200_OK;
}

int orbb200_device_status(orbb200_ctx* ctx, int* status)
{
    CTX_ENTER(ctx);
    int32_t st = 0;
    ORBB200_CUDA_OK(c, cudaMemcpyAsync(&st, c.d_status, sizeof(st), cudaMemcpyDeviceToHost, c.stream));
    ORBB200_CUDA_OK(c, cudaStreamSynchronize(c.stream));
    deliver_host_copies(c);
    if (st != 0) cudaMemsetAsync(c.d_status, 0, sizeof(int32_t), c.stream);
    if (status) *status = st;
    if (st != 0) { c.err = "device-side capacity overflow (status " + std::to_string(st) + "): results of the last calls are incomplete"; return ORBB200_ERR_CAPACITY; }
    return ORBB200_OK;
}

int orbb200_bird_set_mask(orbb200_ctx* ctx, int w, int h, int nfeatures, int max_batch, const uint8_t* mask, size_t mask_stride)
{
    CTX_ENTER(ctx);
    if (w <= 0 || h <= 0 || nfeatures <= 0 || w > 4000 || h > 4000 || (mask && mask_stride < (size_t)w)) { c.err = "bird_set_mask: bad argument"; return ORBB200_ERR_ARG; }
    return bird_set_mask(c, w, h, nfeatures, max_batch, mask, mask_stride);
}

int orbb200_step_enable_stereo(orbb200_ctx* ctx, int enable, float mb, float mbf)
{
    CTX_ENTER(ctx);
    if (enable && !(mb > 0.f)) { c.err = "step_enable_stereo: mb must be positive"; return ORBB200_ERR_ARG; }
    c.stepStereo = enable != 0; c.stepMb = mb; c.stepMbf = mbf;
    return ORBB200_OK;
}

int orbb200_stage_timing(orbb200_ctx* ctx, int enable)
{
    CTX_ENTER(ctx);
    drain_stage_events(c);
    c.timing = enable != 0;
    return ORBB200_OK;
}

int orbb200_stage_times(orbb200_ctx* ctx, float* ms, int32_t* groups, int reset)
{
    CTX_ENTER(ctx);
    drain_stage_events(c);
    for (int i = 0; i < ORBB200_NUM_STAGES; i++) {
        if (ms) ms[i] = c.stageMs[i];
        if (groups) groups[i] = c.stageGroups[i];
        if (reset) { c.stageMs[i] = 0; c.stageGroups[i] = 0; }
    }
    return ORBB200_OK;
}

}  // extern "C"
