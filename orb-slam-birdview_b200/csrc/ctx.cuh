// Context object behind orbb200_ctx: parameters, per-shape geometry cache, HBM-resident pools.
#pragma once
#include <map>
#include <mutex>
#include <utility>

#include "orbb200_internal.cuh"

namespace orbb200 {

struct ShapeTables {
    Geom g;
    int2* d_xtab = nullptr;     // resize: per output x of levels>=1: {sx, a0 | a1<<16}
    int4* d_ytab = nullptr;     // resize: per output y of levels>=1: {sy0, sy1, b0, b1}
    int4* d_cells = nullptr;    // FAST cells of all levels, 3 x int4 each: {x0|y0<<16, x1|y1<<16, level, order}, {level offset, pitch, candOff, candCap}, {2^32/nw+1, 2^32/npr+1, -, -}
    int4* d_blurTiles = nullptr;  // blur tiles of all levels: {level, x0, y0, 0}, 128 x 32 pixels each
    int nBlurTiles = 0;
    DescribeMaps dmaps;         // TMA descriptors of this context's blurred pool for this shape's levels (host copy)
    DescribeMaps rmaps;         // TMA descriptors of the pyramid pool, entry l = source level l-1 with the staging box of resize level l
    CUtensorMap* d_rmaps = nullptr;
    bool resizeMapOk[MAX_LEVELS] = {false};
    CUtensorMap* d_dmaps = nullptr;   // the same in global memory (64-byte aligned): what describe_kernel hands to the TMA unit
    int4* d_resizeTiles = nullptr;  // resize tiles of levels >= 1, two int4 each: {x0, y0, srcRow0, rows}, {srcCol0, vectors, 2^16/vectors+1, 0}
    int resizeTileBase[MAX_LEVELS + 1] = {0};
    int resizeTileCount[MAX_LEVELS] = {0};
    bool resizeNarrow[MAX_LEVELS] = {false};  // taps of any 4 adjacent output columns span <= 8 source bytes (scale <= 2)
    int resizeSmemPitch[MAX_LEVELS] = {0}, resizeSmemRows[MAX_LEVELS] = {0};   // staged source window of a resize tile
    struct GraphExec { cudaGraphExec_t exec; long long launches; const uint8_t* stage; };
    std::map<int, GraphExec> graphs;   // captured extraction pipeline per image count
    int nFastCells = 0;         // entries of d_cells (cells the reference skips at the image edge are not listed)
    FastSmem fastSmem;          // shared-memory carve of fast_cells_kernel
    int4* d_groups = nullptr;   // the same cells as groups of up to 4 per cell row for fast_strip_kernel, 3 x int4 each
    int nFastGroups = 0;
    int fastGroupBase[MAX_LEVELS + 1] = {0};   // groups of level l: [fastGroupBase[l], fastGroupBase[l+1])
    FastStripSmem fastStrip;
};

struct WinJob;

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per (device, kernel) and SETS the limit, so a per-context record of
// "what I asked for" lets a second context on the same GPU lower the first one's limit.  Each kernel that may need
// more than 48 KB is instead raised ONCE per device to the device's opt-in maximum; the record is process-wide.
// Returns that maximum (bytes) or 0 after a CUDA error.  slot: one small integer per kernel instantiation.
enum SmemSlot { SMEM_FAST = 0, SMEM_FAST_SMALL, SMEM_FAST_STRIP, SMEM_OCTREE, SMEM_OCTREE_FEW, SMEM_WINDOW_MATCH, SMEM_BIRD_SELECT, SMEM_BIRD_SELECT_FEW, SMEM_BIRD_SUBPIX, SMEM_BIRD_SUBPIX3, SMEM_BIRD_SUBPIX2, SMEM_SLOTS };
inline size_t ensure_max_dynamic_smem(int device, const void* kernel, int slot)
{
    static std::mutex mu;
    static size_t limit[64][SMEM_SLOTS] = {};
    if (device < 0 || device >= 64) return 0;
    std::lock_guard<std::mutex> lk(mu);
    if (limit[device][slot]) return limit[device][slot];
    int optin = 0;
    if (cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess || optin <= 0) return 0;
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) return 0;
    const int dyn = optin - (int)fa.sharedSizeBytes;        // the opt-in maximum covers static + dynamic shared memory
    if (dyn <= 0 || cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn) != cudaSuccess) { cudaGetLastError(); return 0; }
    limit[device][slot] = (size_t)dyn;
    return (size_t)dyn;
}

// Cached device-side descriptors of one batched step (orbb200_stereo_step_device); the leading fields are
// the cache key.
struct StepPlan {
    orbb200_proj_queries q;
    int nq, n_frames, kpi;
    float th, nnratio, minX, minY, invW, invH;
    void *o0, *o1, *o2;
    int stereo;
    FrameDev* dF;       // <- offsetof(StepPlan, dF) ends the key
    WinJob* dJ;
    void* block;
};

// Cached device-side descriptors of one batched frame step (orbb200_frame_step_device); the fields up to `dF` are the cache key.
struct FramePlan {
    int n, kpi, birdKpi, mapN, stereo, birdWindow, birdOri, hasBird;
    float th, nnratio, minX, minY, invW, invH, birdInvW, birdInvH, birdRatio;
    const void* map; const void* birdKps; const void* birdQx;
    void *oBi, *oBd, *oNm, *oM12, *oBnm;
    FrameDev* dF;       // <- offsetof(FramePlan, dF) ends the key; front frames [0,n), birdview frames [n,2n)
    WinJob* dJ;         // front jobs [0,nFront), birdview jobs after them
    int nFrames, nJobs, maxNq, maxKpCap;
    uint8_t* inView; float *u, *v, *uR, *viewcos; int32_t* level; int32_t* count;     // isInFrustum results [n][mapN], [n]
    void* block;
};

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;          // side branch of the extraction graph (blur runs beside FAST + octree)
    cudaEvent_t evFork = nullptr, evJoin = nullptr;
    cudaStream_t stream4 = nullptr;          // one or two birdview images: the blurred pyramid beside FAST + retainBest + cornerSubPix
    cudaEvent_t evFork4 = nullptr, evJoin4 = nullptr;
    cudaStream_t stream3 = nullptr;          // one or two images: FAST + octree of level 0 beside the resize chain
    cudaEvent_t evFork0 = nullptr, evJoin0 = nullptr;
    bool pdl = true;                         // programmatic dependent launches inside the extraction chain (ORBB200_NO_PDL=1 turns them off)
    bool octreeSmemCand = true;              // one or two images: a level's candidates are swept from shared memory (ORBB200_NO_OCTREE_SMEM=1: from L2)
    bool hostGraph = true;                   // small host calls: upload and result delivery are nodes of the extraction graph (ORBB200_NO_HOST_GRAPH=1: copy-engine operations around it)
    bool splitLevel0 = true;                 // the level-0 branch for one or two images (ORBB200_NO_SPLIT=1 turns it off)
    cudaStream_t streamBird = nullptr;       // the birdview front-end of the batched frame step runs beside the front-camera extraction
    cudaEvent_t evBirdFork = nullptr, evBirdJoin = nullptr;
    cudaEvent_t evBirdCarry = nullptr;       // the carry copies behind the birdview matching (only the next step needs them)
    bool deferBirdJoin = false, birdJoinPending = false;   // orbb200_frame_step_host joins the birdview stream itself, behind its front-camera downloads
    bool birdForkRecorded = false;           // orbb200_frame_step_host has recorded evBirdFork behind the birdview upload
    bool forkBird = false;                   // ORBB200_FORK_BIRD=1 turns it on (measured slower: the GPU is already full)
    std::string err;
    long long launches = 0;
    size_t stageMaxBytes = 3u << 19;         // largest host call (bytes of images) that goes through the pinned staging block (ORBB200_STAGE_MAX_KB)
    bool stageUploads = true;                // small host-API batches upload through the pinned staging block (ORBB200_NO_STAGED_UPLOAD=1: straight from the caller's memory)
    bool fastCells = false;                  // ORBB200_FAST_CELLS=1: grid FAST with one CTA per cell (fast_cells_kernel) instead of the strip form
    int subpixCtasPerSm = 3;                 // resident CTAs per SM of bird_subpix5_kernel (tuning knob, ORBB200_SUBPIX_CTAS: 1-2 reach-5 patches, 3 reach-3, 4 reach-2)

    // ORBextractor parameters and tables (src/ORBextractor.cc:410-470)
    int nfeatures = 0;
    double scaleFactor = 1.2;
    int nlevels = 8, iniTh = 20, minTh = 7;
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> quota;

    int maxW = 0, maxH = 0, maxBatch = 0;
    Geom gmax;                               // geometry of the largest shape: sizes the pools
    std::map<std::pair<int, int>, ShapeTables> shapes;
    const ShapeTables* cur = nullptr;        // shape of the last extraction
    int curN = 0;

    // pools (per image blocks, maxBatch images)
    uint8_t* d_pyr = nullptr;                // [maxBatch][pyrBytes]   pyramid levels (row pitch 128-aligned)
    uint8_t* d_blur = nullptr;               // [maxBatch][pyrBytes]   GaussianBlur'ed levels
    uint32_t* d_cand = nullptr;              // [maxBatch][candPerImg] x:12|y:12|response:8, region coords
    uint16_t* d_nodeOf = nullptr;            // [maxBatch][candPerImg] octree scratch
    int32_t* d_candCount = nullptr;          // [maxBatch][MAX_LEVELS]
    uint32_t* d_lvlKp = nullptr;             // [maxBatch][kpPerImg]   octree winners, same packing
    int32_t* d_lvlCount = nullptr;           // [maxBatch][MAX_LEVELS]
    orbb200_kp_t* d_kps = nullptr;           // [maxBatch][kpPerImg]
    uint8_t* d_desc = nullptr;               // [maxBatch][kpPerImg][32]
    int32_t* d_counts = nullptr;             // [maxBatch]
    int32_t* d_status = nullptr;             // device-side error flags (octree overflow etc.)
    float* d_uRight = nullptr;               // [maxBatch][kpPerImg]   Frame::mvuRight of the last stereo matching (left images)
    float* d_depth = nullptr;                // [maxBatch][kpPerImg]   Frame::mvDepth
    int32_t* d_sad = nullptr;                // [maxBatch][kpPerImg]   SAD distance of each stereo match (-1 none)
    int32_t* d_nKept = nullptr;              // [maxBatch]             stereo matches kept per frame
    float* d_invScale = nullptr;             // [MAX_LEVELS]
    int32_t* d_rowStart = nullptr;           // [maxBatch/2+1][maxH+1] stereo row table (CSR over image rows)
    int32_t* d_rowItems = nullptr;           // [maxBatch/2+1][stereoItemCap]
    int stereoItemCap = 0;
    bool stereoValid = false;
    void* bird = nullptr;                    // birdview front-end plans (bird.cu)
    bool forkBlur = true;                    // blur on the side stream (ORBB200_SERIAL=1 in the environment turns it off)
    bool useGraphs = true;                   // replay the extraction launches from a CUDA graph
    bool stepStereo = false;                 // orbb200_step_enable_stereo: the batched step also runs stereo matching
    float stepMb = 0.f, stepMbf = 0.f;

    // staging for host entry points
    uint8_t* d_scratch = nullptr;            // generic device scratch (matcher uploads)
    size_t d_scratch_bytes = 0;
    uint8_t* h_scratch = nullptr;            // pinned generic
    size_t h_scratch_bytes = 0;
    uint8_t* h_pyrMirror = nullptr;          // pinned copy of one image's pyramid block (orbb200_pyramid_mirror)
    size_t h_pyrMirrorBytes = 0;
    bool mirrorPyramid = false;              // orbb200_set_pyramid_mirror: small host calls also deliver image 0's pyramid to the pinned mirror
    const ShapeTables* pyrMirrorFresh = nullptr;   // shape of the extraction whose pyramid the mirror currently holds (null: stale)
    uint8_t* d_step = nullptr;               // device staging of the host step (images, queries, results)
    size_t d_step_bytes = 0;

    std::vector<StepPlan> plans;
    std::vector<FramePlan> framePlans;
    // Small host frame steps (the one-frame-per-call pattern) replayed as ONE graph: uploads from the pinned staging block, both
    // front-ends, matching, downloads into the staging block.  A key's first call runs eagerly (it creates plans and pools), the
    // second is captured, later ones are a single cudaGraphLaunch.  allocEpoch counts every reallocation of something a captured
    // node may point to (staging blocks, birdview plans and masks, local maps, frame plans): a change drops all captured steps.
    struct FrameGraph { std::vector<uint8_t> key; int state; cudaGraphExec_t exec; long long launches; };
    std::vector<FrameGraph> frameGraphs;
    unsigned long long allocEpoch = 0, frameGraphEpoch = 0;
    bool warpCands = true;                   // small matcher jobs scan their windows with one warp per query (ORBB200_NO_WARP_CANDS=1: one thread per query)
    bool stageMatch = true;                  // matcher calls upload / download through the pinned staging block in one copy each (ORBB200_NO_STAGED_MATCH=1: array by array)
    bool selectTiers = false;                // ORBB200_SELECT_TIERS=1: the tiered 256-thread retainBest launches also for one or two images (A/B)
    bool subpixGenericWarp = false;          // ORBB200_SUBPIX_GENERIC=1: the generic warp-per-corner cornerSubPix also for the 5x5 window (A/B)
    bool frameGraph = true;                  // ORBB200_NO_FRAME_GRAPH=1 turns the replay off
    uint8_t* d_fstep = nullptr;              // device staging of the host frame step
    size_t d_fstep_bytes = 0;

    // per-stage timing (bench)
    bool timing = false;
    struct Pending { int stage; cudaEvent_t e0, e1; };
    std::vector<Pending> pending;
    // results of a small host-buffer frame step staged in pinned memory: delivered to the caller's (pageable) buffers by orbb200_sync()
    struct HostCopy { void* dst; size_t dpitch; const void* src; size_t width, rows; };
    std::vector<HostCopy> hostCopies;
    std::vector<cudaEvent_t> freeEvents;
    float stageMs[ORBB200_NUM_STAGES] = {};
    int stageGroups[ORBB200_NUM_STAGES] = {};
};

// RAII: events around one stage when timing is on
struct StageTimer {
    Ctx& c;
    Ctx::Pending p{};
    bool on;
    StageTimer(Ctx& c_, int stage);
    ~StageTimer();
};
void drain_stage_events(Ctx& c);

bool ensure_scratch(Ctx& c, size_t dev_bytes, size_t host_bytes);
// The pinned staging block of small host calls: uploads in the lower half, the result mirror in the upper half.
constexpr size_t STAGE_LIMIT = 8u << 20;
constexpr size_t STAGE_H2D_OFF = 0, STAGE_D2H_OFF = 4u << 20;   // the upload of a call and its download do not share bytes
void deliver_host_copies(Ctx& c);      // staged results of an earlier frame step -> the caller's buffers (the stream must have drained)
const ShapeTables* get_shape(Ctx& c, int w, int h);   // nullptr + c.err on failure
bool build_geom(const Ctx& c, int w, int h, Geom& g, std::string& err);

}  // namespace orbb200

struct orbb200_ctx { orbb200::Ctx c; };

namespace orbb200 {

#define ORBB200_CUDA_OK(c, call)                                                           \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            (c).err = std::string(#call) + ": " + cudaGetErrorString(e__);                 \
            return ORBB200_ERR_CUDA;                                                       \
        }                                                                                  \
    } while (0)

}  // namespace orbb200
