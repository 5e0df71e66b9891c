// Internal declarations shared by the CUDA translation units of liborbb200.so.
#pragma once
#include <cuda.h>          // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, libcuda is not linked)
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/orbb200.h"

namespace orbb200 {

constexpr int MAX_LEVELS = 12;
constexpr int EDGE_THRESHOLD = 19;      // reference src/ORBextractor.cc:74
constexpr int HALF_PATCH = 15;          // :73
constexpr int PATCH_SIZE = 31;          // :72
constexpr int FAST_BORDER = 16;         // EDGE_THRESHOLD-3, :776
constexpr int GRID_COLS = 64;           // include/Frame.h:40
constexpr int GRID_ROWS = 48;           // include/Frame.h:39
constexpr int GRID_CELLS = GRID_COLS * GRID_ROWS;
constexpr int TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 30;   // src/ORBmatcher.cc:37-39
constexpr int PYR_MARGIN_X = 32;        // bytes of reflect-101 border stored left of each level row (>= 4 used)
constexpr int PYR_MARGIN_Y = 4;         // border rows stored above/below each level (>= 3 used)
#ifndef ORBB200_BL_ROWS
#define ORBB200_BL_ROWS 48
#endif
constexpr int BL_ROWS = ORBB200_BL_ROWS;   // output rows per blur warp tile, 128 columns wide (6 extra rows of horizontal work per tile)
constexpr int RS_ROWS = 32;             // output rows per resize tile (128 columns wide)
constexpr int FT_PITCH = 49;            // u32 words per row of fast_cells_kernel's shared tiles (>= 2*ceil(68/4)+3 = 37), general case
constexpr int FT_PITCH_SMALL = 25;      // the same for shapes whose cell images are at most 44 pixels wide (2*11+3): half the shared memory
constexpr int FS_PITCH = 71;            // u32 words per row of fast_strip_kernel's shared tiles: 1 pad + 68 pair words (136 pixels) + 2 pads
#ifndef ORBB200_FS_THREADS
#define ORBB200_FS_THREADS 128
#endif
constexpr int FS_THREADS = ORBB200_FS_THREADS, FS_WARPS = FS_THREADS / 32;   // threads of a fast_strip_kernel CTA
constexpr int FS_MAX_CELLS = 8;         // cells per group (cells are >= 30 px wide: at most 4 fit the tile)

// Geometry of one pyramid level for one image shape (host-computed, passed to kernels by value).
struct LevelGeom {
    int w, h, pitch;        // level size; row pitch in bytes (multiple of 128)
    unsigned off;           // byte offset of the level inside one image's pyramid block
    float scale;            // mvScaleFactor[l]
    int quota;              // mnFeaturesPerLevel[l]
    // grid FAST (src/ORBextractor.cc:776-787)
    int maxBX, maxBY;       // level coords: FAST region is [16,maxBX) x [16,maxBY)
    int nCols, nRows, wCell, hCell;
    int cellBase, nCells;   // range in the flattened cell table
    // octree (src/ORBextractor.cc:543-546)
    int nIni;
    float hX;
    int candCap;            // candidate slots for this level
    unsigned candOff;       // offset (entries) in one image's candidate block
    int kpCap;              // keypoint slots (>= quota + slack)
    int kpOff;              // offset (entries) in one image's level-keypoint block
    int maxNodes;           // node table size for the octree kernel
    int xtabOff, ytabOff;   // offsets into the resize tables (level >= 1)
    int patchSize;          // (int)(PATCH_SIZE*scale)
};

struct Geom {
    int nlevels;
    int w, h;
    int iniTh, minTh;
    unsigned pyrBytes;      // bytes of one image's pyramid block
    unsigned candPerImg;    // candidate entries per image
    int kpPerImg;           // level-keypoint entries per image (== output capacity per image)
    int totalCells;
    LevelGeom lv[MAX_LEVELS];
};

// TMA descriptors of the blurred pyramid pool, one per level: 3-D u8 tensors {row bytes, rows, images} whose boxes are the
// 64 x 39-byte windows describe_kernel fetches around a keypoint.  The 39 x 39 patch starts at an arbitrary column, but
// a box must start on a 16-byte boundary of the row (measured: an unaligned inner coordinate raises "illegal
// instruction", tools/ubench/tma_box.cu), so the box starts at the patch column rounded down to 16 and is 15 + 39 -> 64
// bytes wide.
constexpr int DS_TBOX_W = 64, DS_TBOX_H = 39;
struct DescribeMaps { CUtensorMap m[MAX_LEVELS]; };

// Device-side frame (keypoints + descriptors + 64x48 CSR grid).
struct FrameDev {
    const orbb200_kp_t* kps;
    const uint8_t* desc;
    const float* uRight;        // may be null
    const int32_t* n_ptr;       // device count (results of an extraction) or null
    int n;                      // host-known count (uploaded frames) or capacity
    int32_t* cellStart;         // [GRID_CELLS+1]
    int32_t* cellItems;         // [cap]
    int4* cellKp;               // [cap] {x bits, y bits, octave, keypoint index} of cellItems[k]: a window scan reads one contiguous slice
    float minX, minY, invW, invH;
};

struct Ctx;

// Shared-memory carve of fast_cells_kernel for a set of cells (host-computed maxima).
struct FastSmem {
    int tileWords = 0, scrWords = 0, clistCap = 0, workCap = 0;   // tile/scr in rows (multiply by the pitch in use)
    int maxRowWords = 0;                                           // widest cell row in words (2*nw+3): selects the tile pitch
    int pitch() const { return maxRowWords <= FT_PITCH_SMALL ? FT_PITCH_SMALL : FT_PITCH; }
    size_t bytes() const { return sizeof(uint32_t) * ((size_t)(tileWords + scrWords) * pitch() + clistCap) + sizeof(uint16_t) * 2 * (size_t)workCap; }
};
// Shared-memory carve of fast_strip_kernel for a set of cell groups (host-computed maxima).
struct FastStripSmem {
    int tileRows = 0, scrRows = 0, segCap = 0, clistCap = 0;
    size_t bytes() const { return sizeof(uint32_t) * ((size_t)(tileRows + scrRows + 5) * FS_PITCH + 8 + clistCap) + sizeof(uint16_t) * FS_WARPS * (size_t)segCap; }
};
// Append one group of `cells` horizontally adjacent FAST cells (3 x int4) and grow `need`: [x0,x1) x [y0,y1) is the image of the
// whole group in level coordinates (the first cell's iniX .. the last cell's maxX), cells are wCell wide.
void push_fast_group(std::vector<int4>& groups, FastStripSmem& need, int x0, int y0, int x1, int y1, int level, int cells, int wCell,
                     unsigned levelOff, int pitch, unsigned candOff, int candCap);
void launch_fast_strips(Ctx& c, const uint8_t* d_pyr, unsigned pyrBytes, unsigned candPerImg, int minTh, int iniTh,
                        const int4* d_groups, int nGroups, const FastStripSmem& need, uint32_t* d_cand, int32_t* d_candCount, int n,
                        cudaStream_t stream = nullptr, bool pdl = false);
// Append one FAST cell (3 x int4) to a host cell table and grow `need`.  [x0,x1) x [y0,y1) is the cell image in level
// coordinates (3-pixel FAST margin included); candidates are only emitted inside [ex0,ex1) x [ey0,ey1) (ex1 == 0: anywhere).
void push_fast_cell(std::vector<int4>& cells, FastSmem& need, int x0, int y0, int x1, int y1, int level, unsigned levelOff, int pitch,
                    unsigned candOff, int candCap, int ex0 = 0, int ey0 = 0, int ex1 = 0, int ey1 = 0);
// passes: 2 = cv::FAST(iniTh) and, when the cell stays empty, cv::FAST(minTh) (ORBextractor); 1 = iniTh only
void launch_fast_cells(Ctx& c, const uint8_t* d_pyr, unsigned pyrBytes, unsigned candPerImg, int minTh, int iniTh, int passes,
                       const int4* d_cells, int nCells, const FastSmem& need, uint32_t* d_cand, int32_t* d_candCount, int n);

void bird_destroy(Ctx& c);   // bird.cu
// bird.cu: the birdview front-end on device-resident images, for the batched frame step (api.cu)
struct BirdStepView {
    const orbb200_kp_t* d_kps; const uint8_t* d_desc; const int32_t* d_counts; int kpPerImg;      // mvKeysBird / mDescriptorsBird per image
    const float* d_qx; const float* d_qy; const float* d_qangle; const int32_t* d_qlevel; const uint8_t* d_qvalid;   // [n+1][kpPerImg], slot 0 = carried frame
    const uint8_t* d_carryDesc;
    void* plan;
};
int bird_set_mask(Ctx& c, int w, int h, int nfeatures, int batch, const uint8_t* mask, size_t stride);
int bird_step_view(Ctx& c, int w, int h, int nfeatures, int n, BirdStepView* out);
int bird_step_enqueue(Ctx& c, int w, int h, int nfeatures, int n, const uint8_t* d_imgs, size_t imgBytes, size_t stride, bool chain, BirdStepView* out);
int bird_step_carry(Ctx& c, const BirdStepView& v, int n);
int bird_step_status(Ctx& c);

// ---- programmatic dependent launch (PDL) ----
// The extraction is a chain of short dependent kernels; launched with cudaLaunchAttributeProgrammaticStreamSerialization a kernel's
// CTAs are scheduled while its predecessor drains, run their prologue (table loads, shared-memory set-up) and block in
// griddepcontrol.wait until the predecessor grid has completed and its writes are visible.  Rule kept by every kernel launched this
// way: nothing that another kernel of the chain writes or reads is touched before pdl_wait() (host-uploaded tables are fair game).
// Second rule, learnt the hard way (profiles/r2c_summary.md): a programmatic edge is only used where every buffer the successor reads was
// written ONCE, before anybody in the chain read it (pyramid levels, candidate slots, partial results).  Buffers that one kernel reads and
// a later one partly rewrites (the birdview corner arrays through cornerSubPix, the border bytes next to pixels) keep full dependencies:
// an SM's L1 may still hold the older line, and only a kernel boundary is documented to drop it.
// Without the attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// ---- kernels launchers (extract.cu) ----
// Level ranges [l0, l1) and the stream are explicit so that, for one or two images, level 0 (no resize needed) can run FAST + octree
// on a side branch of the graph while the other levels are still being built.
void launch_import(Ctx& c, const uint8_t* d_imgs, size_t img_bytes, size_t stride, int n);
void launch_resizes(Ctx& c, int n, cudaStream_t stream, uint8_t* hostPyr = nullptr);
void launch_border(Ctx& c, int n, cudaStream_t stream, bool afterKernel);
void launch_pyramid(Ctx& c, int n, uint8_t* hostPyr = nullptr);
void launch_blur(Ctx& c, int n, cudaStream_t stream, bool afterKernel = false);
void launch_clear_counters(Ctx& c, int n);       // zeroes the per-level candidate counters: once per extraction, before any FAST launch
void launch_fast(Ctx& c, int n, bool afterKernel = false);
void launch_fast_levels(Ctx& c, int n, int l0, int l1, cudaStream_t stream, bool afterKernel);
void launch_octree(Ctx& c, int n);
void launch_octree_levels(Ctx& c, int n, int l0, int l1, cudaStream_t stream, bool afterKernel);
// Pinned (device-mapped) result block of a small host call: describe_kernel writes the final records there as well, so that no
// copy-engine operation follows the graph.  All null: device pools only.
struct HostMirror { orbb200_kp_t* kps; uint8_t* desc; int32_t* counts; int32_t* status; const int32_t* d_status; };
// A small host call staged in pinned memory: n images of rowBytes-pitched rows in, results out through `mirror`.
struct HostStage { const uint8_t* imgs; size_t imgBytes; int rowBytes; HostMirror mirror; uint8_t* hostPyr; };   // hostPyr: pinned mirror of image 0's pyramid block or null
void launch_describe(Ctx& c, int n, bool afterKernel = false, const HostMirror* mirror = nullptr);
void launch_import_host(Ctx& c, const uint8_t* h_imgs, size_t imgBytes, int rowBytes, int n, uint8_t* hostPyr = nullptr);
size_t octree_smem_bytes(int maxNodes, int smemCand = 0);
void launch_stereo(Ctx& c, int n_frames, int left0, int right0, int strideImgs, float mb, float mbf, const float* d_invScale, int32_t* d_nKept);

}  // namespace orbb200
