// Extraction kernels: pyramid resize, 7x7 Gaussian blur, per-cell FAST-9 with threshold fallback,
// octree distribution, intensity-centroid orientation and rBRIEF descriptors.
//
// Integer arithmetic follows OpenCV 4.x exactly (DESIGN.md
// and SURVEY.md Appendix A lists each primitive); float arithmetic uses explicit round-to-nearest intrinsics so that nvcc
// never contracts it into FMAs.
#include <cuda_fp16.h>

#include "ctx.cuh"
#include "device_math.cuh"

namespace orbb200 {

// ---------------------------------------------------------------------------------------------------
// import: user images (arbitrary stride) -> level 0 of the pyramid pool (pitch 128-aligned)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) import_kernel(const uint8_t* __restrict__ src, size_t img_bytes, size_t stride,
                                                     uint8_t* __restrict__ pyr, unsigned pyrBytes, unsigned off0, int w, int h, int pitch)
{
    const int img = blockIdx.z;
    const int y = blockIdx.y;
    const uint8_t* s = src + (size_t)img * img_bytes + (size_t)y * stride;
    uint8_t* d = pyr + (size_t)img * pyrBytes + off0 + (size_t)y * pitch;
    const int x16 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (x16 >= w) return;
    // source rows are rarely aligned (stride 1241): read the aligned words around the 16 source bytes and
    // funnel-shift; words that hold no byte of this row are not touched (the last one could lie past the allocation)
    const uintptr_t a = reinterpret_cast<uintptr_t>(s + x16);
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const int mis = (int)(a & 3), sh = mis * 8;
    const int avail = w - x16 + mis;                      // bytes of this row from wp[0] on
    uint32_t v[5];
#pragma unroll
    for (int k = 0; k < 5; k++) v[k] = (4 * k < avail && (k < 4 || mis != 0)) ? __ldg(wp + k) : 0u;
    uint4 o;
    o.x = __funnelshift_r(v[0], v[1], sh); o.y = __funnelshift_r(v[1], v[2], sh);
    o.z = __funnelshift_r(v[2], v[3], sh); o.w = __funnelshift_r(v[3], v[4], sh);
    *reinterpret_cast<uint4*>(d + x16) = o;               // 128-byte pitched rows: the row padding absorbs the tail
}

// The same for a small host call: the images sit in the context's pinned staging block (rows padded to 16 bytes by the host
// memcpy) and this kernel reads them over PCIe itself, 16 bytes per lane, so that the upload is a node of the extraction graph
// instead of a copy-engine operation in front of it.
// hostPyr != nullptr: image 0's levels are also stored into a pinned mirror of its pyramid block (the drop-in's mvImagePyramid), by the
// kernels that produce them: posted 16-byte / 4-byte stores over PCIe beside the device stores, no copy afterwards.
__global__ void __launch_bounds__(128) import_host_kernel(const uint8_t* __restrict__ src, unsigned imgBytes, int rowVec, int nVec,
                                                          uint8_t* __restrict__ pyr, unsigned pyrBytes, unsigned off0, int pitch, uint8_t* __restrict__ hostPyr)
{
    const int img = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    if (i >= nVec) return;
    const int r = i / rowVec, k = i - r * rowVec;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + (size_t)img * imgBytes) + i);
    *reinterpret_cast<uint4*>(pyr + (size_t)img * pyrBytes + off0 + (size_t)r * pitch + 16 * k) = v;     // the row padding absorbs the tail
    if (hostPyr && img == 0) *reinterpret_cast<uint4*>(hostPyr + off0 + (size_t)r * pitch + 16 * k) = v;
}

// ---------------------------------------------------------------------------------------------------
// cv::resize INTER_LINEAR u8 (reference src/ORBextractor.cc:1120): level l-1 -> l.
// 11-bit fixed-point coefficients from host-built tables; vertical pass
// (((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2.
// A CTA owns 128 output columns x RS_ROWS output rows and stages their source footprint in shared memory; each
// lane owns 4 adjacent output columns, keeps their column coefficients in registers, walks down 8 rows and reuses
// the horizontal interpolation of a source row for the next output row (a source row serves ~1.7 output rows
// at 1/1.2).  (Tried: a separable two-phase form -- every H row once into shared memory, then the vertical pass --
// with fewer instructions per pixel: 0.44 -> 0.53 ms per step, slower.)
// ---------------------------------------------------------------------------------------------------
#ifndef ORBB200_RS_THREADS
#define ORBB200_RS_THREADS 64
#endif
constexpr int RS_THREADS = ORBB200_RS_THREADS;  // one CTA per 128 x RS_ROWS output tile; warp w owns RS_RPW consecutive rows of it
constexpr int RS_RPW = RS_ROWS / (RS_THREADS / 32);

__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t saddr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

// prmt.b32 with a selector whose nibbles are all < 8 (no sign-replication mode): spares the "& 0x7777" __byte_perm adds
__device__ __forceinline__ uint32_t prmt_raw(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

template <bool NARROW>
__global__ void __launch_bounds__(RS_THREADS) resize_kernel(uint8_t* __restrict__ pyr, unsigned pyrBytes, LevelGeom src, LevelGeom dst,
                                                            const int2* __restrict__ xtab, const int4* __restrict__ ytab,
                                                            const int4* __restrict__ tiles, int nTiles, int smemPitch, int smemRows,
                                                            const CUtensorMap* __restrict__ srcMap, uint8_t* __restrict__ hostPyr)
{
    // The source footprint of the tile is staged in shared memory, so the interpolation reads never wait on global
    // memory: one TMA box (cp.async.bulk.tensor.3d of smemPitch x smemRows bytes from the source level, its first column
    // is 16-byte aligned by construction) when the box fits TMA's 256-byte inner limit (srcMap != nullptr), else
    // coalesced 16-byte loads.  Every table entry the CTA needs (tile window, the rows' vertical coefficients, the lane's
    // four column entries) is requested before the wait.
    extern __shared__ __align__(128) uint8_t rsSmem[];
    __shared__ int4 sY[RS_ROWS];
    __shared__ __align__(8) uint64_t sBar;
    pdl_launch_dependents();
    const int4 t = __ldg(tiles + 2 * blockIdx.x);       // {x0, y0, first staged source row, rows}
    const int4 u = __ldg(tiles + 2 * blockIdx.x + 1);   // {first staged source column, vectors per row, 2^16/vectors + 1, -}
    const int img = blockIdx.y;
    const int x0 = t.x, y0 = t.y, ry0 = t.z, nrows = t.w;              // rows <= smemRows, vectors <= smemPitch/16 (host-sized)
    const int cx0 = u.x, nvec = u.y;
    const int tid = threadIdx.x;
    const uint8_t* S = pyr + (size_t)img * pyrBytes + src.off;
    uint8_t* D = pyr + (size_t)img * pyrBytes + dst.off;
    uint8_t* HD = (hostPyr && img == 0) ? hostPyr + dst.off : nullptr;      // pinned mirror of image 0's pyramid block (see import_host_kernel)
    const int x4 = x0 + 4 * (tid & 31);
    int2 xt[4];
#pragma unroll
    for (int i = 0; i < 4; i++) xt[i] = __ldg(xtab + dst.xtabOff + min(x4 + i, dst.w - 1));
    if (tid < RS_ROWS) sY[tid] = __ldg(ytab + dst.ytabOff + min(y0 + tid, dst.h - 1));
    pdl_wait();                                                           // the source level is complete
    if (srcMap != nullptr) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&sBar);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(smemPitch * smemRows) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         :: "r"((uint32_t)__cvta_generic_to_shared(rsSmem)), "l"(reinterpret_cast<uint64_t>(srcMap)), "r"(cx0), "r"(ry0), "r"(img), "r"(bar)
                         : "memory");
        }
        __syncthreads();                                                  // the barrier is initialised (and sY written) for everyone
        asm volatile("{\n\t.reg .pred p;\n\tRSWAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra RSDONE;\n\tbra RSWAIT;\n\tRSDONE:\n\t}"
                     :: "r"(bar), "r"(0) : "memory");
    } else {
        // floor(i / nvec) by a 16-bit reciprocal: exact while i * nvec < 2^16 (i < 64 rows x 16 vectors)
        const unsigned rcp = (unsigned)u.z;
        const uint8_t* Sw = S + (size_t)ry0 * src.pitch + cx0;
        for (int i = tid; i < nrows * nvec; i += RS_THREADS) {
            const int r = (int)(((unsigned)i * rcp) >> 16), k = i - r * nvec;
            // rows are 128-byte pitched and start 32 bytes into the pitch: column multiples of 16 are 16-byte aligned
            *reinterpret_cast<uint4*>(rsSmem + r * smemPitch + 16 * k) =
                *reinterpret_cast<const uint4*>(Sw + (unsigned)(r * src.pitch + 16 * k));
        }
        __syncthreads();
    }
    // warp index through a shuffle: the output row and everything derived from it is warp-uniform for the compiler
    const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (x4 >= dst.w) return;
    // column setup, once per thread.  NARROW (scale <= 2: the taps of 4 adjacent output columns lie within 8 source
    // bytes): one row address, three aligned word loads funnel-shifted to an 8-byte window that starts at the first
    // tap, then per column one PRMT (both taps as bytes 0/1) and one IDP2A with the coefficient pair.
    // General form: two byte loads and two multiplies per column.  Shared memory is addressed through 32-bit
    // shared-space addresses (one add per source row instead of a generic-pointer rebuild).
    const uint32_t smemBase = (uint32_t)__cvta_generic_to_shared(rsSmem);
    uint32_t o0[4], o1[4], sel[4], cf[4];
    const int sx0 = xt[0].x - cx0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int s0 = xt[i].x - cx0, s1 = min(xt[i].x + 1, src.w - 1) - cx0;
        o0[i] = smemBase + s0; o1[i] = smemBase + s1;
        sel[i] = (uint32_t)((s0 - sx0) | ((s1 - sx0) << 4)) | 0x4400u;    // bytes 2,3 <- byte 4 (unused by IDP2A.LO)
        cf[i] = (uint32_t)xt[i].y;                                        // a0 | a1 << 16
    }
    const uint32_t pw = smemBase + (sx0 & ~3);
    const int sh = (sx0 & 3) * 8;
    auto hrow = [&](int sy, uint32_t (&h)[4]) {
        const uint32_t ro = (uint32_t)((sy - ry0) * smemPitch);
        if (NARROW) {
            const uint32_t w0 = lds_u32(pw + ro), w1 = lds_u32(pw + ro + 4), w2 = lds_u32(pw + ro + 8);   // pitch has >= 16 spare bytes
            const uint32_t W0 = __funnelshift_r(w0, w1, sh), W1 = __funnelshift_r(w1, w2, sh);
#pragma unroll
            for (int i = 0; i < 4; i++) h[i] = __dp2a_lo(cf[i], prmt_raw(W0, W1, sel[i]), 0u) >> 4;
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) h[i] = (lds_u8(o0[i] + ro) * (cf[i] & 0xffffu) + lds_u8(o1[i] + ro) * (cf[i] >> 16)) >> 4;
        }
    };
    uint32_t ha[4], hb[4];
    int ia = -1, ib = -1;
    const int yb = y0 + RS_RPW * wid, ye = min(yb + RS_RPW, dst.h);
    unsigned doff = (unsigned)(yb * dst.pitch + x4);
    for (int y = yb; y < ye; y++) {
        const int4 yt = sY[y - y0];                                       // {sy0, sy1, b0, b1}: warp-uniform
        if (yt.x != ia) {
            if (yt.x == ib) {
#pragma unroll
                for (int i = 0; i < 4; i++) ha[i] = hb[i];
            } else {
                hrow(yt.x, ha);
            }
            ia = yt.x;
        }
        if (yt.y != ib) {
            if (yt.y == ia) {
#pragma unroll
                for (int i = 0; i < 4; i++) hb[i] = ha[i];
            } else {
                hrow(yt.y, hb);
            }
            ib = yt.y;
        }
        // (((b0*H0)>>16) + ((b1*H1)>>16) + 2) >> 2, the >>16 as the high word of (b << 16) * H
        const uint32_t bz = (uint32_t)yt.z << 16, bw = (uint32_t)yt.w << 16;
        uint32_t v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) v[i] = (__umulhi(bz, ha[i]) + __umulhi(bw, hb[i]) + 2u) >> 2;     // <= 255
        const uint32_t out = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
        *reinterpret_cast<uint32_t*>(D + doff) = out;                     // row padding absorbs the tail
        if (HD) *reinterpret_cast<uint32_t*>(HD + doff) = out;
        doff += (unsigned)dst.pitch;
    }
}

// Reflect-101 border of every level: PYR_MARGIN_X(>=4 used) columns left, 8 right, 3 rows above/below.
// blockIdx.x = level * BD_CHUNKS + chunk: the (few thousand) border bytes of a level are split over BD_CHUNKS CTAs.
#ifndef ORBB200_BD_CHUNKS
#define ORBB200_BD_CHUNKS 32
#endif
constexpr int BD_CHUNKS = ORBB200_BD_CHUNKS;

__global__ void __launch_bounds__(128) border_kernel(uint8_t* __restrict__ pyr, unsigned pyrBytes, Geom g)
{
    const int level = blockIdx.x / BD_CHUNKS, chunk = blockIdx.x - level * BD_CHUNKS, img = blockIdx.y;
    const LevelGeom L = g.lv[level];
    pdl_launch_dependents();
    pdl_wait();
    if (L.w < 8 || L.h < 4) return;
    uint8_t* B = pyr + (size_t)img * pyrBytes + L.off;
    auto rx = [&](int x) { return x < 0 ? -x : (x >= L.w ? 2 * (L.w - 1) - x : x); };
    auto ry = [&](int y) { return y < 0 ? -y : (y >= L.h ? 2 * (L.h - 1) - y : y); };
    // rows -3..-1 and h..h+2, columns -4 .. w+7
    const int wide = L.w + 12;
    const int nA = 6 * wide, nB = 12 * L.h;
    for (int i = chunk * 128 + threadIdx.x; i < nA + nB; i += BD_CHUNKS * 128) {
        int x, y;
        if (i < nA) {
            const int k = i / wide;
            x = i - k * wide - 4;
            y = k < 3 ? k - 3 : L.h + (k - 3);
        } else {
            const int j = i - nA;
            y = j / 12;
            const int k = j - y * 12;
            x = k < 4 ? k - 4 : L.w + (k - 4);
        }
        B[(ptrdiff_t)y * L.pitch + x] = B[(ptrdiff_t)ry(y) * L.pitch + rx(x)];
    }
}

// ---------------------------------------------------------------------------------------------------
// cv::GaussianBlur 7x7 sigma 2, BORDER_REFLECT_101 (reference src/ORBextractor.cc:1085-1086).
// Q8.8 kernel {18,34,48,56,48,34,18}; H in u16, V in u32, (v + 2^15) >> 16.
// One launch covers all levels: blockIdx.y indexes a flattened (level, tile-row) table.
// ---------------------------------------------------------------------------------------------------
// Streaming form: one warp owns a 128-column x BL_ROWS-row tile; each lane owns 4 adjacent columns (one output
// word), walks down the rows, keeps the horizontal sums of the last 7 rows in a register ring and emits one
// output word per row.  No shared memory, one launch for all levels (tiles come from a flattened table).
constexpr int BL_WARPS = 4;

__device__ __forceinline__ int reflect101(int p, int len)
{
    if (p < 0) p = -p;
    if (p >= len) p = 2 * (len - 1) - p;
    return min(max(p, 0), len - 1);    // second clamp only matters for levels narrower than the kernel
}

#ifndef ORBB200_BL_AHEAD
#define ORBB200_BL_AHEAD 6
#endif
// horizontal 7-tap sums of the 4 pixels in w1 (w0 = the 4 bytes before, w2 = the 4 after): ten 4-way byte dot products
// (IDP.4A) on the aligned words with the taps placed per output pixel (zeros where a word holds no tap of that pixel),
// no funnel shifts
__device__ __forceinline__ void blur_hrow(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t (&h)[4])
{
    constexpr uint32_t A0 = (18u << 8) | (34u << 16) | (48u << 24), B0 = 56u | (48u << 8) | (34u << 16) | (18u << 24);
    constexpr uint32_t A1 = (18u << 16) | (34u << 24), B1 = 48u | (56u << 8) | (48u << 16) | (34u << 24), C1 = 18u;
    constexpr uint32_t A2 = (18u << 24), B2 = 34u | (48u << 8) | (56u << 16) | (48u << 24), C2 = 34u | (18u << 8);
    constexpr uint32_t B3 = 18u | (34u << 8) | (48u << 16) | (56u << 24), C3 = 48u | (34u << 8) | (18u << 16);
    h[0] = __dp4a(w1, B0, __dp4a(w0, A0, 0u));
    h[1] = __dp4a(w2, C1, __dp4a(w1, B1, __dp4a(w0, A1, 0u)));
    h[2] = __dp4a(w2, C2, __dp4a(w1, B2, __dp4a(w0, A2, 0u)));
    h[3] = __dp4a(w2, C3, __dp4a(w1, B3, 0u));
}

__global__ void __launch_bounds__(BL_WARPS * 32) blur_kernel(const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur, unsigned pyrBytes,
                                                             Geom g, const int4* __restrict__ tiles, int nTiles)
{
    // warp index through a shuffle: lets the compiler treat everything derived from the tile (level, row base
    // pointers) as warp-uniform and address the loads as uniform base + 32-bit lane offset
    const int tileIdx = blockIdx.x * BL_WARPS + __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    pdl_launch_dependents();
    if (tileIdx >= nTiles) return;
    const int lane = threadIdx.x & 31;
    const int4 t = __ldg(tiles + tileIdx);          // {level, x0, y0, -}
    pdl_wait();
    const LevelGeom L = g.lv[t.x];
    const int img = blockIdx.y;
    const int x = t.y + 4 * lane, y0 = t.z;
    if (x >= L.w) return;
    const uint8_t* S = pyr + (size_t)img * pyrBytes + L.off;
    uint8_t* D = blur + (size_t)img * pyrBytes + L.off;
    const int rows = min(BL_ROWS, L.h - y0);

    // the level is stored with its reflect-101 border (border_kernel): no edge handling here; rows are
    // visited in order, so the row pointers just advance by the pitch
    const uint32_t* rp = reinterpret_cast<const uint32_t*>(S + (ptrdiff_t)(y0 - 3) * L.pitch + x);
    uint32_t* wp = reinterpret_cast<uint32_t*>(D + (size_t)y0 * L.pitch + x);
    const int pitchW = L.pitch >> 2;
    auto fetch = [&](uint32_t (&w)[3]) {
        w[0] = rp[-1]; w[1] = rp[0]; w[2] = rp[1];
        rp += pitchW;
    };
    auto load_row = [&](uint32_t (&h)[4]) {
        uint32_t w[3];
        fetch(w);
        blur_hrow(w[0], w[1], w[2], h);
    };
    // Vertical pass with 2-way dot products (IDP.2A): the horizontal sums (<= 65280) of rows r and r+1 sit in one
    // register as u16 pairs P_r; window rows r..r+6 = P_r.(18,34) + P_{r+2}.(48,56) + P_{r+4}.(48,34) + 18 * H_{r+6}.
    // The ring holds the six newest pairs; it rotates by renaming (unrolled x6).
    // Memory-level parallelism: a warp that loads a row, uses it and only then loads the next keeps ~136 bytes in
    // flight, and 40 such warps per SM cover ~1.1 TB/s of reads at HBM latency -- what the kernel measured.  The raw
    // words of the next three rows are therefore requested three steps ahead (ring of 3, same renaming).
    constexpr uint32_t KV = 18u | (34u << 8) | (48u << 16) | (56u << 24);       // lo: (18,34)  hi: (48,56)
    constexpr uint32_t KW = 48u | (34u << 8);                                   // lo: (48,34)
    uint32_t p0[4], p1[4], p2[4], p3[4], p4[4], p5[4], hl[4], hn[4];
    auto pack = [&](const uint32_t (&a)[4], const uint32_t (&b)[4], uint32_t (&p)[4]) {
#pragma unroll
        for (int k = 0; k < 4; k++) p[k] = __byte_perm(a[k], b[k], 0x5410);
    };
    {
        uint32_t h0[4], h1[4], h2[4], h3[4], h4[4];
        load_row(h0); load_row(h1); load_row(h2); load_row(h3); load_row(h4); load_row(hl);
        pack(h0, h1, p0); pack(h1, h2, p1); pack(h2, h3, p2); pack(h3, h4, p3); pack(h4, hl, p4);
    }
#if ORBB200_BL_AHEAD == 6
    uint32_t ra[3] = {0, 0, 0}, rb[3] = {0, 0, 0}, rc[3] = {0, 0, 0}, rd[3] = {0, 0, 0}, re[3] = {0, 0, 0}, rf[3] = {0, 0, 0};
    fetch(ra);                                      // row y0 + 3 (rows >= 1)
    if (rows > 1) fetch(rb);
    if (rows > 2) fetch(rc);
    if (rows > 3) fetch(rd);
    if (rows > 4) fetch(re);
    if (rows > 5) fetch(rf);
#define ORBB200_BLUR_NEXT(R) blur_hrow(R[0], R[1], R[2], hn); if (r + 6 < rows) fetch(R);
#elif ORBB200_BL_AHEAD == 3
    uint32_t ra[3] = {0, 0, 0}, rb[3] = {0, 0, 0}, rc[3] = {0, 0, 0};
    uint32_t (&rd)[3] = ra, (&re)[3] = rb, (&rf)[3] = rc;
    fetch(ra);                                      // row y0 + 3 (rows >= 1)
    if (rows > 1) fetch(rb);
    if (rows > 2) fetch(rc);
#define ORBB200_BLUR_NEXT(R) blur_hrow(R[0], R[1], R[2], hn); if (r + 3 < rows) fetch(R);
#else
    uint32_t ra[3];
    uint32_t (&rb)[3] = ra, (&rc)[3] = ra, (&rd)[3] = ra, (&re)[3] = ra, (&rf)[3] = ra;
#define ORBB200_BLUR_NEXT(R) fetch(R); blur_hrow(R[0], R[1], R[2], hn);
#endif
    // state before a step: pairs A..E = P_r..P_{r+4} (rows r..r+5), hl = H_{r+5}; R = raw words of row r+6
#define ORBB200_BLUR_STEP(A, B, C, Dd, E, F, R) \
    if (r < rows) { \
        ORBB200_BLUR_NEXT(R) \
        uint32_t out = 0; \
        _Pragma("unroll") for (int k = 0; k < 4; k++) { \
            uint32_t v = __dp2a_lo(A[k], KV, 18u * hn[k] + 32768u); \
            v = __dp2a_hi(C[k], KV, v); \
            v = __dp2a_lo(E[k], KW, v); \
            out |= (v >> 16) << (8 * k); \
            F[k] = __byte_perm(hl[k], hn[k], 0x5410); \
            hl[k] = hn[k]; \
        } \
        *wp = out; wp += pitchW; r++; \
    }
    for (int r = 0; r < rows;) {
        ORBB200_BLUR_STEP(p0, p1, p2, p3, p4, p5, ra)
        ORBB200_BLUR_STEP(p1, p2, p3, p4, p5, p0, rb)
        ORBB200_BLUR_STEP(p2, p3, p4, p5, p0, p1, rc)
        ORBB200_BLUR_STEP(p3, p4, p5, p0, p1, p2, rd)
        ORBB200_BLUR_STEP(p4, p5, p0, p1, p2, p3, re)
        ORBB200_BLUR_STEP(p5, p0, p1, p2, p3, p4, rf)
    }
#undef ORBB200_BLUR_STEP
#undef ORBB200_BLUR_NEXT
}

// ---------------------------------------------------------------------------------------------------
// Grid FAST (reference src/ORBextractor.cc:789-829 + cv::FAST TYPE_9_16 with NMS).
// One CTA per 30-px cell.  The threshold-independent score S = M-1 (M = best 9-arc contrast) is computed
// for the cell's inner rectangle with two horizontally adjacent pixels per thread packed as s16x2
// (VIMNMX3.S16x2); 3x3 strict-greater NMS inside the rectangle; a local maximum is a candidate if
// S >= iniThFAST, or, when the cell has none, if S >= minThFAST.
// Candidates are appended unordered to the level's pool: the octree only needs (x, y, response).
// ---------------------------------------------------------------------------------------------------
#ifndef ORBB200_FT_THREADS
#define ORBB200_FT_THREADS 64
#endif
#ifndef ORBB200_FT_MINBLK
#define ORBB200_FT_MINBLK 10
#endif
constexpr int FT_THREADS = ORBB200_FT_THREADS;

// (Tried: funnel shifts as two IMADs on the FMA pipe instead of one SHF on the ALU pipe -- slower, the kernel is then
// issue-bound: 0.93 -> 0.96 ms per 128 images.)
__device__ __forceinline__ uint32_t min3s(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_s16x2(a, b, c); }
__device__ __forceinline__ uint32_t max3s(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
// The same per-lane min/max as half2 ops (HMNMX2/VHMNMX): lanes are 0x6400 + byte, i.e. the positive normal
// halves 1024..1279, whose order as halves equals their order as integers, so the results are identical.
// Measured on B200 (tools/ubench/pipes.cu): HMNMX2 issues to the same ALU pipe as VIMNMX (interleaving them
// takes the sum of both times, 64 lanes/clk/SM either way; VIMNMX3 costs the same as VIMNMX; IMAD overlaps
// fully), so splitting the network over both instruction kinds buys nothing: FT_NHALF stays 0.
__device__ __forceinline__ uint32_t hmin2u(uint32_t a, uint32_t b)
{
    const __half2 r = __hmin2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ uint32_t hmax2u(uint32_t a, uint32_t b)
{
    const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
}
constexpr int FT_NHALF = 0;        // first-stage triples computed with HMNMX2/VHMNMX (0..16); see note above
constexpr uint32_t FT_BIAS = 0x64006400u;

// score of the two pixels packed in `c` given the 16 ring pairs; lanes hold u8 values.
// With d_k = v - p_k:  M_dark = max_arcs min_arc d = v - min_arcs max_arc p,  M_bright = max_arcs min_arc (-d)
// = max_arcs min_arc p - v, so the min/max network runs on the ring values themselves and the centre is
// subtracted once at the end.
__device__ __forceinline__ uint32_t fast_score_pair(uint32_t c, const uint32_t (&r)[16])
{
    uint32_t lo3[16], hi3[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (k < FT_NHALF) {
            lo3[k] = hmin2u(hmin2u(r[k], r[(k + 1) & 15]), r[(k + 2) & 15]);
            hi3[k] = hmax2u(hmax2u(r[k], r[(k + 1) & 15]), r[(k + 2) & 15]);
        } else {
            lo3[k] = min3s(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
            hi3[k] = max3s(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
        }
    }
    uint32_t mn[16], mx[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        mn[k] = min3s(lo3[k], lo3[(k + 3) & 15], lo3[(k + 6) & 15]);   // min over the 9-arc starting at k
        mx[k] = max3s(hi3[k], hi3[(k + 3) & 15], hi3[(k + 6) & 15]);
    }
    uint32_t a = max3s(mn[0], mn[1], mn[2]);
    a = max3s(a, mn[3], mn[4]); a = max3s(a, mn[5], mn[6]); a = max3s(a, mn[7], mn[8]);
    a = max3s(a, mn[9], mn[10]); a = max3s(a, mn[11], mn[12]); a = max3s(a, mn[13], mn[14]);
    a = __vmaxs2(a, mn[15]);                        // max_arcs min_arc p  = v + M_bright
    uint32_t b = min3s(mx[0], mx[1], mx[2]);
    b = min3s(b, mx[3], mx[4]); b = min3s(b, mx[5], mx[6]); b = min3s(b, mx[7], mx[8]);
    b = min3s(b, mx[9], mx[10]); b = min3s(b, mx[11], mx[12]); b = min3s(b, mx[13], mx[14]);
    b = __vmins2(b, mx[15]);                        // min_arcs max_arc p  = v - M_dark
    // S + 512 = max(M_dark - 1, M_bright - 1, 0) + 512; per-lane biases keep every lane positive (no borrow)
    const uint32_t dark = (c + 0x01ff01ffu) - b;    // v + 511 - (v - M_dark)  = M_dark - 1 + 512
    const uint32_t bright = (a + 0x01ff01ffu) - c;  // v + M_bright + 511 - v = M_bright - 1 + 512
    const uint32_t s = max3s(dark, bright, 0x02000200u);
    return s - 0x02000200u;
}

// The 16 ring pairs around the pixel pair whose centre word is t[0] (tile pitch FT_PITCH).
// ring order k=0..15: (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)(-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)
template <int P>
__device__ __forceinline__ void fast_load_ring(const uint32_t* t, uint32_t (&r)[16])
{
    {
        const uint32_t* p = t + 3 * P;                          // dy = +3 : dx -1,0,1
        const uint32_t wl = p[-1], wc = p[0], wr = p[1];
        r[15] = __funnelshift_r(wl, wc, 16); r[0] = wc; r[1] = __funnelshift_r(wc, wr, 16);
    }
    {
        const uint32_t* p = t - 3 * P;                          // dy = -3
        const uint32_t wl = p[-1], wc = p[0], wr = p[1];
        r[9] = __funnelshift_r(wl, wc, 16); r[8] = wc; r[7] = __funnelshift_r(wc, wr, 16);
    }
    r[14] = t[2 * P - 1]; r[2] = t[2 * P + 1];                  // dy=+2: dx -2, +2
    r[10] = t[-2 * P - 1]; r[6] = t[-2 * P + 1];                // dy=-2
    {
        const uint32_t* p = t + P;                              // dy = +1 : dx -3, +3
        r[13] = __funnelshift_r(p[-2], p[-1], 16); r[3] = __funnelshift_r(p[1], p[2], 16);
    }
    {
        const uint32_t* p = t - P;                              // dy = -1
        r[11] = __funnelshift_r(p[-2], p[-1], 16); r[5] = __funnelshift_r(p[1], p[2], 16);
    }
    r[12] = __funnelshift_r(t[-2], t[-1], 16); r[4] = __funnelshift_r(t[1], t[2], 16);   // dy = 0
}

// Necessary condition for score >= thr on either pixel of the pair: a 9-arc contains one pixel of every
// opposite ring pair (k, k+8), so a bright arc needs max(p_k, p_k+8) > v + thr for all k and a dark arc needs
// min(p_k, p_k+8) < v - thr.  Tested on the three opposite pairs whose words are aligned with the centre
// word (k = 0, 2, 6: no funnel shifts): 7 loads + ~13 ALU ops instead of the ~130-instruction network.
// Tp = (thr + 1) in both lanes.  Returns non-zero when a pixel of the pair may reach thr.
template <int P>
__device__ __forceinline__ uint32_t fast_may_pass(const uint32_t* t, uint32_t Tp)
{
    const uint32_t c = t[0];
    const uint32_t r0 = t[3 * P], r8 = t[-3 * P];
    const uint32_t r2 = t[2 * P + 1], r10 = t[-2 * P - 1];
    const uint32_t r6 = t[-2 * P + 1], r14 = t[2 * P - 1];
    const uint32_t A = min3s(__vmaxs2(r0, r8), __vmaxs2(r2, r10), __vmaxs2(r6, r14));   // bright: A >= v + thr + 1
    const uint32_t B = max3s(__vmins2(r0, r8), __vmins2(r2, r10), __vmins2(r6, r14));   // dark:   B <= v - thr - 1
    // per-lane >= via bit 15 of a biased difference; every lane stays within [0x8000 - 511, 0x8000 + 255]
    const uint32_t d1 = (A | 0x80008000u) - (c + Tp);
    const uint32_t d2 = ((c | 0x80008000u) - Tp) - B;
    return (d1 | d2) & 0x80008000u;
}

constexpr int FT_WARPS = FT_THREADS / 32;

// entry j of a list kept as FT_WARPS per-warp segments of `seg` slots holding n[0..FT_WARPS) entries
__device__ __forceinline__ int seg_list_at(const uint16_t* list, int seg, const int (&n)[FT_WARPS], int j)
{
    int w = 0;
#pragma unroll
    for (int k = 0; k < FT_WARPS - 1; k++)
        if (j >= n[k]) { j -= n[k]; w++; } else break;
    return list[w * seg + j];
}

template <int P>
__global__ void __launch_bounds__(FT_THREADS, ORBB200_FT_MINBLK) fast_cells_kernel(const uint8_t* __restrict__ pyr, unsigned pyrBytes, unsigned candPerImg,
                                                                   int minTh, int iniTh, const int4* __restrict__ cells,
                                                                   uint32_t* __restrict__ cand, int32_t* __restrict__ candCount,
                                                                   int tileWords, int scrWords, int clistCap, int workCap, int passes)
{
    // tile[r][1+m] = pixels (2m, 2m+1) of cell-image row r as u16x2; score tile in the same layout with a
    // zero row above/below.  Row pitch == pairs-per-row (mod 32): the flattened (row, pair) -> lane mapping
    // then walks consecutive banks across row boundaries (no bank conflicts).
    // Shared memory is sized by the host for the largest cell of this image shape (typically ~19 KB).
    extern __shared__ uint32_t ftSmem[];
    uint32_t* tile = ftSmem;
    uint32_t* scr = tile + tileWords * P;
    uint32_t* clist = scr + scrWords * P;
    uint16_t* wlist = reinterpret_cast<uint16_t*>(clist + clistCap);    // pairs that may hold a local maximum
    uint16_t* qlist = wlist + workCap;                                  // pairs that pass the quick test
    __shared__ int sN, sBase, sQn[FT_WARPS], sWn[FT_WARPS];

    const int img = blockIdx.y;
    const int4 cell = __ldg(cells + 3 * blockIdx.x);         // {x0|y0<<16, x1|y1<<16, level, -}
    const int4 lvl = __ldg(cells + 3 * blockIdx.x + 1);      // {level byte offset, pitch, candOff, candCap}
    const int4 mg = __ldg(cells + 3 * blockIdx.x + 2);       // {2^32/nw + 1, 2^32/npr + 1, emit x0|y0<<16, emit x1|y1<<16}
    const int ex0 = mg.z & 0xffff, ey0 = mg.z >> 16, ex1 = mg.w ? (mg.w & 0xffff) : 0x7fff, ey1 = mg.w ? (mg.w >> 16) : 0x7fff;
    const int x0 = cell.x & 0xffff, y0 = cell.x >> 16, x1 = cell.y & 0xffff, y1 = cell.y >> 16;
    const int level = cell.z;
    const int pitch = lvl.y;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int xa = x0 & ~3;
    const int tw = x1 - xa, th = y1 - y0;
    const int wi = x1 - x0 - 6, hi = th - 6;
    if (wi <= 0 || hi <= 0) return;
    const int cx0 = x0 - xa + 3, cx1 = cx0 + wi;    // inner columns in tile coordinates
    const int m0 = cx0 >> 1, m1 = (cx1 - 1) >> 1, npr = m1 - m0 + 1;
    const int nw = (tw + 3) >> 2;                   // 32-bit words loaded per row
    // Row pitch of both shared tiles: 49 words.  With the usual 16-18 pairs per row the flattened
    // (row, pair) -> lane mapping then touches (almost) disjoint banks for the rows a warp spans, and a
    // compile-time pitch turns every ring offset into an immediate.  (P = 25 for the usual <= 44-pixel cells: more
    // 2-way conflicts between the rows a warp spans, half the shared memory -- measured faster.)

    // zero frame of the score tile: the rows above/below and the words left/right of the inner span
    for (int i = tid; i < P; i += FT_THREADS) { scr[i] = 0; scr[(hi + 1) * P + i] = 0; }
    for (int i = tid; i < hi; i += FT_THREADS) { scr[(i + 1) * P + m0] = 0; scr[(i + 1) * P + m1 + 2] = 0; }
    // words next to the loaded span are read by masked lanes only, but must hold in-range values: a lane
    // outside [0x6400, 0x64ff] would borrow into its neighbour lane in the packed subtraction (or be a NaN)
    for (int r = tid; r < th; r += FT_THREADS) {
        uint32_t* t = tile + r * P;
        t[0] = FT_BIAS; t[2 * nw + 1] = FT_BIAS; t[2 * nw + 2] = FT_BIAS;
    }
    // load the cell image: 32-bit words widened to u16 pairs (+ bias); pixel column tc lives in word 1 + tc/2
    const uint8_t* S = pyr + (size_t)img * pyrBytes + (unsigned)lvl.x + (size_t)y0 * pitch + xa;
    const uint32_t magicNw = (uint32_t)mg.x, magicNpr = (uint32_t)mg.y;
    for (int i = tid; i < th * nw; i += FT_THREADS) {
        const int r = nw == 1 ? i : (int)__umulhi((uint32_t)i, magicNw);
        const int k = i - r * nw;
        const uint32_t v = *reinterpret_cast<const uint32_t*>(S + (unsigned)(r * pitch + 4 * k));
        uint32_t* t = tile + r * P + 2 * k;
        t[1] = __byte_perm(v, FT_BIAS, 0x5150);
        t[2] = __byte_perm(v, FT_BIAS, 0x5352);
    }

    // cv::FAST(cell, iniThFAST) and, only when that leaves the cell empty, cv::FAST(cell, minThFAST)
    // (reference :809-816).  Per pass: quick test on every pair -> full score network on the pairs that pass
    // (every pixel that reaches the threshold is among them, the others score 0 exactly as a non-corner does
    // in cv::FAST's score buffer) -> 3x3 strict-greater NMS.
    // Work lists are kept as one segment per warp, filled with ballots (no atomics).
    const int total = hi * npr;
    const int qseg2s = ((((hi + 1) >> 1) * npr) + FT_WARPS - 1) / FT_WARPS;   // quick-test items (two rows each) per warp
    const int wseg = ((total + FT_THREADS - 1) / FT_THREADS) * 32;      // upper bound of a warp's share of the score loop
    const unsigned ltmask = (1u << lane) - 1u;
    int nEmit = 0;
    for (int pass = 0; pass < passes; pass++) {
        const int thr = pass ? minTh : iniTh;
        const uint32_t Tp = (uint32_t)(min(max(thr, 0), 255) + 1) * 0x00010001u;
        if (tid == 0) sN = 0;
        __syncthreads();                                                // tile loaded / previous pass done

        // ---- quick test; also clears the pairs' score words.  An item is a pixel pair in two vertically adjacent rows:
        //      one index computation, the second row's addresses are immediates ----
        {
            const int total2 = ((hi + 1) >> 1) * npr;
            const int qseg2 = (total2 + FT_WARPS - 1) / FT_WARPS;
            const int cbeg = wid * qseg2, cend = min(cbeg + qseg2, total2);
            uint16_t* myq = qlist + 2 * cbeg;
            int cnt = 0;
            for (int b = cbeg; b < cend; b += 32) {
                const int i = min(b + lane, cend - 1);                  // surplus lanes repeat the last item
                const int rp = npr == 1 ? i : (int)__umulhi((uint32_t)i, magicNpr);
                const int m = m0 + (i - rp * npr);
                const int rr = 2 * rp;
                const uint32_t* t = tile + (rr + 3) * P + m + 1;
                const uint32_t hit0 = fast_may_pass<P>(t, Tp);
                const uint32_t hit1 = fast_may_pass<P>(t + P, Tp);      // row hi of an odd cell: reads stay inside shared memory, result unused
                uint32_t* z = scr + (rr + 1) * P + m + 1;
                z[0] = 0; z[P] = 0;                                     // (row hi + 1 is the zero frame row)
                const bool live = b + lane < cend;
                const unsigned bal0 = __ballot_sync(0xffffffffu, hit0 != 0 && live);
                const unsigned bal1 = __ballot_sync(0xffffffffu, hit1 != 0 && live && rr + 1 < hi);
                const int idx = rr * npr + (m - m0);
                if ((bal0 >> lane) & 1u) myq[cnt + __popc(bal0 & ltmask)] = (uint16_t)idx;
                cnt += __popc(bal0);
                if ((bal1 >> lane) & 1u) myq[cnt + __popc(bal1 & ltmask)] = (uint16_t)(idx + npr);
                cnt += __popc(bal1);
            }
            if (lane == 0) sQn[wid] = cnt;
        }
        __syncthreads();

        // ---- scores of the listed pairs ----
        {
            int nq[FT_WARPS], nQ = 0;
#pragma unroll
            for (int k = 0; k < FT_WARPS; k++) { nq[k] = sQn[k]; nQ += nq[k]; }
            uint16_t* myw = wlist + wid * wseg;
            int cnt = 0;
            for (int b = wid * 32; b < nQ; b += FT_THREADS) {
                const int j = min(b + lane, nQ - 1);
                const int i = seg_list_at(qlist, 2 * qseg2s, nq, j);
                const int rr = npr == 1 ? i : (int)__umulhi((uint32_t)i, magicNpr);
                const int m = m0 + (i - rr * npr);
                const uint32_t* t = tile + (rr + 3) * P + m + 1;
                uint32_t r[16];
                fast_load_ring<P>(t, r);
                uint32_t s = fast_score_pair(t[0], r);
                const int c = 2 * m;
                if (c < cx0 || c >= cx1) s &= 0xffff0000u;              // pixels outside the inner rectangle score 0
                if (c + 1 < cx0 || c + 1 >= cx1) s &= 0x0000ffffu;
                scr[(rr + 1) * P + m + 1] = s;
                // a pair goes on the NMS work list when one of its pixels reaches the threshold
                const bool keep = ((int)(s & 0xffffu) >= thr || (int)(s >> 16) >= thr) && b + lane < nQ;
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                if ((bal >> lane) & 1u) myw[cnt + __popc(bal & ltmask)] = (uint16_t)i;
                cnt += __popc(bal);
            }
            if (lane == 0) sWn[wid] = cnt;
        }
        __syncthreads();

        // ---- NMS + threshold on the listed pairs (raw neighbour scores suffice: a neighbour below the
        //      threshold is below S anyway) ----
        {
            int nwk[FT_WARPS], nWork = 0;
#pragma unroll
            for (int k = 0; k < FT_WARPS; k++) { nwk[k] = sWn[k]; nWork += nwk[k]; }
            for (int j = tid; j < nWork; j += FT_THREADS) {
                const int i = seg_list_at(wlist, wseg, nwk, j);
                const int rr = npr == 1 ? i : (int)__umulhi((uint32_t)i, magicNpr);
                const int m = m0 + (i - rr * npr);
                const uint32_t* q = scr + (rr + 1) * P + m + 1;
                const uint32_t w = q[0];
                const int sl = w & 0xffff, sh = w >> 16;
                const uint32_t u0 = q[-P - 1], u1 = q[-P], u2 = q[-P + 1];
                const uint32_t c0 = q[-1], c2 = q[1];
                const uint32_t d0 = q[P - 1], d1 = q[P], d2 = q[P + 1];
                uint32_t nb = max3s(u1, __funnelshift_r(u0, u1, 16), __funnelshift_r(u1, u2, 16));
                nb = max3s(nb, __funnelshift_r(c0, w, 16), __funnelshift_r(w, c2, 16));
                nb = max3s(nb, d1, __funnelshift_r(d0, d1, 16));
                nb = __vmaxs2(nb, __funnelshift_r(d1, d2, 16));
                const int nl = nb & 0xffff, nh = nb >> 16;
                const int ly = y0 + rr + 3;
                const int ry = ly - FAST_BORDER;                        // region coordinates
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int sv = e ? sh : sl, nv = e ? nh : nl;
                    const int lx = xa + 2 * m + e;
                    if (sv >= thr && sv > nv && lx >= ex0 && lx < ex1 && ly >= ey0 && ly < ey1) {
                        const int slot = atomicAdd(&sN, 1);
                        const int rx = lx - FAST_BORDER;
                        if (slot < clistCap) clist[slot] = (uint32_t)rx | ((uint32_t)ry << 12) | ((uint32_t)sv << 24);
                    }
                }
            }
        }
        __syncthreads();
        nEmit = min(sN, clistCap);
        if (nEmit > 0) break;
    }
    if (nEmit == 0) return;
    if (tid == 0) sBase = atomicAdd(&candCount[img * MAX_LEVELS + level], nEmit);
    __syncthreads();
    uint32_t* out = cand + (size_t)img * candPerImg + (unsigned)lvl.z;
    const int base = sBase, candCap = lvl.w;
    for (int i = tid; i < nEmit; i += FT_THREADS)
        if (base + i < candCap) out[base + i] = clist[i];
}

// ---------------------------------------------------------------------------------------------------
// Grid FAST, strip form: one CTA per GROUP of up to four horizontally adjacent cells of one cell row (the reference's cells,
// src/ORBextractor.cc:789-829: same rectangles, same per-cell isolation, same per-cell threshold fallback).
// fast_cells_kernel above (one CTA per 30-px cell; still used for cv::ORB's whole-level tiles of the birdview path) spends
// 69 % of its executed instructions on integer glue (profiles/r2_summary.md: VIMNMX[3] 21 %, LDS 10 %): per-item index
// divisions, list lookups across warp segments, per-cell setup for 900 pixels.  This form removes the glue instead of the math:
//   * a tile is ~130 x 37 pixels; warp w owns rows w, w+4, ...: a row step is 32 adjacent pixel pairs, every ring offset is an
//     immediate, consecutive lanes read consecutive words (no bank conflicts), no index arithmetic per item;
//   * each warp compacts ITS rows' survivors into its own list segment (ballot + popc), runs the score network on that segment
//     and compacts the pairs that reach the threshold in place: no barrier and no cross-warp lookup between the three steps;
//   * cell isolation is a pair of lane masks per pixel pair (left / right neighbour belongs to the same cell), applied to the
//     packed 3x3 maximum; the per-cell fallback (cv::FAST(minThFAST) only for cells that cv::FAST(iniThFAST) left empty) is a
//     third mask, so the second pass runs on the lanes of the empty cells only;
//   * the score tile is cleared once: a pair that fails the quick test at the lower threshold also failed it at the higher one
//     and was never written.
// Scores, maxima and candidates are identical to fast_cells_kernel's (same network, same NMS rule); tests/test_gpu_parity.py
// compares candidate sets per level with the oracle.
// ---------------------------------------------------------------------------------------------------
#ifndef ORBB200_FS_MINBLK
#define ORBB200_FS_MINBLK 6
#endif
#ifndef ORBB200_FS_LOADS
#define ORBB200_FS_LOADS 10
#endif
#ifndef ORBB200_FS_ROWS
#define ORBB200_FS_ROWS 2
#endif
constexpr int FS_LOADS = ORBB200_FS_LOADS;   // global words in flight per thread while the tile is filled
constexpr int FS_ROWS = ORBB200_FS_ROWS;     // rows per quick-test step

template <int P>
__global__ void __launch_bounds__(FS_THREADS, ORBB200_FS_MINBLK) fast_strip_kernel(const uint8_t* __restrict__ pyr, unsigned pyrBytes, unsigned candPerImg,
                                                                                 int minTh, int iniTh, const int4* __restrict__ groups,
                                                                                 uint32_t* __restrict__ cand, int32_t* __restrict__ candCount,
                                                                                 int tileRows, int scrRows, int segCap, int clistCap)
{
    extern __shared__ uint32_t fsSmem[];
    uint32_t* tile = fsSmem;                             // [tileRows][P]: word 1 + m of a row = pixels (2m, 2m+1) as u16x2 (+ bias)
    uint32_t* scr = tile + ((tileRows * P + 3) & ~3);    // [scrRows][P]: scores, same layout, one zero row above and below (16-byte aligned)
    uint32_t* inMask = scr + ((scrRows * P + 3) & ~3);   // per pair word: lanes inside the group's inner columns
    uint32_t* lMask = inMask + P;                        //                lanes whose left neighbour lies in the same cell
    uint32_t* rMask = lMask + P;                         //                lanes whose right neighbour lies in the same cell
    uint32_t* p2Mask = rMask + P;                        //                lanes of cells the first pass left empty
    uint32_t* cellIdx = p2Mask + P;                      //                cell of lane 0 | cell of lane 1 << 8
    uint32_t* clist = cellIdx + P;                       // [clistCap] candidates of this group
    uint16_t* lists = reinterpret_cast<uint16_t*>(clist + clistCap);   // [FS_WARPS][segCap] pair codes (row << 7 | pair word)
    __shared__ int sN, sBase, sCellHit[FS_MAX_CELLS];

    const int img = blockIdx.y;
    pdl_launch_dependents();
    const int4 ga = __ldg(groups + 3 * blockIdx.x);      // {x0|y0<<16, x1|y1<<16, level, cells}
    const int4 gb = __ldg(groups + 3 * blockIdx.x + 1);  // {level byte offset, pitch, candOff, candCap}
    const int4 gc = __ldg(groups + 3 * blockIdx.x + 2);  // {wCell, -, -, -}
    const int x0 = ga.x & 0xffff, y0 = ga.x >> 16, x1 = ga.y & 0xffff, y1 = ga.y >> 16;
    const int level = ga.z, nCells = ga.w, wCell = gc.x;
    const int pitch = gb.y;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int xa = x0 & ~3;
    const int th = y1 - y0, hi = th - 6;
    const int firstIn = x0 + 3, endIn = x1 - 3;          // inner columns of the group [firstIn, endIn), level coordinates
    if (hi <= 0 || endIn <= firstIn) return;
    const int m0 = (firstIn - xa) >> 1, m1 = (endIn - 1 - xa) >> 1, npr = m1 - m0 + 1;
    const int nw = (x1 - xa + 3) >> 2;                   // 32-bit words loaded per row

    if (tid < FS_MAX_CELLS) sCellHit[tid] = 0;
    if (tid == 0) sN = 0;
    // The group's image: 32-bit words widened to u16 pairs (+ bias), the (row, word) items flattened over the CTA and requested
    // FS_LOADS at a time before any is stored -- for the usual tile that is ALL of a thread's words in one global round trip, and
    // the first batch is requested before the masks and the score clear below, which need no pixel.
    const uint8_t* S = pyr + (size_t)img * pyrBytes + (unsigned)gb.x + (size_t)y0 * pitch + xa;
    const uint32_t magicNw = (uint32_t)gc.y;             // 2^32 / nw + 1
    const int total = th * nw;
    uint32_t v[FS_LOADS];
    int o[FS_LOADS];
    auto request = [&](int base) {
#pragma unroll
        for (int j = 0; j < FS_LOADS; j++) {
            const int i = min(base + j * FS_THREADS, total - 1);
            const int r = nw == 1 ? i : (int)__umulhi((uint32_t)i, magicNw);
            const int k = i - r * nw;
            v[j] = *reinterpret_cast<const uint32_t*>(S + (unsigned)(r * pitch + 4 * k));
            o[j] = r * P + 2 * k;
        }
    };
    auto widen = [&]() {
#pragma unroll
        for (int j = 0; j < FS_LOADS; j++) {             // (surplus items rewrite the last word with the same value)
            tile[o[j] + 1] = __byte_perm(v[j], FT_BIAS, 0x5150);
            tile[o[j] + 2] = __byte_perm(v[j], FT_BIAS, 0x5352);
        }
    };
    pdl_wait();                                          // the level's pixels are final; the candidate counters are zeroed
    request(tid);
    // per-pair lane masks
    for (int i = tid; i < P; i += FS_THREADS) {
        uint32_t in = 0, lm = 0, rm = 0, ci = 0;
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int x = xa + 2 * (i - 1) + e;
            if (x >= firstIn && x < endIn) {
                const int k = min((x - firstIn) / wCell, nCells - 1);
                const int cs = firstIn + k * wCell, ce = min(cs + wCell, endIn);
                const uint32_t L = 0xffffu << (16 * e);
                in |= L;
                if (x > cs) lm |= L;
                if (x < ce - 1) rm |= L;
                ci |= (uint32_t)k << (8 * e);
            }
        }
        inMask[i] = in; lMask[i] = lm; rMask[i] = rm; cellIdx[i] = ci; p2Mask[i] = in;
    }
    for (int i = tid; i < (scrRows * P + 3) >> 2; i += FS_THREADS) reinterpret_cast<uint4*>(scr)[i] = make_uint4(0, 0, 0, 0);
    // words next to the loaded span hold the bias so that masked lanes stay in range (see fast_cells_kernel)
    for (int r = tid; r < th; r += FS_THREADS) {
        uint32_t* t = tile + r * P;
        t[0] = FT_BIAS; t[2 * nw + 1] = FT_BIAS; t[2 * nw + 2] = FT_BIAS;
    }
    widen();
    for (int base = tid + FS_THREADS * FS_LOADS; base < total; base += FS_THREADS * FS_LOADS) { request(base); widen(); }
    __syncthreads();

    const unsigned ltmask = (1u << lane) - 1u;
    uint16_t* myq = lists + wid * segCap;
    for (int pass = 0; pass < 2; pass++) {
        const int thr = pass ? minTh : iniTh;
        const uint32_t Tp = (uint32_t)(min(max(thr, 0), 255) + 1) * 0x00010001u;
        const uint32_t* laneMask = pass ? p2Mask : inMask;

        // ---- quick test: this warp's blocks of FS_ROWS rows, 32 pairs x FS_ROWS rows per step (one address computation; the rows
        //      below are immediates) ----
        int nq = 0;
        for (int rr0 = wid * FS_ROWS; rr0 < hi; rr0 += FS_WARPS * FS_ROWS) {
            const uint32_t* tb = tile + (rr0 + 3) * P + 1 + m0;
            for (int c0 = 0; c0 < npr; c0 += 32) {
                const int c = min(c0 + lane, npr - 1);              // surplus lanes repeat the last pair
                const bool valid = c0 + lane < npr;
                const uint32_t lm = pass ? laneMask[1 + m0 + c] : 0xffffffffu;
                const int code0 = (rr0 << 7) | (m0 + c);
#pragma unroll
                for (int j = 0; j < FS_ROWS; j++) {                 // rows past hi read the words behind the tile (the score tile); not listed
                    const uint32_t hit = fast_may_pass<P>(tb + c + j * P, Tp) & lm;
                    const bool on = hit != 0 && valid && rr0 + j < hi;
                    const unsigned bal = __ballot_sync(0xffffffffu, on);
                    if (on) myq[nq + __popc(bal & ltmask)] = (uint16_t)(code0 + (j << 7));
                    nq += __popc(bal);
                }
            }
        }
        __syncwarp();

        // ---- scores of this warp's survivors; the pairs that reach the threshold are compacted in place ----
        int nk = 0;
        for (int b = 0; b < nq; b += 32) {
            const int code = myq[min(b + lane, nq - 1)];
            const int rr = code >> 7, m = code & 127;
            const uint32_t* t = tile + (rr + 3) * P + 1 + m;
            uint32_t r[16];
            fast_load_ring<P>(t, r);
            uint32_t sc = fast_score_pair(t[0], r) & laneMask[1 + m];   // pixels outside the inner columns (second pass: outside the empty cells) score 0
            scr[(rr + 1) * P + 1 + m] = sc;
            const bool keep = ((int)(sc & 0xffffu) >= thr || (int)(sc >> 16) >= thr) && b + lane < nq;
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            __syncwarp();                                           // every lane has read its code before slots are rewritten
            if ((bal >> lane) & 1u) myq[nk + __popc(bal & ltmask)] = (uint16_t)code;
            nk += __popc(bal);
        }
        __syncthreads();                                            // all scores of the pass are in the tile

        // ---- 3x3 strict-greater NMS inside the pixel's own cell + threshold (raw neighbour scores suffice) ----
        for (int b = 0; b < nk; b += 32) {
            if (b + lane < nk) {
                const int code = myq[b + lane];
                const int rr = code >> 7, m = code & 127;
                const uint32_t* q = scr + (rr + 1) * P + 1 + m;
                const uint32_t w = q[0];
                const uint32_t u0 = q[-P - 1], u1 = q[-P], u2 = q[-P + 1];
                const uint32_t c0 = q[-1], c2 = q[1];
                const uint32_t d0 = q[P - 1], d1 = q[P], d2 = q[P + 1];
                const uint32_t left = max3s(__funnelshift_r(u0, u1, 16), __funnelshift_r(c0, w, 16), __funnelshift_r(d0, d1, 16)) & lMask[1 + m];
                const uint32_t right = max3s(__funnelshift_r(u1, u2, 16), __funnelshift_r(w, c2, 16), __funnelshift_r(d1, d2, 16)) & rMask[1 + m];
                const uint32_t nb = max3s(left, right, __vmaxs2(u1, d1));
                const uint32_t ci = cellIdx[1 + m];
                const int ly = y0 + rr + 3;
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int sv = (w >> (16 * e)) & 0xffff, nv = (nb >> (16 * e)) & 0xffff;     // masked lanes hold sv == 0
                    if (sv >= thr && sv > nv) {
                        const int slot = atomicAdd(&sN, 1);
                        const int lx = xa + 2 * m + e;
                        if (slot < clistCap) clist[slot] = (uint32_t)(lx - FAST_BORDER) | ((uint32_t)(ly - FAST_BORDER) << 12) | ((uint32_t)sv << 24);
                        sCellHit[(ci >> (8 * e)) & 0xff] = 1;
                    }
                }
            }
        }
        __syncthreads();
        if (pass) break;
        // cells without a corner take the second pass (reference :809-816)
        bool again = false;
        for (int k = 0; k < nCells; k++) again |= sCellHit[k] == 0;
        if (!again) break;
        for (int i = tid; i < P; i += FS_THREADS) {
            const uint32_t ci = cellIdx[i], in = inMask[i];
            uint32_t pm = 0;
            if ((in & 0xffffu) && sCellHit[ci & 0xff] == 0) pm |= 0xffffu;
            if ((in >> 16) && sCellHit[(ci >> 8) & 0xff] == 0) pm |= 0xffff0000u;
            p2Mask[i] = pm;
        }
        __syncthreads();
    }
    const int nEmit = min(sN, clistCap);
    if (nEmit == 0) return;
    if (tid == 0) sBase = atomicAdd(&candCount[img * MAX_LEVELS + level], nEmit);
    __syncthreads();
    uint32_t* out = cand + (size_t)img * candPerImg + (unsigned)gb.z;
    const int base = sBase, candCap = gb.w;
    for (int i = tid; i < nEmit; i += FS_THREADS)
        if (base + i < candCap) out[base + i] = clist[i];
}

// ---------------------------------------------------------------------------------------------------
// DistributeOctTree (reference src/ORBextractor.cc:539-763, DivideNode :481-537) as bulk passes over
// arrays, one CTA per (image, level).  The std::list is a position-ordered node table; every pass:
//   1. count the four children of each node that may be split (one sweep over the candidates),
//   2. choose the split set: all nodes with >1 keys ("full" pass, :598-665) or the size-sorted prefix of
//      last pass's children up to the first point where |list| >= N (:676-733),
//   3. new list = reverse(children in processing order) ++ (unsplit nodes in old order),
//   4. relabel the candidates (second sweep).
// Size ties in the sorted pass: most recently created node first == smaller list position first
// (the documented tie rule).  Winner per final node: max response, first in
// (cell row, cell col, y, x) order on ties (:744-759).
// ---------------------------------------------------------------------------------------------------
constexpr int OT_THREADS = 256;          // batches: many (image, level) CTAs in flight
#ifndef ORBB200_OT_FEW
#define ORBB200_OT_FEW 1024
#endif
constexpr int OT_THREADS_FEW = ORBB200_OT_FEW;     // one or two images: the level-0 CTA is the critical path, its sweeps are latency chains

struct OtNode { short x0, y0, x1, y1; };

template <int NT>
__device__ __forceinline__ int block_scan_excl(int v, int* warpSums, int& total)
{
    // exclusive scan across the block of one value per thread; all threads must call
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warpSums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        int s = lane < (NT / 32) ? warpSums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        if (lane < (NT / 32)) warpSums[lane] = s;
    }
    __syncthreads();
    const int wbase = wid ? warpSums[wid - 1] : 0;
    total = warpSums[NT / 32 - 1];
    __syncthreads();
    return wbase + x - v;
}

size_t octree_smem_bytes(int maxNodes, int smemCand)
{
    const int P2 = 1 << (32 - __builtin_clz(std::max(maxNodes, 2) - 1));
    // nodes A,B (8 B) + cnt A,B (4 B) + cc (16 B) + newPos (4 B) + childPos (16 B) + order (4 B) + sort keys (P2*4)
    // + the shared-memory copy of the candidates and their labels (6 B each)
    return (size_t)maxNodes * (8 * 2 + 4 * 2 + 16 + 4 + 16 + 4 + 4) + (size_t)P2 * 4 + 256 + (size_t)smemCand * 6;
}

// One sweep over a level's candidates: the packed candidate and its node label of OT_UNROLL strided items are
// requested before any of them is used.  A thread that loads one item, uses it and only then loads the next keeps a
// single L2 request in flight, and the (image, level) CTAs with ~10^4 candidates then set the kernel's duration by
// their ~20 dependent sweeps (measured: latency-bound, not issue-bound).
constexpr int OT_UNROLL = 4;
template <int NT, bool NEED_NODE, typename F>
__device__ __forceinline__ void ot_sweep(const uint32_t* __restrict__ C, const uint16_t* nodeOf, int n, int tid, F f)
{
    for (int base = tid; base < n; base += OT_UNROLL * NT) {
        uint32_t v[OT_UNROLL];
        int p[OT_UNROLL];
#pragma unroll
        for (int k = 0; k < OT_UNROLL; k++) {
            const int i = base + k * NT;
            v[k] = i < n ? C[i] : 0u;
            p[k] = (NEED_NODE && i < n) ? (int)nodeOf[i] : 0;
        }
#pragma unroll
        for (int k = 0; k < OT_UNROLL; k++) {
            const int i = base + k * NT;
            if (i < n) f(i, v[k], p[k]);
        }
    }
}

template <int NT>
__global__ void __launch_bounds__(NT) octree_kernel(Geom g, const uint32_t* __restrict__ cand, const int32_t* __restrict__ candCount,
                                                            uint16_t* __restrict__ nodeOfAll, uint32_t* __restrict__ lvlKp,
                                                            int32_t* __restrict__ lvlCount, int32_t* __restrict__ status, int levelBase, int smemCand)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int warpSums[NT / 32];
    __shared__ int sFlag;

    // blockIdx.y = level: the CTAs of level 0 (most candidates, longest) are dispatched first, the short ones fill the tail
    const int level = levelBase + blockIdx.y, img = blockIdx.x;
    const LevelGeom L = g.lv[level];
    pdl_launch_dependents();
    pdl_wait();
    const int tid = threadIdx.x;
    const int N = L.quota;
    const int maxNodes = L.maxNodes;
    int n = min(candCount[img * MAX_LEVELS + level], L.candCap);
    const uint32_t* C = cand + (size_t)img * g.candPerImg + L.candOff;
    uint16_t* nodeOf = nodeOfAll + (size_t)img * g.candPerImg + L.candOff;
    // One or two images: a CTA has the SM to itself, and its ~20 sweeps over the candidates are the critical path of the whole
    // extraction.  When the level's candidates fit the shared memory left over (smemCand entries), the packed candidates are copied
    // there by the first sweep and the node labels live there from the start: every later sweep is shared-memory only.
    const bool local = n <= smemCand;
    uint32_t* outKp = lvlKp + (size_t)img * g.kpPerImg + L.kpOff;

    // carve shared memory
    const int P2 = 1 << (32 - __clz(max(maxNodes, 2) - 1));
    OtNode* nodesA = reinterpret_cast<OtNode*>(smem);
    OtNode* nodesB = nodesA + maxNodes;
    int* cntA = reinterpret_cast<int*>(nodesB + maxNodes);
    int* cntB = cntA + maxNodes;
    int* cc = cntB + maxNodes;               // [maxNodes][4] child counts
    int* newPos = cc + 4 * maxNodes;         // [maxNodes] new position of an unsplit node, or -1 if split
    int* childPos = newPos + maxNodes;       // [maxNodes][4] new position of each child (-1 if empty)
    int* order = childPos + 4 * maxNodes;    // [maxNodes] processing order -> node position
    int* aux = order + maxNodes;             // [maxNodes] scratch (split flag / processing rank)
    uint32_t* keys = reinterpret_cast<uint32_t*>(aux + maxNodes);   // [P2] sort keys
    uint32_t* sC = keys + P2;                                       // [smemCand] candidates (few-image launches)
    if (local) nodeOf = reinterpret_cast<uint16_t*>(sC + smemCand); // [smemCand] node labels

    if (n == 0 || L.nIni <= 0 || L.nIni > maxNodes) {
        if (tid == 0) lvlCount[img * MAX_LEVELS + level] = 0;
        return;
    }

    OtNode* cur = nodesA;
    OtNode* nxt = nodesB;
    int* curCnt = cntA;
    int* nxtCnt = cntB;

    // ---- root nodes (:543-592) ----
    const int nIni = L.nIni;
    const float hX = L.hX;
    const int regH = L.maxBY - FAST_BORDER;
    for (int i = tid; i < nIni; i += NT) {
        OtNode nd;
        nd.x0 = (short)(int)__fmul_rn(hX, (float)i);
        nd.x1 = (short)(int)__fmul_rn(hX, (float)(i + 1));
        nd.y0 = 0;
        nd.y1 = (short)regH;
        cur[i] = nd;
        curCnt[i] = 0;
    }
    __syncthreads();
    ot_sweep<NT, false>(C, nodeOf, n, tid, [&](int i, uint32_t v, int) {
        const int x = v & 0xfff;
        int r = (int)__fdiv_rn((float)x, hX);
        r = min(r, nIni - 1);
        nodeOf[i] = (uint16_t)r;
        if (local) sC[i] = v;
        atomicAdd(&curCnt[r], 1);
    });
    if (local) C = sC;
    __syncthreads();
    // drop empty roots (compaction in order)
    int listSize;
    {
        int total = 0;
        // nIni is tiny (<= a handful): serial by one thread
        if (tid == 0) {
            int k = 0;
            for (int i = 0; i < nIni; i++) {
                newPos[i] = curCnt[i] > 0 ? k : -1;
                if (curCnt[i] > 0) { nxt[k] = cur[i]; nxtCnt[k] = curCnt[i]; k++; }
            }
            sFlag = k;
        }
        __syncthreads();
        total = sFlag;
        ot_sweep<NT, true>(C, nodeOf, n, tid, [&](int i, uint32_t, int p) { nodeOf[i] = (uint16_t)newPos[p]; });
        __syncthreads();
        listSize = total;
        OtNode* t = cur; cur = nxt; nxt = t;
        int* tc = curCnt; curCnt = nxtCnt; nxtCnt = tc;
    }

    int frontNew = 0;        // number of nodes at the list front created by the last pass
    bool sortedMode = false;
    bool finish = false;
    int guard = 0;
    while (!finish && guard++ < 64) {
        const int prevSize = listSize;
        // ---- 1. which nodes may be split in this pass, and in which order ----
        // full pass: every node with cnt>1, list order.  sorted pass: front nodes with cnt>1, by (cnt desc, pos asc).
        int nProc = 0;
        if (!sortedMode) {
            int carry = 0;
            for (int base = 0; base < listSize; base += NT) {
                const int p = base + tid;
                const int f = (p < listSize && curCnt[p] > 1) ? 1 : 0;
                int tot;
                const int ex = block_scan_excl<NT>(f, warpSums, tot);
                if (f) order[carry + ex] = p;
                carry += tot;
            }
            nProc = carry;
        } else {
            int carry = 0;
            for (int base = 0; base < frontNew; base += NT) {
                const int p = base + tid;
                const int f = (p < frontNew && curCnt[p] > 1) ? 1 : 0;
                int tot;
                const int ex = block_scan_excl<NT>(f, warpSums, tot);
                if (f) keys[carry + ex] = ((uint32_t)min(curCnt[p], 0xffff) << 16) | (uint32_t)(0xffff - p);
                carry += tot;
            }
            nProc = carry;
            if (NT == OT_THREADS_FEW) {
                // one or two images: the CTA is alone on its SM and the ~50 barrier-separated steps of a sorting network are the
                // cost.  The keys are unique (they carry the position), so a key's place in the descending order is the number of
                // larger keys: every thread counts that for its key over the whole list (broadcast reads), no barrier in between.
                for (int i = nProc + tid; i < ((nProc + 3) & ~3); i += NT) keys[i] = 0;
                __syncthreads();
                for (int i = tid; i < nProc; i += NT) {
                    const uint32_t mine = keys[i];
                    int rank = 0;
                    for (int j = 0; j < nProc; j += 4)
                        rank += (keys[j] > mine) + (keys[j + 1] > mine) + (keys[j + 2] > mine) + (keys[j + 3] > mine);
                    order[rank] = 0xffff - (int)(mine & 0xffff);
                }
            } else {
                int S2 = 1;
                while (S2 < nProc) S2 <<= 1;
                for (int i = nProc + tid; i < S2; i += NT) keys[i] = 0;
                __syncthreads();
                // bitonic sort, descending
                for (int k = 2; k <= S2; k <<= 1) {
                    for (int j = k >> 1; j > 0; j >>= 1) {
                        for (int i = tid; i < S2; i += NT) {
                            const int ixj = i ^ j;
                            if (ixj > i) {
                                const uint32_t a = keys[i], b = keys[ixj];
                                const bool desc = (i & k) == 0;
                                if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[ixj] = a; }
                            }
                        }
                        __syncthreads();
                    }
                }
                for (int i = tid; i < nProc; i += NT) order[i] = 0xffff - (int)(keys[i] & 0xffff);
            }
        }
        __syncthreads();

        // ---- 2. child counts of the nodes in `order` ----
        for (int i = tid; i < listSize; i += NT) aux[i] = -1;
        __syncthreads();
        for (int i = tid; i < nProc; i += NT) {
            const int p = order[i];
            aux[p] = i;                       // processing rank
            cc[4 * p + 0] = 0; cc[4 * p + 1] = 0; cc[4 * p + 2] = 0; cc[4 * p + 3] = 0;
        }
        __syncthreads();
        ot_sweep<NT, true>(C, nodeOf, n, tid, [&](int, uint32_t v, int p) {
            if (aux[p] < 0) return;
            const int x = v & 0xfff, y = (v >> 12) & 0xfff;
            const OtNode nd = cur[p];
            const int mx = nd.x0 + ((nd.x1 - nd.x0 + 1) >> 1), my = nd.y0 + ((nd.y1 - nd.y0 + 1) >> 1);
            const int q = (x < mx) ? (y < my ? 0 : 2) : (y < my ? 1 : 3);
            atomicAdd(&cc[4 * p + q], 1);
        });
        __syncthreads();

        // ---- 3. how many of `order` are actually split (sorted pass stops once |list| >= N) ----
        // prefix over processing order of (non-empty children - 1)
        int nSplit = nProc;
        int nChildrenTotal = 0;
        {
            int carryKids = 0;
            if (tid == 0) sFlag = nProc;      // first processing index at which the list reaches N (sorted mode)
            __syncthreads();
            for (int base = 0; base < nProc; base += NT) {
                const int i = base + tid;
                int kids = 0;
                if (i < nProc) {
                    const int p = order[i];
                    kids = (cc[4 * p] > 0) + (cc[4 * p + 1] > 0) + (cc[4 * p + 2] > 0) + (cc[4 * p + 3] > 0);
                }
                int tot;
                const int ex = block_scan_excl<NT>(kids, warpSums, tot);
                if (i < nProc) {
                    // child sequence start for this node = carryKids + ex; stash in childPos[4p] for now
                    childPos[4 * order[i]] = carryKids + ex;
                    if (sortedMode) {
                        // list size after processing nodes 0..i (inclusive)
                        const int sizeAfter = prevSize + (carryKids + ex + kids) - (i + 1);
                        if (sizeAfter >= N) atomicMin(&sFlag, i);
                    }
                }
                carryKids += tot;
            }
            __syncthreads();
            if (sortedMode && sFlag < nProc) nSplit = sFlag + 1;
            __syncthreads();
            // total children of the split prefix
            if (nSplit > 0) {
                const int pl = order[nSplit - 1];
                const int kidsLast = (cc[4 * pl] > 0) + (cc[4 * pl + 1] > 0) + (cc[4 * pl + 2] > 0) + (cc[4 * pl + 3] > 0);
                nChildrenTotal = childPos[4 * pl] + kidsLast;
            }
        }
        __syncthreads();
        const int newSize = listSize - nSplit + nChildrenTotal;
        if (newSize > maxNodes) {       // cannot happen for sane geometry; fail loudly rather than corrupt
            if (tid == 0) { atomicExch(status, 1); lvlCount[img * MAX_LEVELS + level] = 0; }
            return;
        }

        // ---- 4. build the new list ----
        // children: sequence index s (processing order, n1..n4) -> position nChildrenTotal-1-s
        int nToExpand = 0;
        for (int i = tid; i < nSplit; i += NT) {
            const int p = order[i];
            const OtNode nd = cur[p];
            const int hx = (nd.x1 - nd.x0 + 1) >> 1, hy = (nd.y1 - nd.y0 + 1) >> 1;
            int s = childPos[4 * p];
            const int c0 = cc[4 * p], c1 = cc[4 * p + 1], c2 = cc[4 * p + 2], c3 = cc[4 * p + 3];
            int pos[4] = {-1, -1, -1, -1};
            if (c0 > 0) { pos[0] = nChildrenTotal - 1 - s; s++; }
            if (c1 > 0) { pos[1] = nChildrenTotal - 1 - s; s++; }
            if (c2 > 0) { pos[2] = nChildrenTotal - 1 - s; s++; }
            if (c3 > 0) { pos[3] = nChildrenTotal - 1 - s; s++; }
            if (pos[0] >= 0) { OtNode c; c.x0 = nd.x0; c.y0 = nd.y0; c.x1 = nd.x0 + hx; c.y1 = nd.y0 + hy; nxt[pos[0]] = c; nxtCnt[pos[0]] = c0; }
            if (pos[1] >= 0) { OtNode c; c.x0 = nd.x0 + hx; c.y0 = nd.y0; c.x1 = nd.x1; c.y1 = nd.y0 + hy; nxt[pos[1]] = c; nxtCnt[pos[1]] = c1; }
            if (pos[2] >= 0) { OtNode c; c.x0 = nd.x0; c.y0 = nd.y0 + hy; c.x1 = nd.x0 + hx; c.y1 = nd.y1; nxt[pos[2]] = c; nxtCnt[pos[2]] = c2; }
            if (pos[3] >= 0) { OtNode c; c.x0 = nd.x0 + hx; c.y0 = nd.y0 + hy; c.x1 = nd.x1; c.y1 = nd.y1; nxt[pos[3]] = c; nxtCnt[pos[3]] = c3; }
            childPos[4 * p] = pos[0]; childPos[4 * p + 1] = pos[1]; childPos[4 * p + 2] = pos[2]; childPos[4 * p + 3] = pos[3];
            nToExpand += (c0 > 1) + (c1 > 1) + (c2 > 1) + (c3 > 1);
        }
        // unsplit nodes keep their relative order after the children
        {
            int carry = 0;
            for (int base = 0; base < listSize; base += NT) {
                const int p = base + tid;
                const int keep = (p < listSize && !(aux[p] >= 0 && aux[p] < nSplit)) ? 1 : 0;
                int tot;
                const int ex = block_scan_excl<NT>(keep, warpSums, tot);
                if (p < listSize) {
                    if (keep) {
                        const int np = nChildrenTotal + carry + ex;
                        newPos[p] = np;
                        nxt[np] = cur[p];
                        nxtCnt[np] = curCnt[p];
                    } else {
                        newPos[p] = -1;
                    }
                }
                carry += tot;
            }
        }
        // nToExpand (block sum) only matters after a full pass
        int nToExpandTotal;
        {
            int tot;
            block_scan_excl<NT>(nToExpand, warpSums, tot);
            nToExpandTotal = tot;
        }
        __syncthreads();

        // ---- 5. relabel candidates ----
        ot_sweep<NT, true>(C, nodeOf, n, tid, [&](int i, uint32_t v, int p) {
            const int np = newPos[p];
            if (np >= 0) { nodeOf[i] = (uint16_t)np; return; }
            const int x = v & 0xfff, y = (v >> 12) & 0xfff;
            const OtNode nd = cur[p];
            const int mx = nd.x0 + ((nd.x1 - nd.x0 + 1) >> 1), my = nd.y0 + ((nd.y1 - nd.y0 + 1) >> 1);
            const int q = (x < mx) ? (y < my ? 0 : 2) : (y < my ? 1 : 3);
            nodeOf[i] = (uint16_t)childPos[4 * p + q];
        });
        __syncthreads();
        {
            OtNode* t = cur; cur = nxt; nxt = t;
            int* tc = curCnt; curCnt = nxtCnt; nxtCnt = tc;
        }
        listSize = newSize;
        frontNew = nChildrenTotal;

        // ---- 6. termination (:667-735) ----
        if (listSize >= N || listSize == prevSize) {
            finish = true;
        } else if (!sortedMode) {
            if (listSize + nToExpandTotal * 3 > N) sortedMode = true;
        }
    }

    // ---- winners (:744-759): max response, first in (cell row, cell col, y, x) order ----
    uint32_t* best = reinterpret_cast<uint32_t*>(cc);          // [listSize] max response
    uint32_t* bestKey = reinterpret_cast<uint32_t*>(childPos); // [listSize] min order key among max-response keys
    uint32_t* bestVal = reinterpret_cast<uint32_t*>(newPos);   // [listSize] packed candidate
    for (int i = tid; i < listSize; i += NT) { best[i] = 0; bestKey[i] = 0xffffffffu; }
    __syncthreads();
    ot_sweep<NT, true>(C, nodeOf, n, tid, [&](int, uint32_t v, int p) { atomicMax(&best[p], v >> 24); });
    __syncthreads();
    const int wCell = L.wCell, hCell = L.hCell, nCols = L.nCols;
    ot_sweep<NT, true>(C, nodeOf, n, tid, [&](int, uint32_t v, int p) {
        if ((v >> 24) != best[p]) return;
        const int x = v & 0xfff, y = (v >> 12) & 0xfff;
        const int cj = (x - 3) / wCell, ci = (y - 3) / hCell;
        const uint32_t key = ((uint32_t)(ci * nCols + cj) << 12) | ((uint32_t)(y - 3 - ci * hCell) << 6) | (uint32_t)(x - 3 - cj * wCell);
        atomicMin(&bestKey[p], key);
    });
    __syncthreads();
    ot_sweep<NT, true>(C, nodeOf, n, tid, [&](int, uint32_t v, int p) {
        if ((v >> 24) != best[p]) return;
        const int x = v & 0xfff, y = (v >> 12) & 0xfff;
        const int cj = (x - 3) / wCell, ci = (y - 3) / hCell;
        const uint32_t key = ((uint32_t)(ci * nCols + cj) << 12) | ((uint32_t)(y - 3 - ci * hCell) << 6) | (uint32_t)(x - 3 - cj * wCell);
        if (key == bestKey[p]) bestVal[p] = v;
    });
    __syncthreads();
    const int nOut = min(listSize, L.kpCap);
    for (int i = tid; i < nOut; i += NT) outKp[i] = bestVal[i];
    if (tid == 0) {
        lvlCount[img * MAX_LEVELS + level] = nOut;
        if (listSize > L.kpCap) atomicExch(status, 2);
    }
}

// ---------------------------------------------------------------------------------------------------
// Orientation (IC_Angle, reference src/ORBextractor.cc:77-104 + cv::fastAtan2) and rBRIEF
// (computeOrbDescriptor :108-147): one warp per keypoint.  Writes the final cv::KeyPoint records
// (operator() :1095-1103: pt *= scale for level > 0) and descriptors in level-major order.
// ---------------------------------------------------------------------------------------------------
#ifndef ORBB200_DS_NBUF
#define ORBB200_DS_NBUF 1
#endif
constexpr int DS_NBUF = ORBB200_DS_NBUF;
constexpr int DS_WARPS = 8;
constexpr int DS_KPB = 32;        // keypoints per CTA in batches: one lane each for the scalar (atan2, sincos) part, four per warp
constexpr int DS_KPB_FEW = 8;     // one or two images: one keypoint per warp, four times the CTAs (the four in a row were a latency chain)
constexpr int DS_R = 19;          // reach of the rotated rBRIEF pattern
constexpr int DS_PROWS = 2 * DS_R + 1;
constexpr int DS_TPATCH = ((DS_TBOX_W * DS_TBOX_H + 127) / 128) * 128;     // bytes of one TMA box in shared memory (128-byte aligned)

#ifndef ORBB200_DS_MINBLK
#define ORBB200_DS_MINBLK 8
#endif
template <int KPB>
__global__ void __launch_bounds__(DS_WARPS * 32, ORBB200_DS_MINBLK) describe_kernel(Geom g, const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ blur,
                                                                 const uint32_t* __restrict__ lvlKp, const int32_t* __restrict__ lvlCount,
                                                                 orbb200_kp_t* __restrict__ kps, uint8_t* __restrict__ desc,
                                                                 int32_t* __restrict__ counts, const CUtensorMap* __restrict__ maps, HostMirror M)
{
    constexpr int DS_KPB = KPB;
    __shared__ float sPX[16 * 32], sPY[16 * 32];      // [sample within the byte][lane]: conflict-free
    __shared__ int sLevel[DS_KPB], sX[DS_KPB], sY[DS_KPB], sResp[DS_KPB], sM01[DS_KPB], sM10[DS_KPB];
    __shared__ float sA[DS_KPB], sB[DS_KPB];
    __shared__ __align__(128) uint8_t sPatch[DS_WARPS][DS_NBUF][DS_TPATCH];   // TMA boxes per warp (with two, the next keypoint's lands while this one is used)
    __shared__ __align__(8) uint64_t sBar[DS_WARPS][2];
    const int img = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    pdl_launch_dependents();
    for (int i = tid; i < 512; i += blockDim.x) {
        const int o = (i & 15) * 32 + (i >> 4);
        sPX[o] = (float)c_patX[i]; sPY[o] = (float)c_patY[i];
    }
    pdl_wait();
    // ---- phase 0: which keypoint (output order = level-major, octree list order inside a level) ----
    if (tid < DS_KPB) {
        const int gk = blockIdx.x * DS_KPB + tid;
        int level = -1, off = 0, total = 0;
        for (int l = 0; l < g.nlevels; l++) {
            const int c = lvlCount[img * MAX_LEVELS + l];
            if (level < 0 && gk < total + c) { level = l; off = total; }
            total += c;
        }
        if (blockIdx.x == 0 && tid == 0) {
            counts[img] = total;
            if (M.counts) {                     // small host call: the results also go straight to the pinned mirror
                M.counts[img] = total;
                if (img == 0) *M.status = *M.d_status;
            }
        }
        sLevel[tid] = level;
        if (level >= 0) {
            const uint32_t v = lvlKp[(size_t)img * g.kpPerImg + g.lv[level].kpOff + (gk - off)];
            sX[tid] = (int)(v & 0xfff) + FAST_BORDER;
            sY[tid] = (int)((v >> 12) & 0xfff) + FAST_BORDER;
            sResp[tid] = (int)(v >> 24);
        }
    }
    __syncthreads();
    if (sLevel[0] < 0) return;      // keypoints are dense from index 0: nothing in this CTA

    // TMA plumbing of phase 3 (see there)
    const uint32_t barBase = (uint32_t)__cvta_generic_to_shared(&sBar[wid][0]);
    const uint32_t patchBase = (uint32_t)__cvta_generic_to_shared(&sPatch[wid][0][0]);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(barBase));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(barBase + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto fetch = [&](int j, int buf) {          // lane 0: arm the barrier with the box size and start the copy
        const int level = sLevel[j];
        const uint32_t bar = barBase + 8 * buf, dst = patchBase + DS_TPATCH * buf;
        const int cx = (sX[j] - DS_R) & ~15, cy = sY[j] - DS_R;             // box start: 16-byte aligned column
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(DS_TBOX_W * DS_TBOX_H) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     :: "r"(dst), "l"(reinterpret_cast<uint64_t>(maps + level)), "r"(cx), "r"(cy), "r"(img), "r"(bar) : "memory");
    };


    // ---- phase 1: IC_Angle moments on the un-blurred level, one warp per keypoint (lane = patch column; every load
    //      is one row segment).  (Tried: staging the patch in shared memory and reducing rows with byte dot products --
    //      a third of the instructions, no faster: the kernel is bound by the sectors its scattered patches pull
    //      from L2, 0.48 -> 0.50 ms.) ----
    for (int j = wid; j < DS_KPB; j += DS_WARPS) {
        const int level = sLevel[j];
        if (level < 0) break;
        const LevelGeom& L = g.lv[level];
        const uint8_t* center = pyr + (size_t)img * g.pyrBytes + L.off + (size_t)sY[j] * L.pitch + sX[j];
        int m01 = 0, m10 = 0;
        const int u = lane - HALF_PATCH;      // lanes 0..30 -> u = -15..15
        if (lane < 31) {
            m10 = u * center[u];
#pragma unroll
            for (int vv = 1; vv <= HALF_PATCH; vv++) {
                if (abs(u) <= c_umax[vv]) {
                    const int vp = center[u + vv * L.pitch], vm = center[u - vv * L.pitch];
                    m01 += vv * (vp - vm);
                    m10 += u * (vp + vm);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m01 += __shfl_xor_sync(0xffffffffu, m01, o);
            m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        }
        if (lane == 0) { sM01[j] = m01; sM10[j] = m10; }
    }
    __syncthreads();

    // ---- phase 2: angle, cos/sin and the keypoint record, one lane per keypoint ----
    if (tid < DS_KPB && sLevel[tid] >= 0) {
        const int level = sLevel[tid];
        const LevelGeom& L = g.lv[level];
        const float angle = fast_atan2_deg((float)sM01[tid], (float)sM10[tid]);
        constexpr float factorPI = (float)(3.14159265358979323846 / 180.0);   // (float)(CV_PI/180.f), :106
        const float ang = __fmul_rn(angle, factorPI);
        // cosf/sinf of the host libm are correctly rounded for all but ~1e-8 of inputs; round a double
        // evaluation to float to match them (float parity)
        double sd, cd;
        sincos((double)ang, &sd, &cd);
        sA[tid] = (float)cd; sB[tid] = (float)sd;
        orbb200_kp_t kp;
        kp.x = (float)sX[tid]; kp.y = (float)sY[tid];
        if (level != 0) { kp.x = __fmul_rn(kp.x, L.scale); kp.y = __fmul_rn(kp.y, L.scale); }
        kp.size = (float)L.patchSize;
        kp.angle = angle;
        kp.response = (float)sResp[tid];
        kp.octave = level;
        kp.class_id = -1;
        kps[(size_t)img * g.kpPerImg + blockIdx.x * DS_KPB + tid] = kp;
        if (M.kps) M.kps[(size_t)img * g.kpPerImg + blockIdx.x * DS_KPB + tid] = kp;
    }
    __syncthreads();

    // ---- phase 3: rBRIEF on the blurred level, one warp per keypoint, one descriptor byte per lane.
    //      The rotated pattern reaches at most 19 pixels from the centre (|p| <= sqrt(13^2+13^2) < 18.4, rounded).  The
    //      39 x 39 patch (as a 48 x 39-byte box) is fetched into shared memory by TMA (cp.async.bulk.tensor, one
    //      instruction per keypoint, completion on an mbarrier): the copy bypasses the LSU/L1 data path that bounds this
    //      kernel (the former 14 loads + 14 stores per lane were a third of its L1 wavefronts), and the next keypoint's
    //      box is in flight while the 512 byte gathers of the current one read theirs. ----
    // (requesting the warp's first box before phase 1 was measured slower: 0.418 against 0.406 ms per step)
    if (lane == 0 && sLevel[wid] >= 0) fetch(wid, 0);
    int it = 0;
    for (int j = wid; j < DS_KPB; j += DS_WARPS, it++) {
        const int level = sLevel[j];
        if (level < 0) break;
        const int buf = DS_NBUF == 2 ? (it & 1) : 0;
        const float a = sA[j], b = sB[j];
        // two buffers: the other one was read by the previous iteration (its __syncwarp() is behind us); order those
        // generic-proxy reads before the async-proxy write that refills it
        if (DS_NBUF == 2 && lane == 0 && j + DS_WARPS < DS_KPB && sLevel[j + DS_WARPS] >= 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            fetch(j + DS_WARPS, buf ^ 1);
        }
        {
            const uint32_t bar = barBase + 8 * buf, parity = (uint32_t)(DS_NBUF == 2 ? ((it >> 1) & 1) : (it & 1));
            asm volatile("{\n\t.reg .pred p;\n\tDSWAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DSDONE;\n\tbra DSWAIT;\n\tDSDONE:\n\t}"
                         :: "r"(bar), "r"(parity) : "memory");
        }
        const int mis = (sX[j] - DS_R) & 15;
        const uint8_t* pc = &sPatch[wid][buf][0] + DS_R * DS_TBOX_W + DS_R + mis;   // patch centre
        int val = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            int t[2];
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int idx = (2 * k + e) * 32 + lane;
                const float px = sPX[idx], py = sPY[idx];
                // cvRound == round-half-even: adding 1.5*2^23 rounds to an integer in the float adder
                const int yy = __float_as_int(__fadd_rn(__fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a)), 12582912.f)) - 0x4B400000;
                const int xx = __float_as_int(__fadd_rn(__fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b)), 12582912.f)) - 0x4B400000;
                t[e] = pc[yy * DS_TBOX_W + xx];
            }
            val |= (t[0] < t[1]) << k;
        }
        desc[((size_t)img * g.kpPerImg + blockIdx.x * DS_KPB + j) * 32 + lane] = (uint8_t)val;
        if (M.desc) M.desc[((size_t)img * g.kpPerImg + blockIdx.x * DS_KPB + j) * 32 + lane] = (uint8_t)val;
        __syncwarp();                                                        // every lane is done with this buffer
        if (DS_NBUF == 1 && lane == 0 && j + DS_WARPS < DS_KPB && sLevel[j + DS_WARPS] >= 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            fetch(j + DS_WARPS, 0);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------
void launch_import(Ctx& c, const uint8_t* d_imgs, size_t img_bytes, size_t stride, int n)
{
    const Geom& g = c.cur->g;
    dim3 grid(((g.w + 15) / 16 + 127) / 128, g.h, n);
    import_kernel<<<grid, 128, 0, c.stream>>>(d_imgs, img_bytes, stride, c.d_pyr, g.pyrBytes, g.lv[0].off, g.w, g.h, g.lv[0].pitch);
    c.launches++;
}

void launch_resizes(Ctx& c, int n, cudaStream_t stream, uint8_t* hostPyr)
{
    const Geom& g = c.cur->g;
    const ShapeTables& st = *c.cur;
    for (int l = 1; l < g.nlevels; l++) {
        const LevelGeom& d = g.lv[l];
        if (d.w <= 0 || d.h <= 0 || st.resizeTileCount[l] == 0) break;
        dim3 grid(st.resizeTileCount[l], n);
        const size_t smem = (size_t)st.resizeSmemPitch[l] * st.resizeSmemRows[l];
        // level 1 follows the import kernel or an upload (a full dependency either way); levels >= 2 follow a resize_kernel
        launch_chain(c.pdl && l > 1, st.resizeNarrow[l] ? resize_kernel<true> : resize_kernel<false>, grid, dim3(RS_THREADS), smem, stream, c.d_pyr,
                     g.pyrBytes, g.lv[l - 1], d, st.d_xtab, st.d_ytab, st.d_resizeTiles + 2 * st.resizeTileBase[l], st.resizeTileCount[l],
                     st.resizeSmemPitch[l], st.resizeSmemRows[l], st.resizeMapOk[l] ? st.d_rmaps + l : nullptr, hostPyr);
        c.launches++;
    }
}

// afterKernel (here and below): the previous operation on `stream` is a kernel of this chain, so the launch may be programmatic
void launch_border(Ctx& c, int n, cudaStream_t stream, bool afterKernel)
{
    const Geom& g = c.cur->g;
    launch_chain(c.pdl && afterKernel, border_kernel, dim3(g.nlevels * BD_CHUNKS, n), dim3(128), 0, stream, c.d_pyr, g.pyrBytes, g);
    c.launches++;
}

void launch_pyramid(Ctx& c, int n, uint8_t* hostPyr)
{
    launch_resizes(c, n, c.stream, hostPyr);
    launch_border(c, n, c.stream, c.cur->g.nlevels > 1);
}

void launch_blur(Ctx& c, int n, cudaStream_t stream, bool afterKernel)
{
    const Geom& g = c.cur->g;
    if (c.cur->nBlurTiles == 0) return;
    dim3 grid((c.cur->nBlurTiles + BL_WARPS - 1) / BL_WARPS, n);
    launch_chain(c.pdl && afterKernel, blur_kernel, grid, dim3(BL_WARPS * 32), 0, stream, c.d_pyr, c.d_blur, g.pyrBytes, g, c.cur->d_blurTiles, c.cur->nBlurTiles);
    c.launches++;
}

void push_fast_cell(std::vector<int4>& cells, FastSmem& need, int x0, int y0, int x1, int y1, int level, unsigned levelOff, int pitch,
                    unsigned candOff, int candCap, int ex0, int ey0, int ex1, int ey1)
{
    cells.push_back(make_int4(x0 | (y0 << 16), x1 | (y1 << 16), level, 0));
    cells.push_back(make_int4((int)levelOff, pitch, (int)candOff, candCap));
    cells.push_back(make_int4(0, 0, ex0 | (ey0 << 16), ex1 | (ey1 << 16)));
    // shared-memory needs of this cell, mirroring fast_cells_kernel's carve
    const int th = y1 - y0, xa = x0 & ~3;
    const int wi = x1 - x0 - 6, hi = th - 6;
    if (wi > 0 && hi > 0) {
        const int cx0 = x0 - xa + 3, cx1 = cx0 + wi, npr = ((cx1 - 1) >> 1) - (cx0 >> 1) + 1;
        const int nw = (x1 - xa + 3) >> 2;
        // floor(i / d) == umulhi(i, 2^32 / d + 1) for the small i used; d == 1 is special-cased in the kernel
        cells.back().x = (int)(0xffffffffu / (uint32_t)nw + 1u);
        cells.back().y = (int)(0xffffffffu / (uint32_t)npr + 1u);
        need.tileWords = std::max(need.tileWords, th);
        need.scrWords = std::max(need.scrWords, hi + 2);
        need.maxRowWords = std::max(need.maxRowWords, 2 * nw + 3);
        need.clistCap = std::max(need.clistCap, ((wi + 1) / 2) * ((hi + 1) / 2));
        need.workCap = std::max(need.workCap, ((hi * npr + 1) & ~1) + 128);   // per-warp list segments: total + slack
    }
}

void launch_fast_cells(Ctx& c, const uint8_t* d_pyr, unsigned pyrBytes, unsigned candPerImg, int minTh, int iniTh, int passes,
                       const int4* d_cells, int nCells, const FastSmem& need, uint32_t* d_cand, int32_t* d_candCount, int n)
{
    if (nCells <= 0 || n <= 0) return;
    const size_t smem = need.bytes();
    const bool small = need.pitch() == FT_PITCH_SMALL;
    if (smem > 48 * 1024) {      // a request beyond the device's opt-in maximum fails at the launch below (cudaGetLastError)
        if (small) ensure_max_dynamic_smem(c.device, (const void*)fast_cells_kernel<FT_PITCH_SMALL>, SMEM_FAST_SMALL);
        else ensure_max_dynamic_smem(c.device, (const void*)fast_cells_kernel<FT_PITCH>, SMEM_FAST);
    }
    dim3 grid(nCells, n);
    if (small)
        fast_cells_kernel<FT_PITCH_SMALL><<<grid, FT_THREADS, smem, c.stream>>>(d_pyr, pyrBytes, candPerImg, minTh, iniTh, d_cells, d_cand, d_candCount,
                                                                               need.tileWords, need.scrWords, need.clistCap, need.workCap, passes);
    else
        fast_cells_kernel<FT_PITCH><<<grid, FT_THREADS, smem, c.stream>>>(d_pyr, pyrBytes, candPerImg, minTh, iniTh, d_cells, d_cand, d_candCount,
                                                                         need.tileWords, need.scrWords, need.clistCap, need.workCap, passes);
    c.launches++;
}

void push_fast_group(std::vector<int4>& groups, FastStripSmem& need, int x0, int y0, int x1, int y1, int level, int cells, int wCell,
                     unsigned levelOff, int pitch, unsigned candOff, int candCap)
{
    groups.push_back(make_int4(x0 | (y0 << 16), x1 | (y1 << 16), level, cells));
    groups.push_back(make_int4((int)levelOff, pitch, (int)candOff, candCap));
    const int th = y1 - y0, hi = th - 6, xa = x0 & ~3;
    const int nw = (x1 - xa + 3) >> 2;
    groups.push_back(make_int4(wCell, (int)(0xffffffffu / (uint32_t)std::max(nw, 1) + 1u), 0, 0));
    const int firstIn = x0 + 3, endIn = x1 - 3;
    if (hi <= 0 || endIn <= firstIn) return;
    const int npr = ((endIn - 1 - xa) >> 1) - ((firstIn - xa) >> 1) + 1;
    int clist = 0;
    for (int k = 0; k < cells; k++) {
        const int wi = std::min(firstIn + (k + 1) * wCell, endIn) - (firstIn + k * wCell);
        if (wi > 0) clist += ((wi + 1) / 2) * ((hi + 1) / 2);
    }
    need.tileRows = std::max(need.tileRows, th);
    need.scrRows = std::max(need.scrRows, hi + 2);
    // a warp owns the row blocks w, w + FS_WARPS, ...: at most ceil(blocks / FS_WARPS) * FS_ROWS rows of npr pairs
    const int blocks = (hi + FS_ROWS - 1) / FS_ROWS;
    need.segCap = std::max(need.segCap, ((((blocks + FS_WARPS - 1) / FS_WARPS) * FS_ROWS * npr + 1) & ~1) + 2);
    need.clistCap = std::max(need.clistCap, clist);
}

void launch_fast_strips(Ctx& c, const uint8_t* d_pyr, unsigned pyrBytes, unsigned candPerImg, int minTh, int iniTh,
                        const int4* d_groups, int nGroups, const FastStripSmem& need, uint32_t* d_cand, int32_t* d_candCount, int n,
                        cudaStream_t stream, bool pdl)
{
    if (nGroups <= 0 || n <= 0) return;
    const size_t smem = need.bytes();
    if (smem > 48 * 1024) ensure_max_dynamic_smem(c.device, (const void*)fast_strip_kernel<FS_PITCH>, SMEM_FAST_STRIP);
    dim3 grid(nGroups, n);
    launch_chain(pdl, fast_strip_kernel<FS_PITCH>, grid, dim3(FS_THREADS), smem, stream ? stream : c.stream, d_pyr, pyrBytes, candPerImg, minTh, iniTh,
                 d_groups, d_cand, d_candCount, need.tileRows, need.scrRows, need.segCap, need.clistCap);
    c.launches++;
}

// Grid FAST of levels [l0, l1).  The candidate counters must be zero (launch_fast and the few-image chain clear them first).
void launch_fast_levels(Ctx& c, int n, int l0, int l1, cudaStream_t stream, bool afterKernel)
{
    const Geom& g = c.cur->g;
    const ShapeTables& st = *c.cur;
    const int g0 = st.fastGroupBase[l0], g1 = st.fastGroupBase[l1];
    launch_fast_strips(c, c.d_pyr, g.pyrBytes, g.candPerImg, g.minTh, g.iniTh, st.d_groups + 3 * (size_t)g0, g1 - g0, st.fastStrip, c.d_cand, c.d_candCount, n,
                       stream, c.pdl && afterKernel);
}

void launch_clear_counters(Ctx& c, int n)
{
    cudaMemsetAsync(c.d_candCount, 0, sizeof(int32_t) * MAX_LEVELS * n, c.stream);
}

void launch_fast(Ctx& c, int n, bool afterKernel)
{
    const Geom& g = c.cur->g;
    const ShapeTables& st = *c.cur;
    if (!c.fastCells && st.nFastGroups > 0) {
        launch_fast_levels(c, n, 0, g.nlevels, c.stream, afterKernel);
        return;
    }
    launch_fast_cells(c, c.d_pyr, g.pyrBytes, g.candPerImg, g.minTh, g.iniTh, 2, st.d_cells, st.nFastCells, st.fastSmem, c.d_cand,
                      c.d_candCount, n);
}

void launch_octree_levels(Ctx& c, int n, int l0, int l1, cudaStream_t stream, bool afterKernel)
{
    const Geom& g = c.cur->g;
    if (l1 <= l0) return;
    int maxNodes = 2;
    for (int l = 0; l < g.nlevels; l++) maxNodes = std::max(maxNodes, g.lv[l].maxNodes);
    const size_t smem = octree_smem_bytes(maxNodes, 0);
    dim3 grid(n, l1 - l0);
    const bool pdl = c.pdl && afterKernel;
    if (n <= 2) {
        // the SM's whole opt-in shared memory: what the node tables leave holds the candidates (up to 32768 per level)
        const size_t lim = ensure_max_dynamic_smem(c.device, (const void*)octree_kernel<OT_THREADS_FEW>, SMEM_OCTREE_FEW);
        const int smemCand = c.octreeSmemCand && lim > smem + 1024 ? (int)std::min<size_t>(32768, ((lim - smem - 512) / 6) & ~(size_t)7) : 0;
        launch_chain(pdl, octree_kernel<OT_THREADS_FEW>, grid, dim3(OT_THREADS_FEW), octree_smem_bytes(maxNodes, smemCand), stream, g, c.d_cand, c.d_candCount,
                     c.d_nodeOf, c.d_lvlKp, c.d_lvlCount, c.d_status, l0, smemCand);
    } else {
        if (smem > 48 * 1024) ensure_max_dynamic_smem(c.device, (const void*)octree_kernel<OT_THREADS>, SMEM_OCTREE);
        launch_chain(pdl, octree_kernel<OT_THREADS>, grid, dim3(OT_THREADS), smem, stream, g, c.d_cand, c.d_candCount, c.d_nodeOf, c.d_lvlKp, c.d_lvlCount,
                     c.d_status, l0, 0);
    }
    c.launches++;
}

void launch_octree(Ctx& c, int n)
{
    // after fast_strip_kernel on the same stream (the one-CTA-per-cell form is launched plainly, a full dependency is always right)
    launch_octree_levels(c, n, 0, c.cur->g.nlevels, c.stream, true);
}

void launch_describe(Ctx& c, int n, bool afterKernel, const HostMirror* mirror)
{
    const Geom& g = c.cur->g;
    const HostMirror M = mirror ? *mirror : HostMirror{nullptr, nullptr, nullptr, nullptr, nullptr};
    if (n <= 2)
        launch_chain(c.pdl && afterKernel, describe_kernel<DS_KPB_FEW>, dim3((g.kpPerImg + DS_KPB_FEW - 1) / DS_KPB_FEW, n), dim3(DS_WARPS * 32), 0, c.stream, g,
                     c.d_pyr, c.d_blur, c.d_lvlKp, c.d_lvlCount, c.d_kps, c.d_desc, c.d_counts, c.cur->d_dmaps, M);
    else
        launch_chain(c.pdl && afterKernel, describe_kernel<DS_KPB>, dim3((g.kpPerImg + DS_KPB - 1) / DS_KPB, n), dim3(DS_WARPS * 32), 0, c.stream, g,
                     c.d_pyr, c.d_blur, c.d_lvlKp, c.d_lvlCount, c.d_kps, c.d_desc, c.d_counts, c.cur->d_dmaps, M);
    c.launches++;
}

// n images of rowBytes-pitched rows (a multiple of 16) in pinned host memory -> level 0 of the pyramid pool
void launch_import_host(Ctx& c, const uint8_t* h_imgs, size_t imgBytes, int rowBytes, int n, uint8_t* hostPyr)
{
    const Geom& g = c.cur->g;
    const int rowVec = rowBytes / 16, nVec = rowVec * g.h;
    import_host_kernel<<<dim3((nVec + 127) / 128, n), 128, 0, c.stream>>>(h_imgs, (unsigned)imgBytes, rowVec, nVec, c.d_pyr, g.pyrBytes, g.lv[0].off, g.lv[0].pitch, hostPyr);
    c.launches++;
}

}  // namespace orbb200
