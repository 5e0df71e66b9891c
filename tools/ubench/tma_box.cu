// Micro-test: fetch a W x 39-byte box of a pitched u8 image with cp.async.bulk.tensor and compare with the source.
// usage: tma_box   (prints one line per variant)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

template <int RANK>
__device__ __forceinline__ void fetch_body(const CUtensorMap* map, int cx, int cy, int cz, int bytes, uint8_t* out)
{
    __shared__ __align__(128) uint8_t box[64 * 39 + 128];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(box);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(bytes) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         :: "r"(d), "l"(reinterpret_cast<uint64_t>(map)), "r"(cx), "r"(cy), "r"(cz), "r"(b) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         :: "r"(d), "l"(reinterpret_cast<uint64_t>(map)), "r"(cx), "r"(cy), "r"(b) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" :: "r"(b), "r"(0) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = box[i];
}

template <int RANK>
__global__ void fetch_kernel(const CUtensorMap* map, int cx, int cy, int cz, int bytes, uint8_t* out) { fetch_body<RANK>(map, cx, cy, cz, bytes, out); }
template <int RANK>
__global__ void fetch_kernel_gc(const __grid_constant__ CUtensorMap map, int cx, int cy, int cz, int bytes, uint8_t* out) { fetch_body<RANK>(&map, cx, cy, cz, bytes, out); }

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv)
{
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    const bool gc = argc > 2;
    int vi = -1;
    const int pitch = 1280, rows = 400, nimg = 3, W0 = 1248;
    const size_t imgBytes = (size_t)pitch * rows;
    std::vector<uint8_t> h(imgBytes * nimg);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t *d, *dout;
    cudaMalloc(&d, h.size() + 4096); cudaMalloc(&dout, 8192);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (!fn) { printf("no encoder\n"); return 1; }
    CUtensorMap* dmap; cudaMalloc(&dmap, sizeof(CUtensorMap));
    struct V { int rank, bw, off; } vs[] = {{3, 48, 32}, {2, 48, 32}, {3, 64, 32}, {2, 64, 32}, {3, 48, 0}, {3, 32, 32}, {2, 16, 0}, {2, 128, 0}};
    for (V v : vs) {
        vi++;
        if (only >= 0 && vi != only) continue;
        CUtensorMap m; memset(&m, 0, sizeof(m));
        const cuuint64_t dims[3] = {(cuuint64_t)(W0 - (v.off ? 0 : 0)), (cuuint64_t)rows, (cuuint64_t)nimg};
        const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)imgBytes};
        const cuuint32_t box[3] = {(cuuint32_t)v.bw, 39, 1}, es[3] = {1, 1, 1};
        CUresult r = ((EncodeTiled)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, v.rank, d + v.off, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("rank %d box %dx39 base+%d: encode=%d ", v.rank, v.bw, v.off, (int)r);
        if (r != CUDA_SUCCESS) { printf("\n"); continue; }
        cudaMemcpy(dmap, &m, sizeof(m), cudaMemcpyHostToDevice);
        const int cx = argc > 3 ? atoi(argv[3]) : 101, cy = 57, cz = v.rank == 3 ? 2 : 0, bytes = v.bw * 39;
        if (gc) { if (v.rank == 3) fetch_kernel_gc<3><<<1, 64>>>(m, cx, cy, cz, bytes, dout); else fetch_kernel_gc<2><<<1, 64>>>(m, cx, cy, cz, bytes, dout); }
        else { if (v.rank == 3) fetch_kernel<3><<<1, 64>>>(dmap, cx, cy, cz, bytes, dout); else fetch_kernel<2><<<1, 64>>>(dmap, cx, cy, cz, bytes, dout); }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<uint8_t> o(bytes);
        cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int y = 0; y < 39; y++) for (int x = 0; x < v.bw; x++) bad += o[y * v.bw + x] != h[(size_t)cz * imgBytes + (size_t)(cy + y) * pitch + v.off + cx + x];
        printf("mismatches=%d\n", bad);
    }
    return 0;
}
