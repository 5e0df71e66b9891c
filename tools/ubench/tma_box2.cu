// TMA box fetch through libcu++'s experimental API (the sequence of the CUDA programming guide), 2-D, __grid_constant__ map.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;

constexpr int BW = 64, BH = 39;
__global__ void k(const __grid_constant__ CUtensorMap map, int cx, int cy, uint8_t* out)
{
    __shared__ alignas(128) uint8_t box[BW * BH];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&box, &map, cx, cy, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(box));
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = box[i];
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv)
{
    const int pitch = 1280, rows = 400;
    std::vector<uint8_t> h((size_t)pitch * rows);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t *d, *dout;
    cudaMalloc(&d, h.size()); cudaMalloc(&dout, 8192);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("entry point: %s q=%d fn=%p\n", cudaGetErrorString(ge), (int)q, fn);
    CUtensorMap m; memset(&m, 0, sizeof(m));
    const cuuint64_t dims[2] = {1280, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)pitch};
    const cuuint32_t box[2] = {BW, BH}, es[2] = {1, 1};
    CUresult r = ((EncodeTiled)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode=%d map words:", (int)r);
    for (int i = 0; i < 16; i++) printf(" %016llx", (unsigned long long)((uint64_t*)&m)[i]);
    printf("\n");
    const int CX = argc > 1 ? atoi(argv[1]) : 96;
    k<<<1, 64>>>(m, CX, 57, dout);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint8_t> o(BW * BH);
    cudaMemcpy(o.data(), dout, o.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < BH; y++) for (int x = 0; x < BW; x++) bad += o[y * BW + x] != h[(size_t)(57 + y) * pitch + CX + x];
    printf("mismatches=%d\n", bad);
    return 0;
}
