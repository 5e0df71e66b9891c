// Micro-benchmark: FP64 pipe of this GPU -- issue rate of DADD / DMUL / DFMA / F2F.F64.F32 with 8 independent chains per
// thread, and the latency of a dependent DADD chain (one chain per thread, one warp per SM sub-partition).
// VERDICT r1 asked for this: cornerSubPix accumulates five sums in double, in order; its floor is either the FP64 issue rate
// or the dependent-add latency.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64 fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE> __global__ void tput(double* out, int iters, double seed)
{
    double a[8], b = seed + threadIdx.x * 1e-9, c = 1.0 + seed * 1e-7;
    float f[8];
    for (int i = 0; i < 8; i++) { a[i] = seed * (i + 1) + threadIdx.x; f[i] = (float)(seed + i + threadIdx.x); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) a[i] = __dadd_rn(a[i], b);
            if (MODE == 1) a[i] = __dmul_rn(a[i], c);
            if (MODE == 2) a[i] = __fma_rn(a[i], c, b);
            if (MODE == 3) { a[i] = __dadd_rn(a[i], (double)f[i]); }                    // F2F.F64.F32 + DADD
            if (MODE == 4) { f[i] = (float)__dmul_rn((double)f[i], c); }                // F2F up + DMUL + F2F down (getRectSubPix's prev = t*s)
        }
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += a[i] + f[i];
    if (s == 0.12345) out[0] = s;
}

__global__ void latency(double* out, int iters, double seed, long long* cycles)
{
    double a = seed + threadIdx.x;
    const double b = seed * 0.5;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) a = __dadd_rn(a, b);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
    if (a == 0.12345) out[0] = a;
}

template <int MODE> float run(double* d, int iters, int blocks, int threads)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    tput<MODE><<<blocks, threads>>>(d, 16, 1.5);
    float best = 1e9;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); tput<MODE><<<blocks, threads>>>(d, iters, 1.5); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    return best;
}

int main()
{
    double* d; cudaMalloc(&d, 8);
    long long* cyc; cudaMalloc(&cyc, 8);
    int sms = 148, mhz = 1965;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0); mhz /= 1000;
    const int iters = 2048, blocks = sms * 4, threads = 256;
    const double ops = (double)blocks * threads * iters * 8;
    const char* names[5] = {"DADD", "DMUL", "DFMA", "F2F.F64.F32 + DADD (2 ops)", "F2F up + DMUL + F2F down (3 ops)"};
    float t[5] = {run<0>(d, iters, blocks, threads), run<1>(d, iters, blocks, threads), run<2>(d, iters, blocks, threads),
                  run<3>(d, iters, blocks, threads), run<4>(d, iters, blocks, threads)};
    const int per[5] = {1, 1, 1, 2, 3};
    for (int m = 0; m < 5; m++) {
        const double rate = ops * per[m] / (t[m] * 1e-3);
        printf("%-36s: %8.3f ms  %8.1f Gop/s  = %.2f lanes/clk/SM at %d MHz nominal\n", names[m], t[m], rate * 1e-9, rate / sms / (mhz * 1e6), mhz);
    }
    // occupancy sweep for DADD: does the rate depend on resident warps (latency-bound) or not (pipe-bound)?
    for (int th = 32; th <= 1024; th *= 2) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        tput<0><<<sms, th>>>(d, 16, 1.5);
        cudaEventRecord(e0); tput<0><<<sms, th>>>(d, iters, 1.5); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("DADD, 1 CTA/SM x %4d threads (8 chains each): %.2f lanes/clk/SM\n", th, (double)sms * th * iters * 8 / (ms * 1e-3) / sms / (mhz * 1e6));
    }
    latency<<<1, 32>>>(d, 256, 1.5, cyc);
    latency<<<1, 32>>>(d, 4096, 1.5, cyc);
    long long h = 0; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent DADD chain: %.2f cycles per add (1 warp)\n", (double)h / (4096.0 * 16));
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
