// Where one ORBextractor::operator() call spends its time (one image per call, results on the host before the next call).
//   nvcc -O2 -o call_timeline call_timeline.cu -I../../include -L../../orb-slam-birdview_b200 -lorbb200 -Xlinker -rpath=...
// Legs, wall time per call unless stated:
//   dev_events  : device-resident image, CUDA events around orbb200_extract_device on the context's stream (pure device chain)
//   dev_wall    : the same call + orbb200_sync (adds the graph launch and the wake-up of the waiting thread)
//   host_full   : orbb200_extract from pageable memory (staging memcpy + H2D + chain + result D2H + copy-out)
//   memcpy_only : the host memcpy of the image into a pinned block (what staging costs)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "orbb200.h"

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv)
{
    const int w = argc > 1 ? atoi(argv[1]) : 752, h = argc > 2 ? atoi(argv[2]) : 480, nf = argc > 3 ? atoi(argv[3]) : 1000;
    const int reps = argc > 4 ? atoi(argv[4]) : 300;
    std::vector<uint8_t> img((size_t)w * h);
    unsigned s = 12345;
    // blocky texture with corners (not the bench image, only a load with a realistic keypoint count)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            s = s * 1664525u + 1013904223u;
            const int blk = ((x / 13) * 7 + (y / 11) * 13) % 5;
            img[(size_t)y * w + x] = (uint8_t)(40 + blk * 40 + ((s >> 24) & 7));
        }
    orbb200_ctx* ctx = nullptr;
    if (orbb200_create(&ctx, 0, nf, 1.2f, 8, 20, 7, w, h, 2) != 0) { fprintf(stderr, "create: %s\n", orbb200_last_error(nullptr)); return 1; }
    const int cap = orbb200_max_keypoints(ctx);
    std::vector<orbb200_kp_t> kps(cap);
    std::vector<uint8_t> desc((size_t)cap * 32);
    int n = 0;
    uint8_t* d_img = nullptr;
    cudaMalloc(&d_img, img.size());
    cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice);
    cudaStream_t st = (cudaStream_t)orbb200_stream(ctx);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 5; i++) {
        if (orbb200_extract(ctx, img.data(), w, h, w, kps.data(), desc.data(), cap, &n) != 0) { fprintf(stderr, "extract: %s\n", orbb200_last_error(ctx)); return 1; }
        orbb200_extract_device(ctx, d_img, img.size(), 1, w, h, w); orbb200_sync(ctx);
    }
    double devEv = 0;
    for (int i = 0; i < reps; i++) {
        cudaEventRecord(e0, st);
        orbb200_extract_device(ctx, d_img, img.size(), 1, w, h, w);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); devEv += ms;
    }
    double t0 = now_ms();
    for (int i = 0; i < reps; i++) { orbb200_extract_device(ctx, d_img, img.size(), 1, w, h, w); orbb200_sync(ctx); }
    const double devWall = (now_ms() - t0) / reps;
    t0 = now_ms();
    for (int i = 0; i < reps; i++) orbb200_extract(ctx, img.data(), w, h, w, kps.data(), desc.data(), cap, &n);
    const double full = (now_ms() - t0) / reps;
    uint8_t* pin = nullptr;
    cudaMallocHost(&pin, img.size());
    t0 = now_ms();
    for (int i = 0; i < reps; i++) { memcpy(pin, img.data(), img.size()); asm volatile("" ::: "memory"); }
    const double mc = (now_ms() - t0) / reps;
    // an empty stream round trip: one tiny async copy + sync (the floor of "enqueue something and wait for it")
    int32_t* d_x = nullptr; cudaMalloc(&d_x, 4);
    int32_t* h_x = nullptr; cudaMallocHost(&h_x, 4);
    t0 = now_ms();
    for (int i = 0; i < reps; i++) { cudaMemcpyAsync(h_x, d_x, 4, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st); }
    const double rt = (now_ms() - t0) / reps;
    t0 = now_ms();
    for (int i = 0; i < reps; i++) { cudaMemcpyAsync(d_img, pin, img.size(), cudaMemcpyHostToDevice, st); cudaStreamSynchronize(st); }
    const double up = (now_ms() - t0) / reps;
    printf("{\"w\": %d, \"h\": %d, \"keypoints\": %d, \"dev_events_ms\": %.4f, \"dev_wall_ms\": %.4f, \"host_full_ms\": %.4f, \"memcpy_only_ms\": %.4f, "
           "\"tiny_copy_round_trip_ms\": %.4f, \"pinned_upload_round_trip_ms\": %.4f}\n",
           w, h, n, devEv / reps, devWall, full, mc, rt, up);
    orbb200_destroy(ctx);
    return 0;
}
