// The drop-in call pattern from several host threads at once: T threads, one context each (the reference's Frame::Frame runs
// the left and the right ORBextractor on two threads, src/Frame.cc:124-127; Tracking, LocalMapping and LoopClosing are three more),
// every thread calls orbb200_extract on ONE image per call from pageable memory and has the results on the host before its next
// call.  Prints aggregate calls/s and the wall time per call seen by a thread for T = 1, 2, 4, 8.
//   nvcc -O2 -o concurrent_calls concurrent_calls.cu -I../../include -L../../orb-slam-birdview_b200 -lorbb200 -Xlinker -rpath=...
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include <cstring>

#include <cuda_runtime.h>

#include "orbb200.h"

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv)
{
    const int w = argc > 1 ? atoi(argv[1]) : 1241, h = argc > 2 ? atoi(argv[2]) : 376, nf = argc > 3 ? atoi(argv[3]) : 2000;
    const int reps = argc > 4 ? atoi(argv[4]) : 300;
    // "dev": the image is already on the device and the results stay there (orbb200_extract_device + orbb200_sync): the same kernels
    // without the host side of the call -- tells whether the threads are held up by the device chain or by the copies around it
    const bool dev = argc > 5 && !strcmp(argv[5], "dev");
    const int maxT = 8;
    // one image per thread (different content: no two threads do the same work)
    std::vector<std::vector<uint8_t>> imgs(maxT, std::vector<uint8_t>((size_t)w * h));
    for (int t = 0; t < maxT; t++) {
        unsigned s = 12345u + 977u * t;
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                s = s * 1664525u + 1013904223u;
                const int blk = (((x + 3 * t) / 13) * 7 + ((y + 5 * t) / 11) * 13) % 5;
                imgs[t][(size_t)y * w + x] = (uint8_t)(40 + blk * 40 + ((s >> 24) & 7));
            }
    }
    std::vector<orbb200_ctx*> ctx(maxT, nullptr);
    for (int t = 0; t < maxT; t++)
        if (orbb200_create(&ctx[t], 0, nf, 1.2f, 8, 20, 7, w, h, 2) != 0) { fprintf(stderr, "create: %s\n", orbb200_last_error(nullptr)); return 1; }
    const int cap = orbb200_max_keypoints(ctx[0]);
    std::vector<uint8_t*> dimg(maxT, nullptr);
    if (dev)
        for (int t = 0; t < maxT; t++) {
            cudaMalloc(&dimg[t], imgs[t].size());
            cudaMemcpy(dimg[t], imgs[t].data(), imgs[t].size(), cudaMemcpyHostToDevice);
        }
    std::vector<int> first(maxT, -1);
    printf("{\"w\": %d, \"h\": %d, \"nfeatures\": %d, \"calls_per_thread\": %d, \"mode\": \"%s\", \"threads\": {", w, h, nf, reps, dev ? "device-resident" : "host");
    bool ok = true;
    for (int T = 1, k = 0; T <= maxT; T *= 2, k++) {
        std::atomic<int> ready{0}, bad{0};
        std::atomic<bool> go{false};
        std::vector<double> perCall(T, 0.0);
        std::vector<std::thread> th;
        double t0 = 0;
        for (int t = 0; t < T; t++)
            th.emplace_back([&, t] {
                std::vector<orbb200_kp_t> kps(cap);
                std::vector<uint8_t> desc((size_t)cap * 32);
                int n = 0;
                for (int i = 0; i < 5; i++)
                    if (orbb200_extract(ctx[t], imgs[t].data(), w, h, w, kps.data(), desc.data(), cap, &n) != 0) bad++;
                if (first[t] < 0) first[t] = n;
                for (int i = 0; dev && i < 5; i++) { orbb200_extract_device(ctx[t], dimg[t], imgs[t].size(), 1, w, h, w); orbb200_sync(ctx[t]); }
                ready++;
                while (!go.load(std::memory_order_acquire)) std::this_thread::yield();
                const double s0 = now_ms();
                for (int i = 0; i < reps; i++) {
                    if (dev) {
                        if (orbb200_extract_device(ctx[t], dimg[t], imgs[t].size(), 1, w, h, w) != 0 || orbb200_sync(ctx[t]) != 0) bad++;
                        continue;
                    }
                    if (orbb200_extract(ctx[t], imgs[t].data(), w, h, w, kps.data(), desc.data(), cap, &n) != 0) bad++;
                    if (n != first[t]) bad++;                   // the same image gives the same keypoints whatever runs beside it
                }
                perCall[t] = (now_ms() - s0) / reps;
            });
        while (ready.load() < T) std::this_thread::yield();
        t0 = now_ms();
        go.store(true, std::memory_order_release);
        for (auto& x : th) x.join();
        const double wall = now_ms() - t0;
        double mean = 0;
        for (double v : perCall) mean += v / T;
        printf("%s\"%d\": {\"calls_per_s\": %.0f, \"ms_per_call_seen_by_a_thread\": %.4f}", k ? ", " : "", T, 1000.0 * T * reps / wall, mean);
        if (bad.load()) ok = false;
    }
    printf("}, \"keypoints\": %d, \"consistent\": %s}\n", first[0], ok ? "true" : "false");
    for (auto c : ctx) orbb200_destroy(c);
    return ok ? 0 : 2;
}
