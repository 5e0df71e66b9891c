// Host-side helpers shared by the C++ shims: the device context of the calling thread, device-frame caching and the
// reference-style error handling (print and abort, like the assert()/exit(-1) of src/ORBextractor.cc:1050 and
// src/System.cc:59-84).  ORBmatcher objects are stack temporaries in the reference (src/Tracking.cc:738,1029,1192), so the
// context cannot live in them: each thread that matches (Tracking, LocalMapping, LoopClosing) gets its own, created on first
// use with the scale pyramid of the frame it is first used on, or handed over from that thread's ORBextractor with
// SetThreadContext(extractor.Context()).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/orbb200.h"

namespace orbb200_host
{
inline orbb200_ctx*& slot() { static thread_local orbb200_ctx* ctx = nullptr; return ctx; }

inline void check(int rc, const char* what)
{
    if (rc == ORBB200_OK) return;
    fprintf(stderr, "orbb200: %s failed (%d): %s\n", what, rc, orbb200_last_error(slot()));
    abort();                                                   // no CPU fallback
}

inline void SetThreadContext(orbb200_ctx* ctx) { slot() = ctx; }

// The windowed searches scale their radius with the context's mvScaleFactors table (th * mvScaleFactors[level],
// src/ORBmatcher.cc:72,1385): the table must be the frame's own.  A matcher-only context is created from the first frame's
// pyramid; afterwards every frame / keyframe handed to a shim is checked against it (a different
// ORBextractor.scaleFactor or nLevels must not silently change the matches).
inline orbb200_ctx* ThreadContext(const std::vector<float>& scaleFactors)
{
    const int nlevels = (int)scaleFactors.size();
    if (!slot()) {
        if (nlevels < 1 || nlevels > 12) { fprintf(stderr, "orbb200: frame with %d pyramid levels\n", nlevels); abort(); }
        // matcher-only context: the extraction pools are sized for a 64x64 image and never used
        const float sf = nlevels > 1 ? scaleFactors[1] / scaleFactors[0] : 1.2f;
        orbb200_ctx* c = nullptr;
        const int rc = orbb200_create(&c, 0, 1000, sf, nlevels, 20, 7, 64, 64, 1);
        if (rc != ORBB200_OK) { fprintf(stderr, "orbb200_create failed (%d): %s\n", rc, orbb200_last_error(nullptr)); abort(); }
        slot() = c;
    }
    static thread_local std::vector<float> verified;
    if (verified != scaleFactors) {
        std::vector<float> mine(12);
        const int n = orbb200_get_levels(slot());
        bool ok = n == nlevels && orbb200_get_scale_table(slot(), 0, mine.data()) == ORBB200_OK;
        for (int i = 0; ok && i < nlevels; i++) ok = mine[i] == scaleFactors[i];
        if (!ok) {
            fprintf(stderr, "orbb200: this thread's context was created for another scale pyramid (%d levels) than the frame's (%d levels, factor %g)\n",
                    n, nlevels, nlevels > 1 ? scaleFactors[1] : 1.f);
            abort();
        }
        verified = scaleFactors;
    }
    return slot();
}

// Device copy of a frame's keypoints + descriptors (+ uRight) and its 64x48 lookup grid.  Tracking calls three to five
// matchers on the same Frame (src/Tracking.cc:1227-1245, 1659-1672): the upload + grid build is done once and found again by
// content (a 64-bit hash of the keypoints and descriptors; frames are rebuilt at the same address every image, so an address
// or id alone would not do).  In the reference tree the handle would simply be a member of Frame.
struct FrameCache
{
    struct Entry { orbb200_ctx* ctx; uint64_t key; int n; float g[4]; orbb200_frame* h; unsigned long stamp; };
    std::vector<Entry> e;
    unsigned long clock = 0;
    ~FrameCache() { for (size_t i = 0; i < e.size(); i++) if (e[i].h) orbb200_frame_free(e[i].h); }

    static uint64_t hash(const void* p, size_t n, uint64_t h)
    {
        // four interleaved multiply-xor lanes over 32-byte blocks (a single lane is a chain of dependent multiplies: 20 us for the
        // 120 KB of a 2000-keypoint frame, every matcher call), folded into one value at the end
        const unsigned char* b = (const unsigned char*)p;
        const uint64_t M = 0x100000001b3ull;
        uint64_t h0 = h, h1 = h ^ 0x9e3779b97f4a7c15ull, h2 = h + 0x7f4a7c159e3779b9ull, h3 = ~h;
        size_t i = 0;
        for (; i + 32 <= n; i += 32) {
            uint64_t w[4];
            memcpy(w, b + i, 32);
            h0 = (h0 ^ w[0]) * M; h0 ^= h0 >> 29;
            h1 = (h1 ^ w[1]) * M; h1 ^= h1 >> 29;
            h2 = (h2 ^ w[2]) * M; h2 ^= h2 >> 29;
            h3 = (h3 ^ w[3]) * M; h3 ^= h3 >> 29;
        }
        h = (((h0 * M ^ h1) * M ^ h2) * M ^ h3) * M;
        for (; i + 8 <= n; i += 8) { uint64_t w; memcpy(&w, b + i, 8); h = (h ^ w) * M; h ^= h >> 29; }
        for (; i < n; i++) h = (h ^ b[i]) * M;
        return h;
    }

    // keys: n keypoints of 28 bytes; desc: n rows of 32 bytes, contiguous
    orbb200_frame* get(orbb200_ctx* ctx, const void* keys, const unsigned char* desc, const float* uRight, int n, float minX, float minY, float invW, float invH)
    {
        uint64_t k = hash(keys, (size_t)n * 28, 0xcbf29ce484222325ull);
        k = hash(desc, (size_t)n * 32, k);
        if (uRight) k = hash(uRight, (size_t)n * 4, k);
        const float g[4] = {minX, minY, invW, invH};
        for (size_t i = 0; i < e.size(); i++)
            if (e[i].ctx == ctx && e[i].key == k && e[i].n == n && memcmp(e[i].g, g, sizeof(g)) == 0) { e[i].stamp = ++clock; return e[i].h; }
        if (e.size() >= 8) {                                    // evict the least recently used
            size_t o = 0;
            for (size_t i = 1; i < e.size(); i++) if (e[i].stamp < e[o].stamp) o = i;
            orbb200_frame_free(e[o].h);
            e.erase(e.begin() + o);
        }
        Entry x; x.ctx = ctx; x.key = k; x.n = n; memcpy(x.g, g, sizeof(g)); x.h = nullptr; x.stamp = ++clock;
        check(orbb200_frame_upload(ctx, &x.h, (const orbb200_kp_t*)keys, desc, uRight, n, minX, minY, invW, invH), "orbb200_frame_upload");
        e.push_back(x);
        return x.h;
    }
};
inline FrameCache& frames() { static thread_local FrameCache c; return c; }
}  // namespace orbb200_host
