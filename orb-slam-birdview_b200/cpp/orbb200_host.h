// Host-side helpers shared by the C++ shims: the device context of the calling thread and the reference-style error
// handling (print and abort, like the assert()/exit(-1) of src/ORBextractor.cc:1050 and src/System.cc:59-84).
// ORBmatcher objects are stack temporaries in the reference (src/Tracking.cc:738,1029,1192), so the context cannot live
// in them: each thread that matches (Tracking, LocalMapping, LoopClosing) gets its own, created on first use or handed
// over from that thread's ORBextractor with SetThreadContext(extractor.Context()).
#pragma once
#include <cstdio>
#include <cstdlib>

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

inline orbb200_ctx* ThreadContext()
{
    if (!slot()) {
        // matcher-only context: the extraction pools are sized for a 64x64 image and never used; the scale tables are the
        // reference defaults of every shipped YAML (ORBextractor.scaleFactor 1.2, nLevels 8)
        orbb200_ctx* c = nullptr;
        const int rc = orbb200_create(&c, 0, 1000, 1.2f, 8, 20, 7, 64, 64, 1);
        if (rc != ORBB200_OK) { fprintf(stderr, "orbb200_create failed (%d): %s\n", rc, orbb200_last_error(nullptr)); abort(); }
        slot() = c;
    }
    return slot();
}
}  // namespace orbb200_host
