// gvc_host_ctx.hpp -- process-wide libgvc context for the drop-in translation
// units.  The reference's classes have no room for a device handle and their
// headers must stay untouched, so the GPU state lives here,
// created lazily at first use.  Environment:
//   GVC_DEVICE  CUDA ordinal (default 0)
//   GVC_MODE    "exact" (default, bit-identical scores) or "fast"
// Errors follow the reference driver's convention of print-and-stop: the API is
// void everywhere (include/gnn_inference.hpp:50), so a CUDA failure prints to
// stderr and aborts.  There is no CPU fallback.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "gvc.h"

namespace gvc_host {

[[noreturn]] inline void die(const char *where, int rc) {
    std::fprintf(stderr, "gvc: %s failed (%d): %s\n", where, rc, gvc_last_error());
    std::abort();
}

inline gvc_ctx *context() {
    static gvc_ctx *ctx = [] {
        const char *dev = std::getenv("GVC_DEVICE");
        gvc_ctx *c = nullptr;
        const int rc = gvc_ctx_create(&c, dev ? std::atoi(dev) : 0);
        if (rc != 0) die("gvc_ctx_create", rc);
        return c;
    }();
    return ctx;
}

inline int mode() {
    const char *m = std::getenv("GVC_MODE");
    return (m && std::strcmp(m, "fast") == 0) ? GVC_MODE_FAST : GVC_MODE_EXACT;
}

}  // namespace gvc_host
