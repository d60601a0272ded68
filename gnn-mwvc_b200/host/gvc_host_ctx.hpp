// gvc_host_ctx.hpp -- process-wide libgvc context for the drop-in translation
// units.  The reference's classes have no room for a device handle and their
// headers must stay untouched, so the GPU state lives here,
// created lazily at first use.  Environment:
//   GVC_DEVICE  CUDA ordinal (default 0)
//   GVC_MODE    "exact" (default, bit-identical scores) or "fast"
//   GVC_WARM    0: no early context creation on a helper thread (see warm_start)
//   GVC_WARM_MB device memory the helper thread allocates up front (default 1024; the first cudaMalloc of a
//               process took 0.5 - 90 ms on the pool's boxes; 0: none, the first predict allocates what it needs)
//   GVC_DEVICES "0,1,2,3": predict() shards graphs of at least GVC_MULTI_MIN_VERTICES vertices
//               (default 2 000 000) over these devices (gvc_group, include/gvc.h); smaller graphs and
//               everything else stay on the first of them
// Errors follow the reference driver's convention of print-and-stop: the API is
// void everywhere (include/gnn_inference.hpp:50), so a CUDA failure prints to
// stderr and aborts.  There is no CPU fallback.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "gvc.h"

namespace gvc_host {

[[noreturn]] inline void die(const char *where, int rc) {
    std::fprintf(stderr, "gvc: %s failed (%d): %s\n", where, rc, gvc_last_error());
    std::abort();
}

inline const std::vector<int> &devices() {
    static const std::vector<int> list = [] {
        std::vector<int> v;
        if (const char *e = std::getenv("GVC_DEVICES")) {
            for (const char *p = e; *p;) {
                char *end = nullptr;
                const long d = std::strtol(p, &end, 10);
                if (end == p) break;
                v.push_back((int)d);
                p = *end == ',' ? end + 1 : end;
            }
        }
        return v;
    }();
    return list;
}

// The context is created at first use -- or ahead of time by warm_start(): CUDA initialisation (a few
// hundred ms in a fresh process), the pinning of the upload ring (about 8 ms) and the first device allocation
// (0.5 - 90 ms, measured) then happen on a helper
// thread while the solver still parses and reduces its graph, instead of inside the first predict().
inline gvc_ctx *context_impl(bool fatal, bool warm) {
    static std::mutex mu;
    static gvc_ctx *ctx = nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!ctx) {
        const char *dev = std::getenv("GVC_DEVICE");
        gvc_ctx *c = nullptr;
        const int ordinal = dev ? std::atoi(dev) : (devices().empty() ? 0 : devices()[0]);
        const int rc = gvc_ctx_create(&c, ordinal);
        if (rc != 0) {
            if (fatal) die("gvc_ctx_create", rc);
            return nullptr;                    // the helper thread leaves the complaint to the call that needs the GPU
        }
        if (warm) {                            // best effort: the pinned ring and a first chunk of device memory
            const char *mb = std::getenv("GVC_WARM_MB");
            const unsigned long long warm_mb = mb ? std::strtoull(mb, nullptr, 10) : 1024ull;
            gvc_ctx_warm(c, warm_mb << 20);
        }
        ctx = c;
    }
    return ctx;
}

inline gvc_ctx *context() { return context_impl(true, false); }

// Called where the solver builds its model (operator>>, src/GNN_VC.cpp:263 -- before it parses the graph).
// GVC_WARM=0 turns it off.  The thread is joined at exit.
inline void warm_start() {
    struct helper {
        std::thread t;
        ~helper() { if (t.joinable()) t.join(); }
    };
    static std::once_flag once;
    static helper h;
    std::call_once(once, [] {
        const char *e = std::getenv("GVC_WARM");
        if (e && e[0] == '0') return;
        h.t = std::thread([] { context_impl(false, true); });
    });
}

// the group of GVC_DEVICES (null when fewer than two are named)
inline gvc_group *group() {
    static gvc_group *grp = [] {
        gvc_group *g = nullptr;
        if (devices().size() >= 2) {
            const int rc = gvc_group_create(&g, devices().data(), (int)devices().size());
            if (rc != 0) die("gvc_group_create", rc);
        }
        return g;
    }();
    return grp;
}

inline bool shard_over_devices(uint64_t n_vertices) {
    static const uint64_t min_n = [] {
        const char *e = std::getenv("GVC_MULTI_MIN_VERTICES");
        return e ? std::strtoull(e, nullptr, 10) : 2000000ull;
    }();
    return devices().size() >= 2 && n_vertices >= min_n;
}

inline int mode() {
    const char *m = std::getenv("GVC_MODE");
    return (m && std::strcmp(m, "fast") == 0) ? GVC_MODE_FAST : GVC_MODE_EXACT;
}

}  // namespace gvc_host
