// mock_gvc.cpp -- TEST INFRASTRUCTURE ONLY: a stand-in for libgvc's entry points that does NO
// arithmetic.  tests/test_dropin_extract_cpu.py links it with the drop-in host units
// (gnn-mwvc_b200/host/*.cpp) so that the HOST side of gnn::model::predict -- reading a
// reduction_graph through begin(u)/end(u)/W/NW and describing it to gvc_graph_upload_stream through
// callbacks that several threads call concurrently -- can be checked on a machine without a GPU: the
// mock drives the callbacks the way libgvc does (vertex chunks and span chunks handed to worker
// threads in arbitrary order, each into its own scratch slot), rebuilds the packed CSR from what it
// received and hands it back to the test.  Nothing here is ever part of the product.
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "gvc.h"

namespace {
std::vector<uint64_t> g_row_ptr;
std::vector<uint32_t> g_col, g_w, g_nw;
uint64_t g_span_len = 0;
int g_streamed = 0;
std::vector<float> g_x;          // the input that travelled with the last streamed upload
bool g_have_x = false;
float g_scales[8];
int g_n_scales = 0;
}  // namespace

struct gvc_ctx { int dummy; };

extern "C" {
const char *gvc_last_error(void) { return "mock"; }
int gvc_ctx_create(gvc_ctx **out, int) { static gvc_ctx c; *out = &c; return 0; }
int gvc_ctx_warm(gvc_ctx *, uint64_t) { return 0; }
int gvc_model_upload(gvc_ctx *, int, const int *, const int *, const int *, const float *const *, const float *const *) { return 0; }
int gvc_model_weight_scales(gvc_ctx *, int n, const float *s) { g_n_scales = n; for (int i = 0; i < n && i < 8; ++i) g_scales[i] = s[i]; return 0; }

int gvc_graph_upload_stream_x(gvc_ctx *, uint32_t n, uint64_t span_len, gvc_fill_vertices_fn fv, gvc_fill_span_fn fs, void *user, int, const float *x) {
    g_x.assign(x ? x : nullptr, x ? x + n : nullptr);
    g_have_x = x != nullptr;
    constexpr uint32_t kV = 1000;            // deliberately odd chunk sizes
    constexpr uint64_t kS = 777;
    std::vector<uint32_t> b(n), e(n), span(span_len);
    g_w.assign(n, 0); g_nw.assign(n, 0);
    const uint64_t nv = (n + kV - 1) / kV, ns = (span_len + kS - 1) / kS;
    std::atomic<uint64_t> next{0};
    auto work = [&](int w) {
        std::vector<uint32_t> slot(4 * kV > kS ? 4 * kV : kS);
        for (;;) {
            // items are taken from both ends alternately so that span pieces arrive before vertex pieces too
            const uint64_t k = next.fetch_add(1);
            if (k >= nv + ns) break;
            const uint64_t it = (k & 1) ? (nv + ns - 1 - k / 2) : k / 2;
            if (it < nv) {
                const uint32_t first = (uint32_t)(it * kV), cnt = std::min<uint32_t>(kV, n - first);
                std::memset(slot.data(), 0xAB, slot.size() * 4);
                fv(user, first, cnt, slot.data(), slot.data() + kV, slot.data() + 2 * kV, slot.data() + 3 * kV);
                std::memcpy(b.data() + first, slot.data(), cnt * 4);
                std::memcpy(e.data() + first, slot.data() + kV, cnt * 4);
                std::memcpy(g_w.data() + first, slot.data() + 2 * kV, cnt * 4);
                std::memcpy(g_nw.data() + first, slot.data() + 3 * kV, cnt * 4);
            } else {
                const uint64_t off = (it - nv) * kS, cnt = std::min<uint64_t>(kS, span_len - off);
                std::memset(slot.data(), 0xCD, slot.size() * 4);
                fs(user, off, cnt, slot.data());
                std::memcpy(span.data() + off, slot.data(), cnt * 4);
            }
        }
        (void)w;
    };
    std::vector<std::thread> th;
    for (int w = 0; w < 3; ++w) th.emplace_back(work, w);
    for (auto &t : th) t.join();
    g_row_ptr.assign((size_t)n + 1, 0);
    g_col.clear();
    for (uint32_t u = 0; u < n; ++u) {
        if (b[u] > e[u] || e[u] > span_len) return 1;
        g_row_ptr[u] = g_col.size();
        g_col.insert(g_col.end(), span.begin() + b[u], span.begin() + e[u]);
    }
    g_row_ptr[n] = g_col.size();
    g_span_len = span_len;
    g_streamed = 1;
    return 0;
}
int gvc_graph_upload_stream(gvc_ctx *c, uint32_t n, uint64_t span_len, gvc_fill_vertices_fn fv, gvc_fill_span_fn fs, void *user, int t) {
    return gvc_graph_upload_stream_x(c, n, span_len, fv, fs, user, t, nullptr);
}

int gvc_graph_upload(gvc_ctx *, uint32_t n, const uint64_t *rp, const uint32_t *col, const uint32_t *W, const uint32_t *NW) {
    g_row_ptr.assign(rp, rp + n + 1);
    g_col.assign(col, col + rp[n]);
    g_w.assign(W, W + n);
    g_nw.assign(NW, NW + n);
    g_span_len = rp[n];
    g_streamed = 0;
    g_have_x = false;
    return 0;
}

int gvc_graph_staging(gvc_ctx *, uint32_t, uint64_t, uint64_t **, uint32_t **, uint32_t **, uint32_t **) { return 1; }   // "no pinned buffers": callers fall back to their own
// scores = 0.5 everywhere; called with x == NULL (the input came with the graph) they are 0.5 only where
// that input equals what the test gave predict, so a lost or shuffled x shows up in `out`
int gvc_forward(gvc_ctx *, const float *x, float, float *scores, int) {
    const size_t n = g_row_ptr.empty() ? 0 : g_row_ptr.size() - 1;
    if (!x && (!g_have_x || g_x.size() != n)) return 1;
    for (size_t i = 0; i < n; ++i) scores[i] = 0.5f;
    return 0;
}
uint64_t mock_last_x(float *out) { if (out) std::memcpy(out, g_x.data(), g_x.size() * 4); return g_have_x ? g_x.size() : ~0ull; }
int gvc_graph_layer_host(gvc_ctx *, const float *, int, float *, float) { return 0; }
int gvc_linear_host(gvc_ctx *, uint64_t, int, int, const float *, const float *, const float *, float *, int) { return 0; }
int gvc_relu_host(gvc_ctx *, uint64_t, const float *, float *) { return 0; }
int gvc_sigmoid_host(gvc_ctx *, uint64_t, const float *, float *, int) { return 0; }
int gvc_sgemm_host(gvc_ctx *, int, int, uint64_t, uint64_t, uint64_t, const float *, uint64_t, const float *, uint64_t, float, float *, uint64_t) { return 0; }

// the group entry points exist for the linker only (the mock never shards)
struct gvc_group { int dummy; };
int gvc_group_create(gvc_group **, const int *, int) { return 1; }
void gvc_group_destroy(gvc_group *) {}
int gvc_group_size(const gvc_group *) { return 0; }
int gvc_group_model_upload(gvc_group *, int, const int *, const int *, const int *, const float *const *, const float *const *) { return 1; }
int gvc_group_model_weight_scales(gvc_group *, int, const float *) { return 1; }
int gvc_group_graph_upload(gvc_group *, uint32_t, const uint64_t *, const uint32_t *, const uint32_t *, const uint32_t *) { return 1; }
int gvc_group_forward(gvc_group *, const float *, float, float *, int) { return 1; }

// what the last upload delivered: sizes first (nulls), then the arrays
uint64_t mock_last_graph(uint64_t *row_ptr, uint32_t *col, uint32_t *w, uint32_t *nw, uint64_t *span_len, int *streamed) {
    const size_t n = g_row_ptr.empty() ? 0 : g_row_ptr.size() - 1;
    if (row_ptr) std::memcpy(row_ptr, g_row_ptr.data(), (n + 1) * 8);
    if (col) std::memcpy(col, g_col.data(), g_col.size() * 4);
    if (w) std::memcpy(w, g_w.data(), n * 4);
    if (nw) std::memcpy(nw, g_nw.data(), n * 4);
    if (span_len) *span_len = g_span_len;
    if (streamed) *streamed = g_streamed;
    return g_col.size();
}
int mock_last_scales(float *out) { for (int i = 0; i < g_n_scales && i < 8; ++i) out[i] = g_scales[i]; return g_n_scales; }
}
