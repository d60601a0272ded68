// gvc_dropin_capi.cpp -- C entry points over the reference's C++ interface, for callers that are
// not C++ (bench.py's end-to-end leg, tools/replay, Python users of the drop-in).
//
// Built together with gvc_gnn_inference.cpp + gvc_matrix.cpp against the reference's own headers
// into host/_build/libgvc_dropin.so.  Every call below is the call src/GNN_VC.cpp makes:
//   gvcd_model_create            istringstream(model_data) >> m        src/GNN_VC.cpp:255,263
//   gvcd_model_set_weight_scale  m.set_weight_scale(w_max)             :278
//   gvcd_graph_create            reduction_graph ctor from the sorted edge list parse_graph builds, :34-91
//   gvcd_graph_mutate            the reduction_graph mutators the reductions call between two
//                                predicts (include/reduction_graph.hpp:248-587), relable_graph :175
//   gvcd_predict                 x(u,0) = ...; m.predict(x, out, g)     :189-192
// so a timing of gvcd_predict is a timing of gnn::model::predict as the solver sees it: CSR
// extraction from the reduction_graph, upload, forward, scores back into the host matrix.
#include "gnn_inference.hpp"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <numeric>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "gvc.h"
#include "gvc_host_ctx.hpp"

namespace {
struct dropin_graph {
    reduction_graph<gnn::Tn, gnn::Tw> g;
    matrix in, out;
};
}  // namespace

extern "C" {

void *gvcd_model_create(const char *text) {
    auto *m = new gnn::model();
    std::istringstream is{std::string(text)};
    is >> *m;
    return m;
}

void gvcd_model_destroy(void *m) { delete static_cast<gnn::model *>(m); }

void gvcd_model_set_weight_scale(void *m, float ws) { static_cast<gnn::model *>(m)->set_weight_scale(ws); }

// eu[i] < ev[i], sorted, unique: what parse_graph hands to the constructor
void *gvcd_graph_create(uint32_t n, uint64_t n_edges, const uint32_t *eu, const uint32_t *ev, const uint32_t *w) {
    std::vector<gnn::Tw> weights(w, w + n);
    std::vector<std::pair<gnn::Tn, gnn::Tn>> edges(n_edges);
    for (uint64_t i = 0; i < n_edges; ++i) edges[i] = {eu[i], ev[i]};
    return new dropin_graph{reduction_graph<gnn::Tn, gnn::Tw>(weights, edges), matrix(), matrix()};
}

void gvcd_graph_destroy(void *g) { delete static_cast<dropin_graph *>(g); }

uint32_t gvcd_graph_size(void *g) { return static_cast<dropin_graph *>(g)->g.size(); }

// op: 0 remove_node(u)  1 remove_neighborhood(u)  2 fold_neighborhood(u)  3 fold_twin(u, v)
//     4 fold_isolated(u)  5 relable_graph()  6 actions_pop()
// Returns -1 without calling when the mutator's own precondition does not hold.
int gvcd_graph_mutate(void *gh, int op, uint32_t u, uint32_t v) {
    auto &g = static_cast<dropin_graph *>(gh)->g;
    auto live = [&](uint32_t a) { return a < g.size() && g.is_active(a); };
    switch (op) {
    case 0: if (!live(u)) return -1; g.remove_node(u); return 0;
    case 1: if (!live(u)) return -1; g.remove_neighborhood(u); return 0;
    case 2: if (!live(u) || g.NW(u) <= g.W(u) || !g.has_independent_neighbors(u)) return -1; g.fold_neighborhood(u); return 0;
    case 3: if (!live(u) || !live(v) || !g.is_twin(u, v)) return -1; g.fold_twin(u, v); return 0;
    case 4: if (!live(u) || !g.is_isolated(u)) return -1; g.fold_isolated(u); return 0;
    case 5: g.relable_graph(); return 0;
    case 6: if (g.get_timestamp() == 0) return -1; g.actions_pop(); return 0;
    default: return -2;
    }
}

// out(u,0) for u < g.size(); *seconds = wall time of predict() alone.  Returns 0, or -1 if predict
// left `out` with another shape than N x 1.
int gvcd_predict(void *mh, void *gh, const float *x, float *scores, double *seconds) {
    auto *m = static_cast<gnn::model *>(mh);
    auto *d = static_cast<dropin_graph *>(gh);
    const uint32_t n = d->g.size();
    d->in.resize(n, 1);
    for (uint32_t u = 0; u < n; ++u) d->in(u, 0) = x[u];
    const auto t0 = std::chrono::steady_clock::now();
    m->predict(d->in, d->out, d->g);
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (d->out.get_height() != n || (n && d->out.get_width() != 1)) return -1;
    for (uint32_t u = 0; u < n; ++u) scores[u] = d->out(u, 0);
    return 0;
}

// predict, then the vertex order src/GNN_VC.cpp:186-206 derives from it: nodes = 0..N-1 sorted with the
// driver's tolerance comparator.  The comparator's inputs -- min(out, 1 - out) and out > 0.5 -- are not
// recomputed on the host: they come off the device, where the stage-2 kernel wrote them beside the scores
// (gvc_last_keys).  Same std::sort, same initial order, same decisions => same
// permutation as the driver's own sort (the comparator is not a strict weak order, so this only holds
// because every decision input is bit-identical).
int gvcd_predict_order(void *mh, void *gh, const float *x, float *scores, uint32_t *nodes, double *seconds) {
    auto *m = static_cast<gnn::model *>(mh);
    auto *d = static_cast<dropin_graph *>(gh);
    const uint32_t n = d->g.size();
    d->in.resize(n, 1);
    for (uint32_t u = 0; u < n; ++u) d->in(u, 0) = x[u];
    const auto t0 = std::chrono::steady_clock::now();
    m->predict(d->in, d->out, d->g);
    std::vector<float> key(n);
    std::vector<unsigned char> above(n);
    if (n) {
        const int rc = gvc_last_keys(gvc_host::context(), key.data(), above.data());
        if (rc != 0) gvc_host::die("gvc_last_keys", rc);
    }
    std::vector<uint32_t> w(n), deg(n);
    for (uint32_t u = 0; u < n; ++u) { w[u] = d->g.W(u); deg[u] = (uint32_t)(d->g.end(u) - d->g.begin(u)); }
    std::iota(nodes, nodes + n, 0u);
    const float eps = 0.0001f;
    std::sort(nodes, nodes + n, [&](uint32_t a, uint32_t b) {
        const float ka = key[a], kb = key[b];
        if (ka < (kb + eps) && ka > (kb - eps)) {                  // within tolerance: side, then weight, then degree
            const bool a_up = above[a], b_up = above[b];
            const bool a_down = !a_up && ka != 0.5f, b_down = !b_up && kb != 0.5f;   // out < 0.5 (key == 0.5 <=> out == 0.5)
            if (a_down && b_up) return true;
            if (a_up && b_up) return w[a] < w[b] || (w[a] == w[b] && deg[a] > deg[b]);
            if (a_down && b_down) return w[a] > w[b] || (w[a] == w[b] && deg[a] < deg[b]);
            return false;
        }
        return ka < kb;
    });
    if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (uint32_t u = 0; u < n; ++u) scores[u] = d->out(u, 0);
    return 0;
}

}  // extern "C"
