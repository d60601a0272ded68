// ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C entry points around the UNMODIFIED reference forward pass.  oracle/Makefile
// compiles this file together with /root/reference/src/matrix.cpp and
// /root/reference/src/gnn_inference.cpp (where they lie, never copied) against
// /root/reference/include and the wheel-bundled OpenBLAS, into
// oracle/_ref/libgnnref.so.  It is used to pin the C restatement
// (oracle/gnn_oracle.c), to generate tests/golden/, and as bench.py's
// "reference" CPU arm.  Nothing in the product links it.
//
// Calls into the reference:
//   gnn::operator>>            src/gnn_inference.cpp:120-139
//   model::set_weight_scale    src/gnn_inference.cpp:83-90
//   model::predict             src/gnn_inference.cpp:67-81
//   graph_layer::forward       src/gnn_inference.cpp:27-42
//   linear_layer::forward      src/gnn_inference.cpp:20-25
//   reduction_graph ctor       include/reduction_graph.hpp:103-128
#include "gnn_inference.hpp"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#ifdef GVC_HARNESS_DROPIN
// The same harness over the DROP-IN host units (gnn-mwvc_b200/host/*.cpp + libgvc) instead of the
// reference's src/*.cpp: `make dropin_harness`.  Lets the tests call the replacement through the
// reference's own C++ interface (parse, print, constructors on the CPU; forward on a GPU).
static void openblas_set_num_threads(int) {}
static char *openblas_get_config(void) {
    static char s[] = "drop-in host units over libgvc (no BLAS)";
    return s;
}
#else
extern "C" {
void openblas_set_num_threads(int n);
char *openblas_get_config(void);
}
#endif

namespace {
struct ref_state {
    gnn::model m;
    std::vector<gnn::component> parsed; // not accessible in model (private) -> layer probes use own copies
};

reduction_graph<uint32_t, uint32_t> make_graph(uint32_t n, uint64_t n_edges, const uint32_t *eu,
                                               const uint32_t *ev, const uint32_t *w) {
    std::vector<uint32_t> weights(w, w + n);
    std::vector<std::pair<uint32_t, uint32_t>> edges(n_edges);
    for (uint64_t i = 0; i < n_edges; i++) edges[i] = {eu[i], ev[i]};
    return reduction_graph<uint32_t, uint32_t>(weights, edges);
}
} // namespace

extern "C" {

const char *ref_blas_config() { return openblas_get_config(); }
void ref_blas_threads(int n) { openblas_set_num_threads(n); }

void *ref_model_create(const char *text) {
    auto *s = new ref_state();
    std::string src(text);
    std::istringstream is(src);
    is >> s->m;
    return s;
}

// A model assembled in code, as old_files' training tools do (model(name), add_layer, the layer
// structs' public members; src/gnn_inference.cpp:54-59): kinds[i] in the order of gnn::component's
// alternatives (0 linear, 1 graph, 2 ReLU, 3 sigmoid); linear layers take rows[i] x cols[i] weights
// W[i] and cols[i] bias values; graph layers take their own WEIGHT_SCALE scales[i] (the one thing
// set_weight_scale cannot express: different scales per layer).
void *ref_model_build(int n, const int *kinds, const int *rows, const int *cols, const float *const *W,
                      const float *const *bias, const float *scales) {
    auto *s = new ref_state();
    s->m = gnn::model("built");
    for (int i = 0; i < n; i++) {
        switch (kinds[i]) {
        case 0: {
            gnn::linear_layer l(rows[i], cols[i], 0);
            std::copy(W[i], W[i] + (size_t)rows[i] * cols[i], l.W.raw().begin());
            std::copy(bias[i], bias[i] + cols[i], l.bias.raw().begin());
            s->m.add_layer(l);
            break;
        }
        case 1: {
            gnn::graph_layer gl;
            gl.WEIGHT_SCALE = scales[i];
            s->m.add_layer(gl);
            break;
        }
        case 2: s->m.add_layer(gnn::ReLU()); break;
        default: s->m.add_layer(gnn::sigmoid()); break;
        }
    }
    return s;
}

void ref_model_destroy(void *h) { delete static_cast<ref_state *>(h); }

void ref_model_set_weight_scale(void *h, float ws) { static_cast<ref_state *>(h)->m.set_weight_scale(ws); }

// Serialise through the reference's operator<< (gnn_inference.cpp:92-118).
// Returns the number of bytes needed (excluding NUL); copies up to cap-1.
size_t ref_model_text(void *h, char *buf, size_t cap) {
    std::ostringstream os;
    os << static_cast<ref_state *>(h)->m;
    std::string t = os.str();
    if (cap) {
        size_t k = t.size() < cap - 1 ? t.size() : cap - 1;
        std::memcpy(buf, t.data(), k);
        buf[k] = 0;
    }
    return t.size();
}

// The graph is given as the sorted, de-duplicated list of undirected edges
// (u < v) that parse_graph (src/GNN_VC.cpp:34-91) would hand to the
// reduction_graph ctor.  x: n floats, out: n floats.  If csr_* are non-null the
// CSR as seen through g.begin(u)/end(u) (and g.NW) is written back so callers
// can feed the very same adjacency order to the code under test.
int ref_predict(void *h, uint32_t n, uint64_t n_edges, const uint32_t *eu, const uint32_t *ev,
                const uint32_t *w, const float *x, float *out, uint64_t *csr_row_ptr,
                uint32_t *csr_col, uint32_t *nw_out, int reps, double *best_seconds) {
    auto *s = static_cast<ref_state *>(h);
    auto g = make_graph(n, n_edges, eu, ev, w);
    if (csr_row_ptr) {
        uint64_t p = 0;
        for (uint32_t u = 0; u < n; u++) {
            csr_row_ptr[u] = p;
            for (auto it = g.begin(u); it != g.end(u); ++it) csr_col[p++] = *it;
        }
        csr_row_ptr[n] = p;
    }
    if (nw_out)
        for (uint32_t u = 0; u < n; u++) nw_out[u] = g.NW(u);
    matrix in(n, 1), res;
    for (uint32_t u = 0; u < n; u++) in(u, 0) = x[u];
    double best = 1e300;
    if (reps < 1) reps = 1;
    for (int r = 0; r < reps; r++) {
        auto t0 = std::chrono::steady_clock::now();
        s->m.predict(in, res, g);
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (dt < best) best = dt;
    }
    if (best_seconds) *best_seconds = best;
    if (res.get_height() != n || (n && res.get_width() != 1)) return -1;
    for (uint32_t u = 0; u < n; u++) out[u] = res(u, 0);
    return 0;
}

// ---- a graph that lives across calls ------------------------------------------------------------
// GNN_VC builds its graph once (src/GNN_VC.cpp:265) and calls predict on it again and again while
// the reductions shrink it (:171-192).  bench.py's reference arm and the drop-in's end-to-end leg
// time predict() alone on such a resident graph; the tests mutate it through the reference's own
// reduction_graph members (include/reduction_graph.hpp:248-587) to get the holes, rotated lists,
// appended fold vertices and relabelled ids that predict sees inside a real run.
struct ref_graph_state {
    reduction_graph<uint32_t, uint32_t> g;
    matrix in, res;
};

void *ref_graph_create(uint32_t n, uint64_t n_edges, const uint32_t *eu, const uint32_t *ev, const uint32_t *w) {
    return new ref_graph_state{make_graph(n, n_edges, eu, ev, w), matrix(), matrix()};
}

void ref_graph_destroy(void *gh) { delete static_cast<ref_graph_state *>(gh); }

uint32_t ref_graph_size(void *gh) { return static_cast<ref_graph_state *>(gh)->g.size(); }

// op: 0 remove_node(u)  1 remove_neighborhood(u)  2 fold_neighborhood(u)  3 fold_twin(u, v)
//     4 fold_isolated(u)  5 relable_graph()  6 actions_pop() (undo the last one)
// Preconditions are the reference's (asserts are compiled out): returns -1 instead of calling when
// u/v are out of range or inactive, or when the fold's own precondition does not hold.
int ref_graph_mutate(void *gh, int op, uint32_t u, uint32_t v) {
    auto &g = static_cast<ref_graph_state *>(gh)->g;
    auto ok = [&](uint32_t a) { return a < g.size() && g.is_active(a); };
    switch (op) {
    case 0: if (!ok(u)) return -1; g.remove_node(u); return 0;
    case 1: if (!ok(u)) return -1; g.remove_neighborhood(u); return 0;
    case 2: if (!ok(u) || !g.has_independent_neighbors(u) || g.NW(u) <= g.W(u)) return -1; g.fold_neighborhood(u); return 0;
    case 3: if (!ok(u) || !ok(v) || !g.is_twin(u, v)) return -1; g.fold_twin(u, v); return 0;
    case 4: if (!ok(u) || !g.is_isolated(u)) return -1; g.fold_isolated(u); return 0;
    case 5: g.relable_graph(); return 0;
    case 6: if (g.get_timestamp() == 0) return -1; g.actions_pop(); return 0;
    default: return -2;
    }
}

// The graph as predict sees it: through size(), begin(u)/end(u), W, NW only.  Pass nulls to get the
// sizes first (returns nnz).  Inactive vertices (before a relabel) are reported with active[u] = 0.
uint64_t ref_graph_csr(void *gh, uint64_t *row_ptr, uint32_t *col, uint32_t *w, uint32_t *nw, uint8_t *active) {
    auto &g = static_cast<ref_graph_state *>(gh)->g;
    uint64_t p = 0;
    for (uint32_t u = 0; u < g.size(); u++) {
        if (row_ptr) row_ptr[u] = p;
        for (auto it = g.begin(u); it != g.end(u); ++it, ++p)
            if (col) col[p] = *it;
        if (w) w[u] = g.W(u);
        if (nw) nw[u] = g.NW(u);
        if (active) active[u] = g.is_active(u) ? 1 : 0;
    }
    if (row_ptr) row_ptr[g.size()] = p;
    return p;
}

// model::predict on the resident graph; x and out have g.size() entries.  Wall time of predict() alone
// (best of reps) in *seconds.
int ref_predict_on(void *h, void *gh, const float *x, float *out, int reps, double *seconds) {
    auto *s = static_cast<ref_state *>(h);
    auto *gs = static_cast<ref_graph_state *>(gh);
    const uint32_t n = gs->g.size();
    gs->in.resize(n, 1);
    for (uint32_t u = 0; u < n; u++) gs->in(u, 0) = x[u];
    double best = 1e300;
    if (reps < 1) reps = 1;
    for (int r = 0; r < reps; r++) {
        auto t0 = std::chrono::steady_clock::now();
        s->m.predict(gs->in, gs->res, gs->g);
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (dt < best) best = dt;
    }
    if (seconds) *seconds = best;
    if (gs->res.get_height() != n || (n && gs->res.get_width() != 1)) return -1;
    for (uint32_t u = 0; u < n; u++) out[u] = gs->res(u, 0);
    return 0;
}

// What the driver does with the scores (src/GNN_VC.cpp:186-206): nodes = 0..N-1, predict, then std::sort
// with its tolerance comparator on min(out, 1 - out), ties decided by the side of 0.5, the weight and the
// degree.  Restated here over whichever predict this harness is built on, as the checker for the
// device-side selection keys.
int ref_selection_order(void *h, void *gh, const float *x, float *out, uint32_t *nodes) {
    auto *s = static_cast<ref_state *>(h);
    auto *gs = static_cast<ref_graph_state *>(gh);
    auto &g = gs->g;
    const uint32_t n = g.size();
    gs->in.resize(n, 1);
    for (uint32_t u = 0; u < n; u++) gs->in(u, 0) = x[u];
    s->m.predict(gs->in, gs->res, g);
    const matrix &o = gs->res;
    for (uint32_t u = 0; u < n; u++) { nodes[u] = u; out[u] = o(u, 0); }
    std::sort(nodes, nodes + n, [&](uint32_t a, uint32_t b) {
        const float pa = o(a, 0), pb = o(b, 0);
        const float ka = std::min(pa, 1.0f - pa), kb = std::min(pb, 1.0f - pb), tol = 0.0001f;
        const bool close = ka < (kb + tol) && ka > (kb - tol);
        if (!close) return ka < kb;
        if (pa < 0.5 && pb > 0.5) return true;
        if (pa > 0.5 && pb > 0.5) return g.W(a) < g.W(b) || (g.W(a) == g.W(b) && g.D(a) > g.D(b));
        if (pa < 0.5 && pb < 0.5) return g.W(a) > g.W(b) || (g.W(a) == g.W(b) && g.D(a) < g.D(b));
        return false;
    });
    return 0;
}

// graph_layer::forward alone: in n x w  ->  out n x (2w+3)
int ref_graph_layer(uint32_t n, uint64_t n_edges, const uint32_t *eu, const uint32_t *ev,
                    const uint32_t *w, float scale, const float *in, int width, float *out) {
    auto g = make_graph(n, n_edges, eu, ev, w);
    gnn::graph_layer gl;
    gl.WEIGHT_SCALE = scale;
    matrix a(n, width), b;
    std::copy(in, in + (size_t)n * width, a.raw().begin());
    gl.forward(a, b, g);
    std::copy(b.raw().begin(), b.raw().end(), out);
    return (int)b.get_width();
}

// linear_layer::forward alone: in n x K, W K x Nout, bias Nout -> out n x Nout
void ref_linear_layer(size_t n, int K, int Nout, const float *in, const float *W,
                      const float *bias, float *out) {
    gnn::linear_layer l(K, Nout, 0);
    std::copy(W, W + (size_t)K * Nout, l.W.raw().begin());
    std::copy(bias, bias + Nout, l.bias.raw().begin());
    matrix a(n, K), b;
    std::copy(in, in + n * (size_t)K, a.raw().begin());
    l.forward(a, b);
    std::copy(b.raw().begin(), b.raw().end(), out);
}

void ref_relu(size_t n, const float *in, float *out) {
    matrix a(n, 1), b;
    std::copy(in, in + n, a.raw().begin());
    gnn::ReLU().forward(a, b);
    std::copy(b.raw().begin(), b.raw().end(), out);
}

void ref_sigmoid(size_t n, const float *in, float *out) {
    matrix a(n, 1), b;
    std::copy(in, in + n, a.raw().begin());
    gnn::sigmoid().forward(a, b);
    std::copy(b.raw().begin(), b.raw().end(), out);
}

// dot() itself (include/matrix.hpp:49, src/matrix.cpp:106-122): C = op(A) op(B) + beta C.
// A is stored (at ? k x m : m x k), B (bt ? n x k : k x n), C m x n; all row-major.
void ref_dot(int at, int bt, size_t m, size_t n, size_t k, const float *A, const float *B, float beta, float *Cio) {
    matrix a(at ? k : m, at ? m : k), b(bt ? n : k, bt ? k : n), c(m, n);
    std::copy(A, A + m * k, a.raw().begin());
    std::copy(B, B + k * n, b.raw().begin());
    std::copy(Cio, Cio + m * n, c.raw().begin());
    dot(a, b, c, at != 0, bt != 0, beta);
    std::copy(c.raw().begin(), c.raw().end(), Cio);
}

// linear_layer random init (gnn_inference.cpp:7-18), for API-parity tests of
// the drop-in host code: returns K*Nout weights followed by Nout bias values.
void ref_linear_init(int K, int Nout, size_t seed, float *out) {
    gnn::linear_layer l(K, Nout, seed);
    std::copy(l.W.raw().begin(), l.W.raw().end(), out);
    std::copy(l.bias.raw().begin(), l.bias.raw().end(), out + (size_t)K * Nout);
}

// class matrix (include/matrix.hpp, src/matrix.cpp:8-104) exercised through its public interface
// only: construction, element access, the stateful row selection of operator[] / raw(), the
// iterator pairs, the three kinds of resize, and the text form both ways.  Every observation is
// appended to `out` (as floats) and the text forms to `text`; the test runs this once over the
// reference's matrix.cpp and once over the drop-in's and compares the two records.
size_t ref_matrix_probe(float *out, size_t cap, char *text, size_t text_cap) {
    std::vector<float> o;
    std::ostringstream all;
    auto dump = [&](matrix &m) {
        o.push_back((float)m.get_height());
        o.push_back((float)m.get_width());
        matrix &r = m.raw();
        o.push_back((float)(r.end() - r.begin()));
        for (auto it = r.begin(); it != r.end(); ++it) o.push_back(*it);
    };
    matrix e;
    dump(e);
    matrix a(3, 4);
    dump(a);                                             // initial contents
    for (size_t i = 0; i < 3; ++i)
        for (size_t j = 0; j < 4; ++j) a(i, j) = (float)(10 * i + j) + 0.5f;
    dump(a);
    // row selection: a[i] narrows begin()/end() to row i until raw() is called
    for (size_t i = 0; i < 3; ++i) {
        matrix &row = a[i];
        o.push_back((float)(row.end() - row.begin()));
        for (auto it = row.begin(); it != row.end(); ++it) o.push_back(*it);
        o.push_back((float)row.get_height());
        o.push_back((float)row.get_width());
    }
    a[2];
    o.push_back((float)(a.end() - a.begin()));           // still selected?
    o.push_back(*a.begin());
    a.raw();
    o.push_back((float)(a.end() - a.begin()));
    for (size_t i = 0; i < 3; ++i) {
        o.push_back((float)(a.end(i) - a.begin(i)));
        o.push_back(*a.begin(i));
    }
    const matrix &ca = a;
    o.push_back(ca(1, 2));
    o.push_back((float)(ca.end() - ca.begin()));
    o.push_back((float)(ca.end(1) - ca.begin(1)));
    o.push_back((float)(ca[1].end() - ca[1].begin()));
    ca.raw();
    all << a << "|";                                      // text form of a whole matrix
    a[1];
    all << a << "|";                                      // ... and with a row selected
    a.raw();
    a.resize(3, 4);                                       // same shape: contents stay
    dump(a);
    a.resize(2, 5);
    dump(a);
    a.resize(4, 4);
    dump(a);
    a.resize(0, 7);
    dump(a);
    {
        std::istringstream is("2 3\n1 2 3\n4.5 -6e-1 7\n");
        matrix b;
        is >> b;
        dump(b);
        all << b << "|";
        std::istringstream is2("1 2 9 8 trailing");
        is2 >> b;                                         // re-read into a used matrix
        dump(b);
        all << b << "|";
    }
    std::string t = all.str();
    if (text && text_cap) {
        size_t k = t.size() < text_cap - 1 ? t.size() : text_cap - 1;
        std::memcpy(text, t.data(), k);
        text[k] = 0;
    }
    for (size_t i = 0; i < o.size() && i < cap; ++i) out[i] = o[i];
    return o.size();
}

// model built in code (model(name), add_layer, the layer constructors; src/gnn_inference.cpp:54-59)
// and printed: the programmatic half of the interface, used by the training
// code in old_files.  Returns the length of the text.
size_t ref_model_build_probe(char *text, size_t text_cap) {
    gnn::model m("built_in_code");
    m.add_layer(gnn::graph_layer());
    m.add_layer(gnn::linear_layer(5, 3, 42));
    m.add_layer(gnn::ReLU());
    m.add_layer(gnn::linear_layer(3, 1, 7));
    m.add_layer(gnn::sigmoid());
    gnn::graph_layer g2;
    g2.WEIGHT_SCALE = 33.0f;                               // not part of the text form
    m.add_layer(g2);
    m.set_weight_scale(200.0f);
    gnn::model copy = m;                                   // value semantics of the class
    std::ostringstream os;
    os << m << "|" << copy << "|" << gnn::model() << "|";
    std::string t = os.str();
    if (text && text_cap) {
        size_t k = t.size() < text_cap - 1 ? t.size() : text_cap - 1;
        std::memcpy(text, t.data(), k);
        text[k] = 0;
    }
    return t.size();
}

} // extern "C"
