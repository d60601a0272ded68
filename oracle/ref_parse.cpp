// ref_parse.cpp -- TEST INFRASTRUCTURE ONLY.
//
// The reference's own METIS reader, parse_graph (src/GNN_VC.cpp:34-91), reachable from the tests: the
// driver's translation unit is included where it lies (its main renamed), so that the function that
// runs is the reference's, unmodified.  Built into oracle/_ref/libgnnref.so (oracle/Makefile);
// checker for gnn-mwvc_b200/host/gvc_metis.cpp.
#define main gnn_vc_reference_main
#include GVC_REF_SRC
#undef main

#include <cstring>

extern "C" {
// sizes first (nulls), then the arrays: weights[n], eu/ev[E] as the reduction_graph constructor receives them
int ref_parse_graph(const char *path, uint64_t *n, uint64_t *e, uint32_t *weights, uint32_t *eu, uint32_t *ev) {
    static test_graph *last = nullptr;
    static std::string last_path;
    if (!last || last_path != path) {
        delete last;
        last = new test_graph(parse_graph(path));
        last_path = path;
    }
    if (n) *n = last->weights.size();
    if (e) *e = last->edges.size();
    if (weights) std::memcpy(weights, last->weights.data(), last->weights.size() * sizeof(uint32_t));
    if (eu) for (size_t i = 0; i < last->edges.size(); ++i) { eu[i] = last->edges[i].first; ev[i] = last->edges[i].second; }
    return 0;
}
}
