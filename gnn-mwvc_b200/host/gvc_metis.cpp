// gvc_metis.cpp -- the METIS-format reader of GNN_VC, memory-mapped and parallel.
//
// Stands in for parse_graph (reference src/GNN_VC.cpp:34-91; format README.md:45-60), which reads the
// file with one getline + stringstream per vertex -- minutes at 100 M edges.  Same result, i.e. the
// weights and the sorted, de-duplicated list of undirected edges (u < v) that parse_graph hands to the
// reduction_graph constructor, including the reference's treatment of odd input:
//   * first line "N E ..." -- anything after the two numbers is ignored (:46-48);
//   * then one line per vertex: weight, then 1-indexed neighbours; only neighbours with a larger id than
//     the vertex itself are kept (:62-64), so each undirected edge is taken from its smaller endpoint;
//   * a line is read up to its first token that is not an unsigned decimal number (stream extraction
//     fails there, :57-60); a missing or unreadable weight is 0 and ends the line;
//   * lines missing at the end of the file are empty lines (weight 0, no neighbours);
//   * the edge array is pre-sized to the header's E (:51): when the file holds FEWER kept edges, the unused
//     entries stay (0,0) and survive sort + unique as ONE self-loop on vertex 0 (:86-87) -- reproduced;
//     when it holds MORE, the reference writes out of bounds -- reported as an error here;
//   * neighbour ids beyond N are not checked by the reference (its graph constructor then indexes out of
//     bounds) -- reported as an error here.
// Also builds the CSR exactly as the reduction_graph constructor does (include/reduction_graph.hpp:103-128:
// per-vertex lists in edge-list order, NW = sum of neighbour weights), so the device can be fed without
// going through the host graph.  Pure host code; part of libgvc.so, declared in include/gvc.h.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gvc.h"

extern "C" int gvc_internal_fail(int code, const char *msg);    // sets gvc_last_error (gvc_api.cu)

struct gvc_metis {
    uint64_t n = 0, header_e = 0;
    std::vector<uint32_t> w, eu, ev;
};

namespace {

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// one unsigned decimal token at p (after blanks); false at end of line or on anything else
inline bool next_number(const char *&p, const char *end, uint64_t &v) {
    while (p < end && is_space(*p)) ++p;
    if (p >= end || *p < '0' || *p > '9') {
        if (p < end && *p == '+' && p + 1 < end && p[1] >= '0' && p[1] <= '9') ++p;   // num_get accepts a plus sign
        else return false;
    }
    uint64_t x = 0;
    while (p < end && *p >= '0' && *p <= '9') { x = x * 10 + (uint64_t)(*p - '0'); ++p; }
    v = x;
    return true;
}

struct slice_result {
    std::vector<uint32_t> eu, ev;
    uint64_t bad_id = 0;        // a neighbour id beyond N (1-indexed value), 0 = none
};

}  // namespace

extern "C" {

int gvc_metis_parse(const char *path, int n_threads, gvc_metis **out) {
    if (!path || !out) return gvc_internal_fail(GVC_ERR_ARG, "null argument");
    *out = nullptr;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return gvc_internal_fail(GVC_ERR_ARG, "cannot open the graph file");
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return gvc_internal_fail(GVC_ERR_ARG, "cannot stat the graph file"); }
    const size_t size = (size_t)st.st_size;
    const char *data = size ? static_cast<const char *>(mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0)) : "";
    close(fd);
    if (size && data == MAP_FAILED) return gvc_internal_fail(GVC_ERR_ALLOC, "mmap of the graph file failed");
    if (size) madvise(const_cast<char *>(data), size, MADV_SEQUENTIAL);
    auto unmap = [&] { if (size) munmap(const_cast<char *>(data), size); };
    const char *end = data + size;

    gvc_metis *m = new (std::nothrow) gvc_metis();
    if (!m) { unmap(); return gvc_internal_fail(GVC_ERR_ALLOC, "out of host memory"); }
    // header
    const char *p = data;
    const char *eol = static_cast<const char *>(memchr(p, '\n', (size_t)(end - p)));
    const char *hend = eol ? eol : end;
    uint64_t N = 0, E = 0;
    {
        const char *q = p;
        if (next_number(q, hend, N)) { if (!next_number(q, hend, E)) E = 0; } else N = 0;
    }
    const char *body = eol ? eol + 1 : end;
    if (N >= (1ull << 32)) { delete m; unmap(); return gvc_internal_fail(GVC_ERR_UNSUPPORTED, "more than 2^32 vertices"); }
    m->n = N;
    m->header_e = E;
    m->w.assign(N, 0);

    // slices of the body at line boundaries; lines counted per slice, then parsed in parallel
    unsigned T = n_threads > 0 ? (unsigned)n_threads : std::max(1u, std::thread::hardware_concurrency());
    const size_t body_size = (size_t)(end - body);
    if (body_size < (1u << 20)) T = 1;
    T = std::min<unsigned>(T, 64);
    std::vector<const char *> cut(T + 1);
    cut[0] = body;
    cut[T] = end;
    for (unsigned t = 1; t < T; ++t) {
        const char *c = body + body_size / T * t;
        const char *nl = static_cast<const char *>(memchr(c, '\n', (size_t)(end - c)));
        cut[t] = nl ? nl + 1 : end;
        if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
    }
    std::vector<uint64_t> lines(T, 0);
    auto count_lines = [&](unsigned t) {
        uint64_t k = 0;
        for (const char *c = cut[t]; c < cut[t + 1];) {
            const char *nl = static_cast<const char *>(memchr(c, '\n', (size_t)(cut[t + 1] - c)));
            ++k;                                   // a last line without a newline counts too
            if (!nl) break;
            c = nl + 1;
        }
        lines[t] = k;
    };
    auto run = [&](auto &&fn) {
        if (T == 1) { fn(0u); return; }
        std::vector<std::thread> th;
        for (unsigned t = 1; t < T; ++t) th.emplace_back([&fn, t] { fn(t); });
        fn(0u);
        for (auto &x : th) x.join();
    };
    run(count_lines);
    std::vector<uint64_t> first(T + 1, 0);
    for (unsigned t = 0; t < T; ++t) first[t + 1] = first[t] + lines[t];

    std::vector<slice_result> res(T);
    auto parse_slice = [&](unsigned t) {
        slice_result &r = res[t];
        uint64_t i = first[t];
        for (const char *c = cut[t]; c < cut[t + 1] && i < N; ++i) {
            const char *nl = static_cast<const char *>(memchr(c, '\n', (size_t)(cut[t + 1] - c)));
            const char *le = nl ? nl : cut[t + 1];
            const char *q = c;
            uint64_t v = 0;
            if (next_number(q, le, v)) {
                m->w[i] = (uint32_t)v;
                while (next_number(q, le, v)) {
                    if (v == 0) { r.bad_id = r.bad_id ? r.bad_id : UINT64_MAX; continue; }   // "0" wraps to a huge id in the reference
                    const uint64_t e = v - 1;
                    if (e <= i) continue;
                    if (e >= N) { if (!r.bad_id) r.bad_id = v; continue; }
                    r.eu.push_back((uint32_t)i);
                    r.ev.push_back((uint32_t)e);
                }
            }
            if (!nl) break;
            c = nl + 1;
        }
    };
    run(parse_slice);
    unmap();
    uint64_t kept = 0;
    for (auto &r : res) {
        if (r.bad_id) { delete m; return gvc_internal_fail(GVC_ERR_ARG, "a neighbour id is 0 or larger than the vertex count"); }
        kept += r.eu.size();
    }
    if (kept > E) { delete m; return gvc_internal_fail(GVC_ERR_ARG, "the file holds more edges than its header says (the reference writes out of bounds here)"); }
    m->eu.reserve(kept + 1);
    m->ev.reserve(kept + 1);
    const bool pad = kept < E;                       // unused pre-sized entries are (0,0): one of them survives unique
    if (pad) { m->eu.push_back(0); m->ev.push_back(0); }
    for (auto &r : res) {
        m->eu.insert(m->eu.end(), r.eu.begin(), r.eu.end());
        m->ev.insert(m->ev.end(), r.ev.begin(), r.ev.end());
        std::vector<uint32_t>().swap(r.eu);
        std::vector<uint32_t>().swap(r.ev);
    }
    // sort + unique (:86-87); files written by sane tools are sorted already
    const size_t k = m->eu.size();
    bool sorted = true;
    for (size_t i = 1; i < k && sorted; ++i)
        sorted = m->eu[i - 1] < m->eu[i] || (m->eu[i - 1] == m->eu[i] && m->ev[i - 1] < m->ev[i]);
    if (!sorted) {
        std::vector<uint64_t> key(k);
        for (size_t i = 0; i < k; ++i) key[i] = ((uint64_t)m->eu[i] << 32) | m->ev[i];
        std::sort(key.begin(), key.end());
        key.erase(std::unique(key.begin(), key.end()), key.end());
        m->eu.resize(key.size());
        m->ev.resize(key.size());
        for (size_t i = 0; i < key.size(); ++i) { m->eu[i] = (uint32_t)(key[i] >> 32); m->ev[i] = (uint32_t)key[i]; }
    }
    *out = m;
    return 0;
}

void gvc_metis_free(gvc_metis *m) { delete m; }
uint64_t gvc_metis_vertices(const gvc_metis *m) { return m ? m->n : 0; }
uint64_t gvc_metis_edges(const gvc_metis *m) { return m ? m->eu.size() : 0; }
const uint32_t *gvc_metis_weights(const gvc_metis *m) { return m && !m->w.empty() ? m->w.data() : nullptr; }
const uint32_t *gvc_metis_edge_u(const gvc_metis *m) { return m && !m->eu.empty() ? m->eu.data() : nullptr; }
const uint32_t *gvc_metis_edge_v(const gvc_metis *m) { return m && !m->ev.empty() ? m->ev.data() : nullptr; }

// The adjacency the reduction_graph constructor builds from (weights, edges)
// (include/reduction_graph.hpp:103-128): degree counts, prefix sums, then every edge (u, v) in list order
// appended to u's and to v's list; NW(u) = sum of the neighbours' weights.  row_ptr[n + 1], col[2 E], nw[n].
int gvc_metis_csr(const gvc_metis *m, uint64_t *row_ptr, uint32_t *col, uint32_t *nw) {
    if (!m || !row_ptr) return gvc_internal_fail(GVC_ERR_ARG, "null argument");
    const uint64_t n = m->n, e = m->eu.size();
    for (uint64_t u = 0; u <= n; ++u) row_ptr[u] = 0;
    for (uint64_t i = 0; i < e; ++i) { row_ptr[m->eu[i] + 1]++; row_ptr[m->ev[i] + 1]++; }
    for (uint64_t u = 0; u < n; ++u) row_ptr[u + 1] += row_ptr[u];
    if (!col && !nw) return 0;
    std::vector<uint64_t> at(row_ptr, row_ptr + n);
    if (nw) for (uint64_t u = 0; u < n; ++u) nw[u] = 0;
    for (uint64_t i = 0; i < e; ++i) {
        const uint32_t u = m->eu[i], v = m->ev[i];
        if (col) { col[at[u]++] = v; col[at[v]++] = u; }
        if (nw) { nw[u] += m->w[v]; nw[v] += m->w[u]; }
    }
    return 0;
}

}  // extern "C"
