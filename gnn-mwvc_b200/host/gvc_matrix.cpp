// gvc_matrix.cpp -- drop-in translation unit for the reference's src/matrix.cpp.
//
// Compiled against the reference's own, untouched include/matrix.hpp (class
// layout and every signature come from there: include/matrix.hpp:6-49), so the
// reference driver src/GNN_VC.cpp links against it unchanged.  The container is
// plain host memory exactly as in the reference (row-major std::vector<float>
// plus the "selected row" cursor used by operator[] / begin() / end(),
// reference src/matrix.cpp:8-85).  The one compute routine, dot()
// (src/matrix.cpp:106-122, a cblas_sgemm call there), runs on the GPU through
// libgvc; OpenBLAS is not linked.
#include "matrix.hpp"

#include <cstdio>
#include <cstdlib>

#include "gvc.h"
#include "gvc_host_ctx.hpp"

// ---- construction / shape (reference src/matrix.cpp:8-21) -------------------------
matrix::matrix(size_t rows, size_t cols) : m(rows), n(cols), data(rows * cols) {}

void matrix::resize(size_t rows, size_t cols) {
    const bool same_shape = (rows == m) && (cols == n);
    if (same_shape) return;            // contents are kept when nothing changes (:11-12)
    m = rows;
    n = cols;
    data.resize(rows * cols);          // std::vector semantics: old prefix kept, growth zero-filled
    selected_row.reset();
}

size_t matrix::get_height() const { return m; }
size_t matrix::get_width() const { return n; }

// ---- row cursor (reference src/matrix.cpp:23-43) -----------------------------------
matrix &matrix::raw() { selected_row.reset(); return *this; }
const matrix &matrix::raw() const { selected_row.reset(); return *this; }
matrix &matrix::operator[](size_t i) { selected_row = i; return *this; }
const matrix &matrix::operator[](size_t i) const { selected_row = i; return *this; }

// ---- element access (reference src/matrix.cpp:45-53) --------------------------------
float &matrix::operator()(size_t i, size_t j) { return data[i * n + j]; }
const float &matrix::operator()(size_t i, size_t j) const { return data[i * n + j]; }

// ---- iteration: whole matrix, or the selected row when one is set (:55-85) -----------
std::vector<float>::iterator matrix::begin() {
    return data.begin() + (selected_row ? *selected_row * n : 0);
}
std::vector<float>::iterator matrix::end() {
    return selected_row ? data.begin() + (*selected_row + 1) * n : data.end();
}
std::vector<float>::const_iterator matrix::begin() const {
    return data.cbegin() + (selected_row ? *selected_row * n : 0);
}
std::vector<float>::const_iterator matrix::end() const {
    return selected_row ? data.cbegin() + (*selected_row + 1) * n : data.cend();
}
std::vector<float>::iterator matrix::begin(size_t i) { return data.begin() + i * n; }
std::vector<float>::iterator matrix::end(size_t i) { return data.begin() + (i + 1) * n; }
std::vector<float>::const_iterator matrix::begin(size_t i) const { return data.cbegin() + i * n; }
std::vector<float>::const_iterator matrix::end(size_t i) const { return data.cbegin() + (i + 1) * n; }

// ---- text I/O (reference src/matrix.cpp:87-104): "h w" then rows, values followed by a blank
std::ostream &operator<<(std::ostream &os, const matrix &mat) {
    os << mat.get_height() << " " << mat.get_width() << std::endl;
    for (size_t i = 0; i < mat.get_height(); ++i) {
        for (auto it = mat.begin(i); it != mat.end(i); ++it) os << *it << " ";
        os << std::endl;
    }
    return os;
}

std::istream &operator>>(std::istream &is, matrix &mat) {
    size_t rows = 0, cols = 0;
    is >> rows >> cols;
    mat.resize(rows, cols);
    const size_t count = rows * cols;
    auto it = mat.raw().begin();
    for (size_t i = 0; i < count; ++i, ++it) is >> *it;
    return is;
}

// ---- dot (reference src/matrix.cpp:106-122): C = op(A) * op(B) + beta * C on the GPU
void dot(const matrix &A, const matrix &B, matrix &C, bool at, bool bt, float beta) {
    const size_t rows = at ? A.get_width() : A.get_height();
    const size_t cols = bt ? B.get_height() : B.get_width();
    const size_t inner = at ? A.get_height() : A.get_width();
    C.resize(rows, cols);
    if (rows == 0 || cols == 0) return;
    gvc_ctx *ctx = gvc_host::context();
    const int rc = gvc_sgemm_host(ctx, at ? 1 : 0, bt ? 1 : 0, rows, cols, inner, A.data.data(), A.get_width(),
                                  B.data.data(), B.get_width(), beta, C.data.data(), C.get_width());
    if (rc != 0) gvc_host::die("dot", rc);
}
