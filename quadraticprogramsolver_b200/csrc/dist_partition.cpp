// dist_partition.cpp -- the row / column partition of ONE large sparse QP over R ranks, behind the C ABI
// (qpb200_dist_create_full): SURVEY.md 8(e) "row partition done inside".  Host only, no device code.
//
// Rank r owns a contiguous, nnz-balanced block of rows I_r of A (and of l, u, z, y) and a contiguous, nnz-balanced
// block of columns J_r of P, so that with H_r = [P[:, J_r]  A_r'] the operator of LinearSystemSolvers.jl:152-157 is
//     K u = sum_r ( P[:, J_r] u[J_r] + rho A_r' (A_r u) ) + sigma u .
// The boundaries are the ones quadraticprogramsolver_b200/partition.py::plan computes (tests/test_host.py compares
// them), so pre-sliced callers and the in-library path agree on who owns what.
#include <algorithm>
#include <cstring>

#include "host_common.h"

namespace qpb {

// Boundaries b[0..parts] of contiguous blocks of (nearly) equal total weight, weight = counts[i] + 1 (so that empty
// rows / columns are spread too): first index whose running sum reaches k/parts of the total.
void balanced_blocks(const int64_t *counts, int64_t len, int parts, std::vector<int64_t> &b) {
    std::vector<double> c((size_t)len + 1);
    c[0] = 0.0;
    for (int64_t i = 0; i < len; ++i) c[(size_t)i + 1] = c[(size_t)i] + ((double)counts[i] + 1.0);
    b.assign((size_t)parts + 1, 0);
    b[(size_t)parts] = len;
    for (int k = 1; k < parts; ++k) {
        const double target = c[(size_t)len] * (double)k / (double)parts;
        b[(size_t)k] = std::lower_bound(c.begin(), c.end(), target) - c.begin();
    }
    for (int k = 1; k <= parts; ++k) b[(size_t)k] = std::max(b[(size_t)k], b[(size_t)k - 1]);
}

// (row_bounds[R+1], col_bounds[R+1]) from the CSC arrays of A (m x n) and the column pointers of P
void dist_plan(int64_t n, int64_t m, const int64_t *Pp, const int64_t *Ap, const int64_t *Ai, int64_t base, int nranks,
               std::vector<int64_t> &rows, std::vector<int64_t> &cols) {
    std::vector<int64_t> cnt((size_t)std::max<int64_t>(m, 1), 0);
    const int64_t nnzA = Ap[n] - base;
    parallel_chunks(nnzA, [&](int, int64_t b, int64_t e) {
        for (int64_t k = b; k < e; ++k) __atomic_fetch_add(&cnt[(size_t)(Ai[k] - base)], (int64_t)1, __ATOMIC_RELAXED);
    }, 1 << 18);
    balanced_blocks(cnt.data(), m, nranks, rows);
    std::vector<int64_t> pc((size_t)n);
    for (int64_t j = 0; j < n; ++j) pc[(size_t)j] = Pp[j + 1] - Pp[j];
    balanced_blocks(pc.data(), n, nranks, cols);
}

// The slice of rank `rank`: P's columns [j0, j1) need no copy (column pointers clamped to the slice, the index / value
// arrays are the caller's, offset by p_off); A's rows [i0, i1) are filtered out of every column (two passes, all host
// threads; rows need not be sorted inside a column).  All arrays keep the caller's index base.
void dist_slice(int64_t n, int64_t m, const int64_t *Pp, const int64_t *Ap, const int64_t *Ai, const double *Av, int64_t base,
                int rank, const std::vector<int64_t> &rows, const std::vector<int64_t> &cols, DistSlice &out) {
    (void)m;
    out.i0 = rows[(size_t)rank]; out.i1 = rows[(size_t)rank + 1];
    out.j0 = cols[(size_t)rank]; out.j1 = cols[(size_t)rank + 1];
    const int64_t lo = Pp[out.j0] - base, hi = Pp[out.j1] - base;
    out.p_off = lo;
    out.Pcolptr.resize((size_t)n + 1);
    for (int64_t j = 0; j <= n; ++j) out.Pcolptr[(size_t)j] = base + std::min(std::max(Pp[j] - base, lo), hi) - lo;
    out.Acolptr.assign((size_t)n + 1, 0);
    const int64_t i0 = out.i0 + base, i1 = out.i1 + base;
    parallel_chunks(n, [&](int, int64_t b, int64_t e) {
        for (int64_t j = b; j < e; ++j) {
            int64_t c = 0;
            for (int64_t k = Ap[j] - base; k < Ap[j + 1] - base; ++k) c += (Ai[k] >= i0 && Ai[k] < i1);
            out.Acolptr[(size_t)j + 1] = c;
        }
    }, 1 << 12);
    out.Acolptr[0] = base;
    for (int64_t j = 0; j < n; ++j) out.Acolptr[(size_t)j + 1] += out.Acolptr[(size_t)j];
    const int64_t nnz = out.Acolptr[(size_t)n] - base;
    out.Arowval.resize((size_t)std::max<int64_t>(nnz, 1));
    out.Anzval.resize((size_t)std::max<int64_t>(nnz, 1));
    parallel_chunks(n, [&](int, int64_t b, int64_t e) {
        for (int64_t j = b; j < e; ++j) {
            int64_t pos = out.Acolptr[(size_t)j] - base;
            for (int64_t k = Ap[j] - base; k < Ap[j + 1] - base; ++k)
                if (Ai[k] >= i0 && Ai[k] < i1) {
                    out.Arowval[(size_t)pos] = Ai[k] - out.i0;      // local row, same base
                    out.Anzval[(size_t)pos] = Av[k];
                    ++pos;
                }
        }
    }, 1 << 12);
}

}  // namespace qpb
