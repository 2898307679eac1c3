// capi.cu -- the extern "C" surface declared in include/qpb200.h.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <new>

#include "host_common.h"
#include "sparse_solver.h"

extern "C" {

int qpb200_version(void) { return QPB200_VERSION; }

const char *qpb200_last_error(void) { return qpb::last_error().c_str(); }

int qpb200_device_count(void) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) return qpb::fail(QPB200_ERR_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    int ok = 0;
    for (int d = 0; d < count; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
    }
    return ok;
}

void qpb200_default_settings(qpb200_settings *s) {
    if (!s) return;
    std::memset(s, 0, sizeof(*s));
    s->max_iter = 5000;          // SolveQuadraticProgram.jl:15
    s->eps_abs = 1e-6;
    s->eps_rel = 1e-6;
    s->rho = 1.0;                // :16
    s->sigma = 1e-6;
    s->alpha = 1.6;
    s->delta = 1e-6;
    s->adaptive_rho = 0;
    s->lin_solver = QPB200_LINSOLVE_PCG;
    s->rho_factor = 5.0;         // :17
    s->check_every = 25;
    s->polish_iter = 10;
    s->minres_eps = 1e-6;
    s->minres_iter = 500;
    s->pcg_eps = 1e-6;           // LinearSystemSolvers.jl:125
    s->pcg_max_iter = 1000;
    s->pcg_rel_eps = -1.0;       // IterativeSolvers default sqrt(eps)
    s->precond = QPB200_PRECOND_JACOBI;
    s->device = -1;
    s->spmv_loader = 0;
}

int qpb200_create(qpb200_handle **out, int64_t n, int64_t m, const int64_t *P_colptr, const int64_t *P_rowval,
                  const double *P_nzval, const int64_t *A_colptr, const int64_t *A_rowval, const double *A_nzval,
                  const double *q, const double *l, const double *u, const qpb200_settings *settings, int32_t index_base) {
    if (!out) return qpb::fail(QPB200_ERR_ARG, "qpb200_create: out is NULL");
    *out = nullptr;
    qpb200_settings s;
    if (settings) s = *settings;
    else qpb200_default_settings(&s);
    qpb200_handle *h = new (std::nothrow) qpb200_handle();
    if (!h) return qpb::fail(QPB200_ERR_ARG, "out of host memory");
    const int rc = h->solver.init(n, m, P_colptr, P_rowval, P_nzval, A_colptr, A_rowval, A_nzval, q, l, u, s, index_base);
    if (rc != QPB200_OK) {
        delete h;
        return rc;
    }
    *out = h;
    return QPB200_OK;
}

int qpb200_solve(qpb200_handle *h, double *x_inout, double *z_out, double *y_out, qpb200_info *info) {
    if (!h) return qpb::fail(QPB200_ERR_ARG, "qpb200_solve: handle is NULL");
    if (h->dist) return qpb::fail(QPB200_ERR_ARG, "qpb200_solve: distributed handle, call qpb200_dist_solve on every rank");
    return h->solver.solve(x_inout, z_out, y_out, info);
}

int qpb200_update_vectors(qpb200_handle *h, const double *q, const double *l, const double *u) {
    if (!h) return qpb::fail(QPB200_ERR_ARG, "qpb200_update_vectors: handle is NULL");
    if (h->dist) return qpb::fail(QPB200_ERR_ARG, "qpb200_update_vectors: not implemented for distributed handles");
    return h->solver.update_vectors(q, l, u);
}

int qpb200_update_settings(qpb200_handle *h, const qpb200_settings *settings) {
    if (!h || !settings) return qpb::fail(QPB200_ERR_ARG, "qpb200_update_settings: NULL argument");
    if (h->dist) return qpb::fail(QPB200_ERR_ARG, "qpb200_update_settings: not implemented for distributed handles");
    return h->solver.settings_to_dev(*settings);
}

int qpb200_set_rho_scale(qpb200_handle *h, const double *rho_scale) {
    if (!h) return qpb::fail(QPB200_ERR_ARG, "qpb200_set_rho_scale: handle is NULL");
    if (h->dist) return qpb::fail(QPB200_ERR_ARG, "qpb200_set_rho_scale: not implemented for distributed handles");
    return h->solver.set_rho_scale(rho_scale);
}

void qpb200_proxqp_default_settings(qpb200_settings *s) {
    if (!s) return;
    qpb200_default_settings(s);
    s->max_iter = 2000;          // ProxQP.jl:118
    s->eps_abs = 1e-7;
    s->eps_rel = 1e-6;
    s->check_every = 50;
    s->rho = 1e2;
    s->sigma = 1e-2;
    s->adaptive_rho = 1;
    s->rho_factor = 10.0;        // tau
    s->lin_solver = QPB200_LINSOLVE_CHOLESKY;
}

int qpb200_proxqp_solve(qpb200_handle *h, int64_t m_eq, const qpb200_settings *settings, double *x, double *y, double *z,
                        double *s, int32_t init_slack, qpb200_proxqp_report *report) {
    if (!h) return qpb::fail(QPB200_ERR_ARG, "qpb200_proxqp_solve: handle is NULL");
    if (h->dist) return qpb::fail(QPB200_ERR_ARG, "qpb200_proxqp_solve: not implemented for distributed handles");
    qpb200_settings ps;
    if (settings) ps = *settings;
    else qpb200_proxqp_default_settings(&ps);
    return h->solver.solve_proxqp(m_eq, ps, x, y, z, s, init_slack != 0, report);
}

void qpb200_destroy(qpb200_handle *h) {
    if (!h) return;
    const auto t0 = std::chrono::steady_clock::now();
    delete h;
    if (getenv("QPB200_TIMING"))
        fprintf(stderr, "[qpb200_destroy] %.1f ms\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
}

int qpb200_apply(qpb200_handle *h, int32_t which, const double *x, double *y) {
    if (!h) return qpb::fail(QPB200_ERR_ARG, "qpb200_apply: handle is NULL");
    return h->solver.apply(which, x, y);
}

int qpb200_time_apply(qpb200_handle *h, int32_t which, int32_t reps, int32_t flush_l2, double *ms_out) {
    if (!h) return qpb::fail(QPB200_ERR_ARG, "qpb200_time_apply: handle is NULL");
    return h->solver.time_apply(which, reps, flush_l2, ms_out);
}

int64_t qpb200_apply_bytes(qpb200_handle *h, int32_t which) {
    if (!h) return 0;
    return which == 100 ? h->solver.solve_bytes() : h->solver.spmv_bytes(which);
}

int64_t qpb200_debug_tile_plan(int32_t rows, const int32_t *rowptr, int32_t grid, int32_t *tiles_out, int64_t tiles_cap,
                               int32_t *cta_begin_out, int32_t *lpr_out) {
    if (rows < 0 || !rowptr || grid <= 0) return qpb::fail(QPB200_ERR_ARG, "qpb200_debug_tile_plan: bad argument");
    qpb::HostCsr M;
    M.rows = rows;
    M.cols = 0;
    M.ptr.assign(rowptr, rowptr + rows + 1);
    qpb::HostTiles T;
    qpb::build_tiles(M, qpb::kTileFill, qpb::kTileRows, T);
    qpb::assign_tiles(T, grid);
    const int64_t nt = (int64_t)T.tiles.size();
    if (tiles_out && nt <= tiles_cap)
        for (int64_t i = 0; i < nt; ++i) {
            tiles_out[4 * i + 0] = T.tiles[(size_t)i].x;
            tiles_out[4 * i + 1] = T.tiles[(size_t)i].y;
            tiles_out[4 * i + 2] = T.tiles[(size_t)i].z;
            tiles_out[4 * i + 3] = T.tiles[(size_t)i].w;
        }
    if (cta_begin_out)
        for (int b = 0; b <= grid; ++b) cta_begin_out[b] = T.cta_begin[(size_t)b];
    if (lpr_out) *lpr_out = T.lpr;
    return nt;
}

int32_t qpb200_debug_tile_nnz(void) { return qpb::kTileFill; }

int qpb200_debug_assemble_h(int64_t n, int64_t m, const int64_t *Pp, const int64_t *Pi, const double *Pv, const int64_t *Ap,
                            const int64_t *Ai, const double *Av, int32_t base, int32_t *rowptr_out, int32_t *rowmid_out,
                            int32_t *col_out, double *val_out, double *diagP_out, double *colsqA_out) {
    if (n <= 0 || m < 0 || (base != 0 && base != 1)) return qpb::fail(QPB200_ERR_ARG, "qpb200_debug_assemble_h: bad argument");
    int rc = qpb::validate_csc("P", n, n, Pp, Pi, Pv, base);
    if (rc) return rc;
    if ((rc = qpb::validate_csc("A", m, n, Ap, Ai, Av, base))) return rc;
    qpb::HostCsr H;
    std::vector<double> dP, dAA;
    qpb::assemble_h_direct(n, m, Pp, Pi, Pv, Ap, Ai, Av, base, H, dP, dAA);
    if (rowptr_out) std::copy(H.ptr.begin(), H.ptr.end(), rowptr_out);
    if (rowmid_out) std::copy(H.mid.begin(), H.mid.end(), rowmid_out);
    if (col_out) std::copy(H.idx.p, H.idx.p + H.nnz(), col_out);
    if (val_out) std::copy(H.val.p, H.val.p + H.nnz(), val_out);
    if (diagP_out) std::copy(dP.begin(), dP.end(), diagP_out);
    if (colsqA_out) std::copy(dAA.begin(), dAA.end(), colsqA_out);
    return QPB200_OK;
}

int qpb200_debug_equilibrate(int64_t n, int64_t m, const int64_t *Pp, const int64_t *Pi, const double *Pv, const int64_t *Ap,
                             const int64_t *Ai, const double *Av, const double *q, int32_t iters, int32_t base, double *D_out,
                             double *E_out, double *c_out, double *q_out, double *Pv_out, double *Av_out) {
    if (n <= 0 || m < 0 || !q || iters < 0 || (base != 0 && base != 1))
        return qpb::fail(QPB200_ERR_ARG, "qpb200_debug_equilibrate: bad argument");
    int rc = qpb::validate_csc("P", n, n, Pp, Pi, Pv, base);
    if (rc) return rc;
    if ((rc = qpb::validate_csc("A", m, n, Ap, Ai, Av, base))) return rc;
    qpb::HostCsr P, A, At;
    qpb::csc_to_csr((int)n, (int)n, Pp, Pi, Pv, base, P);
    qpb::csc_to_csr((int)m, (int)n, Ap, Ai, Av, base, A);
    qpb::csc_as_csr_of_transpose((int)m, (int)n, Ap, Ai, Av, base, At);
    std::vector<double> qs(q, q + n);
    qpb::RuizScaling sc;
    qpb::ruiz_equilibrate(P, A, At, qs, iters, sc);
    if (D_out) std::copy(sc.D.begin(), sc.D.end(), D_out);
    if (E_out) std::copy(sc.E.begin(), sc.E.end(), E_out);
    if (c_out) *c_out = sc.c;
    if (q_out) std::copy(qs.begin(), qs.end(), q_out);
    // CSR(A') rows are A's columns in CSC order: the scaled values map back 1:1; P by symmetry of the pattern
    // is returned through a CSR -> CSC pass of its own (a second transpose)
    if (Av_out) std::copy(At.val.p, At.val.p + At.nnz(), Av_out);
    if (Pv_out) {
        std::vector<int64_t> ptr(P.ptr.begin(), P.ptr.end()), idx((size_t)P.nnz());
        for (int64_t k = 0; k < P.nnz(); ++k) idx[(size_t)k] = P.idx.p[k];
        qpb::HostCsr Pt;   // CSR of P' from "CSC of P'" = CSR(P): rows of Pt are columns of P, i.e. CSC order of P
        qpb::csc_to_csr((int)n, (int)n, ptr.data(), idx.data(), P.val.p, 0, Pt);
        std::copy(Pt.val.p, Pt.val.p + Pt.nnz(), Pv_out);
    }
    return QPB200_OK;
}

}  // extern "C"
