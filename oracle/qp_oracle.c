/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * Compiled CPU restatement (C11 + OpenMP, built with -ffp-contract=off so that, like Julia, no
 * FMA contraction happens) of the reference's ADMM hot path, used (a) as a second, independently
 * written check of oracle/qp_oracle.py and (b) as the timed CPU baseline of bench.py
 * ("cpu_baseline.kind" = "port": Julia is not available in this environment).
 *
 * Follows, line by line:
 *   SolveQuadraticProgram!   /root/reference/SolveQuadraticProgram.jl:14-76
 *   CheckConvergence         /root/reference/SolveQuadraticProgram.jl:79-112
 *   LinOpCg! (matrix-free)   /root/reference/LinearSystemSolvers.jl:145-186
 *   IterativeSolvers.cg!     third-party, not vendored, version un-pinned; v0.9.x algorithm
 *                            (CGIterable / PCGIterable), call site LinearSystemSolvers.jl:181
 *   reduced dense solve      the direct plugins LinearSystemSolvers.jl:16-107 after eliminating
 *                            nu: (P + sigma I + rho A'A) x~ = sigma x - q + A'(rho z - y), z~ = A x~
 *
 * PARITY UNPINNED: no golden vectors / KATs exist in the reference and it cannot run here; see
 * the header of qp_oracle.py and DESIGN.md.
 *
 * Input layout = Julia's SparseMatrixCSC: colptr[ncols+1], rowval[nnz], nzval[nnz], 64-bit
 * indices, index_base 0 or 1.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int64_t rows, cols;
    int64_t *ptr;   /* rows+1 */
    int64_t *idx;   /* nnz    */
    double *val;    /* nnz    */
} csr_t;

typedef struct {
    int64_t max_iter;      /* numIterations = 5000 */
    double eps_abs;        /* 1e-6 */
    double eps_rel;        /* 1e-6 */
    double rho;            /* 1    */
    double sigma;          /* 1e-6 */
    double alpha;          /* 1.6  */
    int32_t adaptive_rho;  /* adptRho = false */
    double rho_factor;     /* fctrRho = 5 */
    int64_t check_every;   /* numItrConv = 25 */
    double pcg_eps;        /* 1e-6  (plugin kwarg) */
    int64_t pcg_max_iter;  /* 1000 */
    int32_t precond;       /* 0 = none (reference), 1 = Jacobi */
    double time_limit_s;   /* <= 0: none.  Bounded-sample timing: stop at the first iteration
                              boundary after this many seconds (flag stays convNumItr). */
} oracle_settings;

typedef struct {
    int32_t conv_flag;     /* 1 convNumItr, 2 convAdmm, 3 convPrimDual */
    int64_t iterations;
    double rho_final;
    double res_prim, res_dual;
    int64_t rho_updates;
    int64_t cg_iters_total;
    double solve_seconds;
} oracle_info;

static double now_s(void) {
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return 0.0;
#endif
}

/* CSC (cols = ncols) -> CSR of the same matrix, and CSR of its transpose (= the CSC arrays). */
static void csc_to_csr(int64_t nrows, int64_t ncols, const int64_t *colptr, const int64_t *rowval,
                       const double *nzval, int64_t base, csr_t *out) {
    int64_t nnz = colptr[ncols] - base;
    out->rows = nrows; out->cols = ncols;
    out->ptr = (int64_t *)calloc((size_t)nrows + 1, sizeof(int64_t));
    out->idx = (int64_t *)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(int64_t));
    out->val = (double *)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(double));
    for (int64_t k = 0; k < nnz; ++k) out->ptr[rowval[k] - base + 1]++;
    for (int64_t i = 0; i < nrows; ++i) out->ptr[i + 1] += out->ptr[i];
    int64_t *cur = (int64_t *)malloc((size_t)(nrows > 0 ? nrows : 1) * sizeof(int64_t));
    memcpy(cur, out->ptr, (size_t)nrows * sizeof(int64_t));
    for (int64_t j = 0; j < ncols; ++j)
        for (int64_t k = colptr[j] - base; k < colptr[j + 1] - base; ++k) {
            int64_t i = rowval[k] - base;
            int64_t p = cur[i]++;
            out->idx[p] = j; out->val[p] = nzval[k];
        }
    free(cur);
}

static void csc_as_csr_transpose(int64_t nrows, int64_t ncols, const int64_t *colptr, const int64_t *rowval,
                                 const double *nzval, int64_t base, csr_t *out) {
    int64_t nnz = colptr[ncols] - base;
    out->rows = ncols; out->cols = nrows;
    out->ptr = (int64_t *)malloc(((size_t)ncols + 1) * sizeof(int64_t));
    out->idx = (int64_t *)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(int64_t));
    out->val = (double *)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(double));
    for (int64_t j = 0; j <= ncols; ++j) out->ptr[j] = colptr[j] - base;
    for (int64_t k = 0; k < nnz; ++k) { out->idx[k] = rowval[k] - base; out->val[k] = nzval[k]; }
}

static void csr_free(csr_t *a) { free(a->ptr); free(a->idx); free(a->val); }

/* y = M x */
static void spmv(const csr_t *M, const double *x, double *y) {
#pragma omp parallel for schedule(static, 512)
    for (int64_t i = 0; i < M->rows; ++i) {
        double s = 0.0;
        for (int64_t k = M->ptr[i]; k < M->ptr[i + 1]; ++k) s += M->val[k] * x[M->idx[k]];
        y[i] = s;
    }
}

/* Dot product with a summation order that does NOT depend on the number of threads: fixed blocks of DOT_BLOCK
 * elements are summed left to right, the block sums are added left to right.  (An OpenMP `reduction(+)` partitions by
 * thread count, so the same test gave different exit checks on an 8-, a 16- and a 32-core host whenever the ADMM
 * stop test fired on rounding noise.  With this the oracle is bit-reproducible from box to box.) */
#define DOT_BLOCK 4096
static double dot(const double *a, const double *b, int64_t n) {
    const int64_t nb = (n + DOT_BLOCK - 1) / DOT_BLOCK;
    if (nb <= 1) {
        double s = 0.0;
        for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
        return s;
    }
    double stack_part[1024];
    double *part = nb <= 1024 ? stack_part : (double *)malloc((size_t)nb * sizeof(double));
#pragma omp parallel for schedule(static)
    for (int64_t blk = 0; blk < nb; ++blk) {
        const int64_t i0 = blk * DOT_BLOCK, i1 = i0 + DOT_BLOCK < n ? i0 + DOT_BLOCK : n;
        double s = 0.0;
        for (int64_t i = i0; i < i1; ++i) s += a[i] * b[i];
        part[blk] = s;
    }
    double s = 0.0;
    for (int64_t blk = 0; blk < nb; ++blk) s += part[blk];
    if (part != stack_part) free(part);
    return s;
}

static double norm_inf(const double *a, int64_t n) {
    double s = 0.0;
#pragma omp parallel for reduction(max : s) schedule(static)
    for (int64_t i = 0; i < n; ++i) { double v = fabs(a[i]); if (v > s || v != v) s = v; }
    return s;
}

typedef struct {
    const csr_t *P, *A, *At;
    int64_t n, m;
    double *vZZ;      /* the plugin's m-vector, also the operator's scratch (LinearSystemSolvers.jl:153) */
    double *tmp_n;
    double rho, sigma;
} kop_t;

/* u = P w + rho A'(A w) + sigma w   (LinearSystemSolvers.jl:152-157) */
static void apply_K(kop_t *K, const double *w, double *u) {
    spmv(K->A, w, K->vZZ);
    spmv(K->At, K->vZZ, u);
    spmv(K->P, w, K->tmp_n);
    const double rho = K->rho, sigma = K->sigma;
    double *t = K->tmp_n;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < K->n; ++i) {
        double v = t[i] + rho * u[i];
        u[i] = v + sigma * w[i];
    }
}

/* IterativeSolvers.cg!(x, K, b; abstol, maxiter[, Pl = Diagonal(d)]) -- returns #iterations */
static int64_t cg(kop_t *K, double *x, const double *b, double abstol, int64_t maxiter, const double *dinv_or_null,
                  double *u, double *r, double *c) {
    const int64_t n = K->n;
    const double reltol = sqrt(2.220446049250313e-16);
    memset(u, 0, (size_t)n * sizeof(double));
    apply_K(K, x, c);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) r[i] = b[i] - c[i];
    double residual = sqrt(dot(r, r, n));
    double tol = fmax(reltol * residual, abstol);
    double prev_residual = 1.0, rho_pcg = 1.0;
    int64_t it = 0;
    while (it < maxiter && !(residual <= tol)) {
        if (!dinv_or_null) {
            double beta = (residual * residual) / (prev_residual * prev_residual);
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < n; ++i) u[i] = r[i] + beta * u[i];
            apply_K(K, u, c);
            double alpha = (residual * residual) / dot(u, c, n);
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < n; ++i) { x[i] += alpha * u[i]; r[i] -= alpha * c[i]; }
            prev_residual = residual;
        } else {
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < n; ++i) c[i] = r[i] * dinv_or_null[i];
            double rho_prev = rho_pcg;
            rho_pcg = dot(c, r, n);
            double beta = rho_pcg / rho_prev;
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < n; ++i) u[i] = c[i] + beta * u[i];
            apply_K(K, u, c);
            double alpha = rho_pcg / dot(u, c, n);
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < n; ++i) { x[i] += alpha * u[i]; r[i] -= alpha * c[i]; }
        }
        residual = sqrt(dot(r, r, n));
        ++it;
    }
    return it;
}

static double clampd(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); }

/* Sparse solve, modes M (precond = 0) and J (precond = 1).  x_inout: start point in, solution out. */
int oracle_solve_sparse(int64_t n, int64_t m,
                        const int64_t *Pcolptr, const int64_t *Prowval, const double *Pnzval,
                        const int64_t *Acolptr, const int64_t *Arowval, const double *Anzval,
                        const double *q, const double *l, const double *u_bound, int64_t index_base,
                        const oracle_settings *s, double *x_inout, double *z_out, double *y_out, oracle_info *info) {
    csr_t P, A, At;
    csc_to_csr(n, n, Pcolptr, Prowval, Pnzval, index_base, &P);
    csc_to_csr(m, n, Acolptr, Arowval, Anzval, index_base, &A);
    csc_as_csr_transpose(m, n, Acolptr, Arowval, Anzval, index_base, &At);

    double *vX = x_inout;
    double *vXX = (double *)calloc((size_t)n + 1, sizeof(double));
    double *vZZ = (double *)calloc((size_t)m + 1, sizeof(double));
    double *vT = (double *)calloc((size_t)n + 1, sizeof(double));
    double *vXP = (double *)calloc((size_t)n + 1, sizeof(double));
    double *vZ = (double *)calloc((size_t)m + 1, sizeof(double));
    double *vY = (double *)calloc((size_t)m + 1, sizeof(double));
    double *vZP = (double *)calloc((size_t)m + 1, sizeof(double));
    double *cu = (double *)calloc((size_t)n + 1, sizeof(double));
    double *cr = (double *)calloc((size_t)n + 1, sizeof(double));
    double *cc = (double *)calloc((size_t)n + 1, sizeof(double));
    double *tmpn = (double *)calloc((size_t)n + 1, sizeof(double));
    double *tmpm = (double *)calloc((size_t)m + 1, sizeof(double));
    double *tmpn2 = (double *)calloc((size_t)n + 1, sizeof(double));
    double *dP = NULL, *dAA = NULL, *dinv = NULL;
    if (s->precond) {
        dP = (double *)calloc((size_t)n + 1, sizeof(double));
        dAA = (double *)calloc((size_t)n + 1, sizeof(double));
        dinv = (double *)calloc((size_t)n + 1, sizeof(double));
        for (int64_t i = 0; i < n; ++i)
            for (int64_t k = P.ptr[i]; k < P.ptr[i + 1]; ++k) if (P.idx[k] == i) dP[i] += P.val[k];
        for (int64_t j = 0; j < n; ++j)
            for (int64_t k = At.ptr[j]; k < At.ptr[j + 1]; ++k) dAA[j] += At.val[k] * At.val[k];
    }

    double rho = s->rho, rho1 = 1.0 / rho;                            /* :30 */
    const double alpha = s->alpha, alpha1 = 1.0 - alpha;              /* :31 */
    const double sigma = s->sigma;
    int32_t convFlag = 1;                                             /* :33 */
    const double epsAdmm = fmin(s->eps_abs, s->eps_rel) * 1e-2;       /* :34 */
    double rhorho = rho;                                              /* :43 */
    const double normQ = norm_inf(q, n);
    kop_t K = {&P, &A, &At, n, m, vZZ, tmpn, rho, sigma};
    int64_t ii = 0, cg_total = 0, rho_updates = 0;
    double resP = NAN, resD = NAN;
    int dinv_dirty = 1;
    const double t0 = now_s();

    for (ii = 1; ii <= s->max_iter; ++ii) {                           /* :45 */
        if (s->adaptive_rho && ((rhorho * s->rho_factor < rho) || (rhorho > s->rho_factor * rho))) {   /* :47 */
            rho = rhorho; rho1 = 1.0 / rho; ++rho_updates; dinv_dirty = 1;
        }
        K.rho = rho;
        if (s->precond && dinv_dirty) {
            for (int64_t i = 0; i < n; ++i) dinv[i] = 1.0 / (dP[i] + sigma + rho * dAA[i]);
            dinv_dirty = 0;
        }
        /* LinOpCg!  (LinearSystemSolvers.jl:178-183) */
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < m; ++i) vZZ[i] = rho * vZ[i] - vY[i];
        spmv(&At, vZZ, vT);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) vT[i] = sigma * vX[i] - q[i] + vT[i];
        cg_total += cg(&K, vXX, vT, s->pcg_eps, s->pcg_max_iter, s->precond ? dinv : NULL, cu, cr, cc);
        spmv(&A, vXX, vZZ);

        /* :56-61 */
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) { vXP[i] = vX[i]; vX[i] = alpha * vXX[i] + alpha1 * vX[i]; }
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < m; ++i) {
            double zp = vZ[i];
            vZP[i] = zp;
            double znew = clampd(alpha * vZZ[i] + alpha1 * zp + rho1 * vY[i], l[i], u_bound[i]);
            vZ[i] = znew;
            vY[i] = vY[i] + rho * (alpha * vZZ[i] + alpha1 * zp - znew);
        }

        if (ii % s->check_every == 0) {                               /* :63 */
            /* CheckConvergence :79-112 */
            spmv(&A, vX, tmpm);                 /* A x  */
            spmv(&P, vX, tmpn);                 /* P x  */
            spmv(&At, vY, tmpn2);               /* A' y */
            double nAx = norm_inf(tmpm, m), nZ = norm_inf(vZ, m);
            double nPx = norm_inf(tmpn, n), nAty = norm_inf(tmpn2, n);
            double rp = 0.0, rd = 0.0;
#pragma omp parallel for reduction(max : rp) schedule(static)
            for (int64_t i = 0; i < m; ++i) { double v = fabs(tmpm[i] - vZ[i]); if (v > rp || v != v) rp = v; }
#pragma omp parallel for reduction(max : rd) schedule(static)
            for (int64_t i = 0; i < n; ++i) { double v = fabs(tmpn[i] + q[i] + tmpn2[i]); if (v > rd || v != v) rd = v; }
            resP = rp; resD = rd;
            double maxNormPrim = fmax(nAx, nZ);
            double maxNormDual = fmax(fmax(nPx, nAty), normQ);
            if (s->adaptive_rho) {
                double num = rp * maxNormDual, den = rd * maxNormPrim;
                rhorho = clampd(rho * sqrt(num / den), 1e-3, 1e6);
            }
            double epsPrim = s->eps_abs + s->eps_rel * maxNormPrim;
            double epsDual = s->eps_abs + s->eps_rel * maxNormDual;
            if (rp < epsPrim && rd < epsDual) convFlag = 3;
            double dx = 0.0, dz = 0.0;
#pragma omp parallel for reduction(max : dx) schedule(static)
            for (int64_t i = 0; i < n; ++i) { double v = fabs(vX[i] - vXP[i]); if (v > dx || v != v) dx = v; }
#pragma omp parallel for reduction(max : dz) schedule(static)
            for (int64_t i = 0; i < m; ++i) { double v = fabs(vZ[i] - vZP[i]); if (v > dz || v != v) dz = v; }
            if (dx <= epsAdmm && dz <= epsAdmm) convFlag = 2;
            if (convFlag != 1) break;
        }
        if (s->time_limit_s > 0.0 && now_s() - t0 > s->time_limit_s) break;
    }
    if (ii > s->max_iter) ii = s->max_iter;

    info->conv_flag = convFlag;
    info->iterations = ii;
    info->rho_final = rho;
    info->res_prim = resP; info->res_dual = resD;
    info->rho_updates = rho_updates;
    info->cg_iters_total = cg_total;
    info->solve_seconds = now_s() - t0;
    if (z_out) memcpy(z_out, vZ, (size_t)m * sizeof(double));
    if (y_out) memcpy(y_out, vY, (size_t)m * sizeof(double));

    free(vXX); free(vZZ); free(vT); free(vXP); free(vZ); free(vY); free(vZP);
    free(cu); free(cr); free(cc); free(tmpn); free(tmpm); free(tmpn2);
    free(dP); free(dAA); free(dinv);
    csr_free(&P); csr_free(&A); csr_free(&At);
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Dense batched solve (mode D on the reduced system).  Layout: P[b] n x n column-major,
 * A[b] m x n column-major, q[b][n], l/u[b][m], X[b][n] in/out.  One problem per OpenMP task.
 * ------------------------------------------------------------------------------------------- */
static int chol_factor(double *K, int64_t n) {   /* lower, column-major, in place */
    for (int64_t j = 0; j < n; ++j) {
        double d = K[j + j * n];
        for (int64_t k = 0; k < j; ++k) d -= K[j + k * n] * K[j + k * n];
        if (!(d > 0.0)) return -1;
        d = sqrt(d);
        K[j + j * n] = d;
        for (int64_t i = j + 1; i < n; ++i) {
            double v = K[i + j * n];
            for (int64_t k = 0; k < j; ++k) v -= K[i + k * n] * K[j + k * n];
            K[i + j * n] = v / d;
        }
    }
    return 0;
}

static void chol_solve(const double *L, int64_t n, double *b) {
    for (int64_t i = 0; i < n; ++i) {
        double v = b[i];
        for (int64_t k = 0; k < i; ++k) v -= L[i + k * n] * b[k];
        b[i] = v / L[i + i * n];
    }
    for (int64_t i = n - 1; i >= 0; --i) {
        double v = b[i];
        for (int64_t k = i + 1; k < n; ++k) v -= L[k + i * n] * b[k];
        b[i] = v / L[i + i * n];
    }
}

static void build_K(const double *P, const double *A, int64_t n, int64_t m, double rho, double sigma, double *K) {
    for (int64_t j = 0; j < n; ++j)
        for (int64_t i = j; i < n; ++i) {
            double s = 0.0;
            for (int64_t k = 0; k < m; ++k) s += A[k + i * m] * A[k + j * m];
            K[i + j * n] = P[i + j * n] + rho * s + (i == j ? sigma : 0.0);
        }
}

int oracle_solve_dense_batch(int64_t batch, int64_t n, int64_t m, const double *P, const double *A,
                             const double *q, const double *l, const double *u_bound,
                             const oracle_settings *s, double *X, int32_t *flags, int64_t *iters, double *seconds) {
    int fail = 0;
    const double t0 = now_s();
#pragma omp parallel for schedule(dynamic, 4) reduction(| : fail)
    for (int64_t b = 0; b < batch; ++b) {
        const double *Pb = P + b * n * n, *Ab = A + b * m * n, *qb = q + b * n, *lb = l + b * m, *ub = u_bound + b * m;
        double *vX = X + b * n;
        double *K = (double *)malloc((size_t)(n * n) * sizeof(double));
        double *vXX = (double *)calloc((size_t)n, sizeof(double));
        double *vXP = (double *)calloc((size_t)n, sizeof(double));
        double *vZZ = (double *)calloc((size_t)m, sizeof(double));
        double *vZ = (double *)calloc((size_t)m, sizeof(double));
        double *vY = (double *)calloc((size_t)m, sizeof(double));
        double *vZP = (double *)calloc((size_t)m, sizeof(double));
        double *tn = (double *)calloc((size_t)n, sizeof(double));
        double *tn2 = (double *)calloc((size_t)n, sizeof(double));
        double *tm = (double *)calloc((size_t)m, sizeof(double));
        double rho = s->rho, rho1 = 1.0 / rho, rhorho = rho;
        const double alpha = s->alpha, alpha1 = 1.0 - alpha, sigma = s->sigma;
        const double epsAdmm = fmin(s->eps_abs, s->eps_rel) * 1e-2;
        int32_t convFlag = 1;
        double normQ = 0.0;
        for (int64_t i = 0; i < n; ++i) normQ = fmax(normQ, fabs(qb[i]));
        build_K(Pb, Ab, n, m, rho, sigma, K);
        if (chol_factor(K, n)) fail |= 1;
        int64_t ii;
        for (ii = 1; ii <= s->max_iter; ++ii) {
            if (s->adaptive_rho && ((rhorho * s->rho_factor < rho) || (rhorho > s->rho_factor * rho))) {
                rho = rhorho; rho1 = 1.0 / rho;
                build_K(Pb, Ab, n, m, rho, sigma, K);
                if (chol_factor(K, n)) fail |= 1;
            }
            for (int64_t i = 0; i < m; ++i) tm[i] = rho * vZ[i] - vY[i];
            for (int64_t j = 0; j < n; ++j) {
                double sacc = 0.0;
                for (int64_t i = 0; i < m; ++i) sacc += Ab[i + j * m] * tm[i];
                vXX[j] = sigma * vX[j] - qb[j] + sacc;
            }
            chol_solve(K, n, vXX);
            for (int64_t i = 0; i < m; ++i) vZZ[i] = 0.0;
            for (int64_t j = 0; j < n; ++j) { double xj = vXX[j]; for (int64_t i = 0; i < m; ++i) vZZ[i] += Ab[i + j * m] * xj; }
            for (int64_t i = 0; i < n; ++i) { vXP[i] = vX[i]; vX[i] = alpha * vXX[i] + alpha1 * vX[i]; }
            for (int64_t i = 0; i < m; ++i) {
                double zp = vZ[i]; vZP[i] = zp;
                double znew = clampd(alpha * vZZ[i] + alpha1 * zp + rho1 * vY[i], lb[i], ub[i]);
                vZ[i] = znew;
                vY[i] = vY[i] + rho * (alpha * vZZ[i] + alpha1 * zp - znew);
            }
            if (ii % s->check_every == 0) {
                for (int64_t i = 0; i < m; ++i) tm[i] = 0.0;
                for (int64_t j = 0; j < n; ++j) { double xj = vX[j]; for (int64_t i = 0; i < m; ++i) tm[i] += Ab[i + j * m] * xj; }
                for (int64_t i = 0; i < n; ++i) tn[i] = 0.0;
                for (int64_t j = 0; j < n; ++j) { double xj = vX[j]; for (int64_t i = 0; i < n; ++i) tn[i] += Pb[i + j * n] * xj; }
                for (int64_t j = 0; j < n; ++j) { double sacc = 0.0; for (int64_t i = 0; i < m; ++i) sacc += Ab[i + j * m] * vY[i]; tn2[j] = sacc; }
                double rp = 0, rd = 0, nAx = 0, nZ = 0, nPx = 0, nAty = 0, dx = 0, dz = 0;
                for (int64_t i = 0; i < m; ++i) { rp = fmax(rp, fabs(tm[i] - vZ[i])); nAx = fmax(nAx, fabs(tm[i])); nZ = fmax(nZ, fabs(vZ[i])); dz = fmax(dz, fabs(vZ[i] - vZP[i])); }
                for (int64_t i = 0; i < n; ++i) { rd = fmax(rd, fabs(tn[i] + qb[i] + tn2[i])); nPx = fmax(nPx, fabs(tn[i])); nAty = fmax(nAty, fabs(tn2[i])); dx = fmax(dx, fabs(vX[i] - vXP[i])); }
                double maxNormPrim = fmax(nAx, nZ), maxNormDual = fmax(fmax(nPx, nAty), normQ);
                if (s->adaptive_rho) rhorho = clampd(rho * sqrt((rp * maxNormDual) / (rd * maxNormPrim)), 1e-3, 1e6);
                if (rp < s->eps_abs + s->eps_rel * maxNormPrim && rd < s->eps_abs + s->eps_rel * maxNormDual) convFlag = 3;
                if (dx <= epsAdmm && dz <= epsAdmm) convFlag = 2;
                if (convFlag != 1) break;
            }
        }
        if (ii > s->max_iter) ii = s->max_iter;
        if (flags) flags[b] = convFlag;
        if (iters) iters[b] = ii;
        free(K); free(vXX); free(vXP); free(vZZ); free(vZ); free(vY); free(vZP); free(tn); free(tn2); free(tm);
    }
    if (seconds) *seconds = now_s() - t0;
    return fail ? -5 : 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py's CPU arm: use `nt` threads whatever OMP_NUM_THREADS says (torch.distributed.run exports
 * OMP_NUM_THREADS=1 to every rank, which would silently time a single-threaded baseline). */
int oracle_set_num_threads(int nt) {
#ifdef _OPENMP
    if (nt > 0) omp_set_num_threads(nt);
    return omp_get_max_threads();
#else
    (void)nt;
    return 1;
#endif
}
