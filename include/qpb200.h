/* qpb200.h -- C ABI of libqpb200.so: the B200-native (sm_100a) OSQP-style ADMM QP solver.
 *
 *     minimise 0.5 x'Px + q'x   subject to   l <= Ax <= u
 *
 * This is the boundary a Julia `ccall` (or any FFI) binds to replace the reference's hot path
 *     SolveQuadraticProgram!(vX, mP, vQ, mA, vL, vU, LinSysSolInit, LinSysSol!; kw...)
 *         /root/reference/SolveQuadraticProgram.jl:14-76   (ADMM driver)
 *         /root/reference/SolveQuadraticProgram.jl:79-112  (CheckConvergence)
 *         /root/reference/LinearSystemSolvers.jl:16-229    (the (Init, Sol!) KKT-solve plugins)
 * Plain C: pointers and sizes only, no exceptions, no torch/CUDA types.  Everything below a
 * handle runs on the GPU in hand-written CUDA kernels; there is no CPU fallback: on a machine
 * without an sm_100 device `*_create` fails with QPB200_ERR_DEVICE.
 *
 * Sparse matrices come in exactly as Julia's SparseMatrixCSC{Float64,Int64} stores them:
 * colptr[ncols+1], rowval[nnz], nzval[nnz], 64-bit indices, `index_base` = 1 from Julia (0 from
 * C / scipy).  P must hold both triangles (the reference multiplies by mP as given).
 * Host arrays are borrowed only for the duration of the call.
 */
#ifndef QPB200_H
#define QPB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QPB200_VERSION 100 /* 0.1.0 */

/* ---- status codes (0 = OK, negative = error; "did not converge" is NOT an error) ---------- */
#define QPB200_OK 0
#define QPB200_ERR_ARG (-1)       /* bad argument / inconsistent dimensions                      */
#define QPB200_ERR_NONFINITE (-2) /* NaN/Inf in P, A, q (l/u may be +-Inf) or l > u              */
#define QPB200_ERR_CUDA (-3)      /* CUDA runtime error (see qpb200_last_error)                  */
#define QPB200_ERR_NCCL (-4)      /* NCCL error / NCCL library not loadable                      */
#define QPB200_ERR_FACTOR (-5)    /* Cholesky breakdown (non-positive pivot) in the dense path   */
#define QPB200_ERR_DEVICE (-6)    /* no CUDA device of compute capability 10.x                   */

/* ---- ConvergenceFlag, numerically identical to the Julia enum (SolveQuadraticProgram.jl:12) */
#define QPB200_CONV_NUM_ITR 1   /* convNumItr   : iteration limit reached                        */
#define QPB200_CONV_ADMM 2      /* convAdmm     : ||x-xprev||inf, ||z-zprev||inf <= min(eps)*1e-2 */
#define QPB200_CONV_PRIM_DUAL 3 /* convPrimDual : primal and dual residual tests passed          */

/* ---- linear-system solver for the x~ step (replaces the (Init, Sol!) pair) ----------------- */
#define QPB200_LINSOLVE_PCG 0      /* matrix-free (P)CG on K = P + sigma I + rho A'A
                                      (LinOpCg!/LinMapsCg!, LinearSystemSolvers.jl:145-229)       */
#define QPB200_LINSOLVE_CHOLESKY 1 /* exact solve, the reduced form of the direct plugins LaLdl/QDLdl/FacLdl
                                      (:16-107; FacLdl is what RunTests.jl:55-56 runs).  qpb200_batch_*: dense
                                      Cholesky of K per QP in shared memory.  qpb200_create (n <= 32768): dense K in
                                      HBM inverted in place by a blocked symmetric sweep (FP64 tensor-pipe trailing
                                      updates), refactorised on every rho change, applied as x~ += K^-1 (b - K x~)   */
#define QPB200_PRECOND_NONE 0      /* IterativeSolvers.cg! exactly as the reference calls it      */
#define QPB200_PRECOND_JACOBI 1    /* Pl = Diagonal(diag(P) + sigma + rho colsumsq(A))            */

/* Settings: the keyword arguments of SolveQuadraticProgram! 1:1 (SolveQuadraticProgram.jl:15-17),
 * then the plugin kwargs (LinearSystemSolvers.jl:125), then what is new here.                    */
typedef struct qpb200_settings {
    int64_t max_iter;     /* numIterations = 5000                                                 */
    double eps_abs;       /* eps_abs (ϵAbs) = 1e-6                                                */
    double eps_rel;       /* eps_rel (ϵRel) = 1e-6                                                */
    double rho;           /* ρ = 1                                                                */
    double sigma;         /* σ = 1e-6                                                             */
    double alpha;         /* α = 1.6                                                              */
    double delta;         /* δ = 1e-6        regularisation of the polish KKT system (unused unless
                             reserved_i[QPB200_RSV_POLISH] is set; the reference never uses it)    */
    int32_t adaptive_rho; /* adptΡ = false                                                        */
    int32_t lin_solver;   /* QPB200_LINSOLVE_*                                                    */
    double rho_factor;    /* fctrΡ = 5                                                            */
    int64_t check_every;  /* numItrConv = 25                                                      */
    int64_t polish_iter;  /* numItrPolish = 10   refinement rounds of the polish (see QPB200_RSV_POLISH)  */
    double minres_eps;    /* ϵMinres = 1e-6      relative residual of each MINRES solve of the polish     */
    int64_t minres_iter;  /* numItrMinres = 500  iteration cap of each MINRES solve                       */
    double pcg_eps;       /* ϵPcg = 1e-6  (abstol of cg!)                                         */
    int64_t pcg_max_iter; /* numItrPcg = 1000                                                     */
    double pcg_rel_eps;   /* reltol of cg!; <0 means the library default sqrt(eps(Float64))       */
    int32_t precond;      /* QPB200_PRECOND_* (default JACOBI)                                    */
    int32_t device;       /* CUDA device ordinal; -1 = the calling thread's current device        */
    int32_t spmv_loader;  /* accepted and ignored since round 2: one tile loader (TMA bulk staged)   */
    int32_t reserved_i[7]; /* [QPB200_RSV_*] below; the rest must be 0                            */
    double reserved_d[4];
} qpb200_settings;

/* Meaning of settings.reserved_i[k] (all default 0 = the reference's behaviour)                     */
#define QPB200_RSV_CHOL_UNBLOCKED 0 /* dense batch: 1 = unblocked Cholesky (A/B runs)                  */
#define QPB200_RSV_DIST_MODE 1      /* qpb200_dist_*: 0 auto / peer memory, 1 NCCL, 2 peer required    */
#define QPB200_RSV_SCALING_ITERS 2  /* sparse single-GPU path: k > 0 = k iterations of modified Ruiz
                                       equilibration of [P A'; A 0] with cost scaling (OSQP paper,
                                       Algorithm 2; README.md:71-72 of the reference lists it as a
                                       TODO).  The scaled QP is iterated; CheckConvergence's norms
                                       (SolveQuadraticProgram.jl:85-105) are evaluated on the
                                       UNSCALED residuals, x / z / y are returned unscaled.            */

#define QPB200_RSV_DENSE_VARIANT 3  /* dense batch, m padded to 96: 0 = default, 1 = matrix-vector products out of
                                       shared memory, 2 = A held in registers during the iterations, 3 = A and
                                       K^-1 in registers (A/B runs)                                               */

#define QPB200_RSV_CG_RECURRENCE 4  /* sparse single-GPU path: 1 = the recurrence exactly as IterativeSolvers' CGIterable /
                                       PCGIterable writes it (two reductions, four grid barriers per iteration);
                                       2 = one-reduction arrangement of the same (P)CG (Chronopoulos-Gear: same iterates
                                       in exact arithmetic, same stopping rule at the same point, three grid barriers
                                       per iteration); 0 = auto: 2 for problems whose iteration is barrier-latency bound
                                       (nnz(P) + 2 nnz(A) <= 16 M: measured 14.5 vs 15.5 us per CG iteration at
                                       configs[0], 30.2 vs 32.8 at configs[1]), 1 for larger ones (the extra vector
                                       traffic costs more than the barrier there)                                     */

#define QPB200_RSV_POLISH 5         /* sparse single-GPU path: 1 = polish the solution after the ADMM loop.  The Julia
                                       driver accepts numItrPolish, delta, eps_minres, numItrMinres and never uses them
                                       (SolveQuadraticProgram.jl:16-17), so 0 (default) is the reference's behaviour; 1 runs
                                       the MATLAB twin's polish (SolveQuadraticProgram.m:289-325): numItrPolish rounds of
                                       iterative refinement of the delta-regularised KKT system of the active constraints,
                                       each solved by MINRES (eps_minres, numItrMinres) on the device; x is replaced only if
                                       the last MINRES solve converged.  Active rows: z_i - l_i < -y_i (lower),
                                       u_i - z_i < y_i (upper) instead of MATLAB's sign(y_i) -- see polish_kernels.cuh.     */

#define QPB200_RSV_BATCH_CHUNK 6    /* qpb200_batch_solve_once: problems per pipeline chunk (0 = 4096)                */

typedef struct qpb200_info {
    int32_t conv_flag;       /* QPB200_CONV_*                                                      */
    int32_t polish_status;   /* 0 = polish not requested, 1 = applied, 2 = MINRES did not converge (x untouched) */
    int64_t iterations;      /* ADMM iterations executed (a multiple of check_every unless capped) */
    double rho_final;
    double res_prim;         /* ||Ax - z||inf at the last check                                    */
    double res_dual;         /* ||Px + q + A'y||inf at the last check                              */
    int64_t rho_updates;     /* times the rho trigger fired (= refactorisations)                   */
    int64_t pcg_iters_total; /* CG iterations summed over all ADMM iterations                      */
    int64_t pcg_maxed;       /* ADMM iterations whose CG stopped on pcg_max_iter                   */
    double solve_ms;         /* device time of the solve, CUDA events                              */
    double setup_ms;         /* host wall time of create (conversion + upload)                     */
    int64_t kernel_launches; /* kernels launched by the last solve                                 */
    int64_t polish_minres_iters; /* MINRES iterations summed over the polish rounds                */
    int64_t polish_active;   /* constraints in the polish's active set                             */
} qpb200_info;

typedef struct qpb200_handle qpb200_handle;             /* one sparse QP on one GPU               */
typedef struct qpb200_batch qpb200_batch;               /* a batch of small dense QPs on one GPU  */

/* ---- library -------------------------------------------------------------------------------- */
int qpb200_version(void);
int qpb200_device_count(void);                          /* sm_100 devices visible, or error < 0   */
void qpb200_default_settings(qpb200_settings *s);       /* defaults of SolveQuadraticProgram.jl:15-17 */
const char *qpb200_last_error(void);                    /* thread-local message of the last failure */

/* ---- single sparse QP: replaces SolveQuadraticProgram! + LinOpCgInit/LinOpCg! ---------------- */
int qpb200_create(qpb200_handle **out, int64_t n, int64_t m,
                  const int64_t *P_colptr, const int64_t *P_rowval, const double *P_nzval,
                  const int64_t *A_colptr, const int64_t *A_rowval, const double *A_nzval,
                  const double *q, const double *l, const double *u,
                  const qpb200_settings *settings, int32_t index_base);
/* x_inout[n]: start point in (vX), solution out.  z_out[m], y_out[m] may be NULL.              */
int qpb200_solve(qpb200_handle *h, double *x_inout, double *z_out, double *y_out, qpb200_info *info);
int qpb200_update_vectors(qpb200_handle *h, const double *q, const double *l, const double *u); /* any may be NULL */
int qpb200_update_settings(qpb200_handle *h, const qpb200_settings *settings);
/* Per-constraint step size (OSQP's rho vector; the reference's README.md:71-72 TODO):
 * rho_i = rho * rho_scale[i] in every place SolveQuadraticProgram.jl:56-61 and LinearSystemSolvers.jl:152-157,
 * 178 use the scalar, i.e. K = P + sigma I + A' diag(rho_i) A, z = clamp(.. + y_i / rho_i), y += rho_i (..);
 * CheckConvergence and the adaptive update keep working on the scalar rho.  rho_scale[m] > 0 (OSQP: 1e3 on rows
 * with l == u); NULL restores the scalar.  Takes effect at the next qpb200_solve.  Single-GPU sparse path only. */
int qpb200_set_rho_scale(qpb200_handle *h, const double *rho_scale);
void qpb200_destroy(qpb200_handle *h);

/* ---- the reference's second solver, ProxQP.jl, on the same kernels ------------------------
 *     min 0.5 x'Px + q'x   s.t.   A x = b,   C x <= d
 * replaces SolveQuadraticProgram!(sQpProb::ProxQP; numIterations = 2000, eps_abs = 1e-7, eps_rel = 1e-6,
 * numItrConv = 50, rho = 1e2, sigma = 1e-2, adptRho = true, tau = 10)  (/root/reference/ProxQP.jl:118-173) with its
 * Cholesky of M = P + rho (A'A + C'C) + sigma I (:175-205) and CheckConvergence! (:250-296).
 * The handle comes from qpb200_create with lin_solver = QPB200_LINSOLVE_CHOLESKY, constraint matrix [A; C] (the first
 * m_eq rows are the equalities), l ignored, u = [b; d].  Settings read here: max_iter, eps_abs, eps_rel, check_every,
 * rho, sigma, adaptive_rho, rho_factor (= tau); qpb200_proxqp_default_settings fills in the reference's defaults.
 * x[n], y[m_eq], z[m - m_eq], s[m - m_eq] are the start point in (the inner constructor ProxQP.jl:36 takes them
 * explicitly) and the final iterates out; init_slack != 0: s_in is ignored and the slack starts from
 * s = max(d - C x, 0) (:109), computed on the device (s may then also be NULL: the final slack is not returned).
 * As in the reference all max_iter iterations run; report->iterations is the iteration of the last converged check. */
typedef struct qpb200_proxqp_report {
    int32_t converged;       /* dReport["Converged"]: convFlag of the last check                   */
    int32_t reserved;
    int64_t iterations;      /* dReport["Iterations"]                                              */
    double rho;              /* dReport["rho"] (final)                                             */
    double sigma;
    double res_prim;         /* dReport["PrimalResidual"]                                          */
    double res_dual;         /* dReport["DualResidual"]                                            */
    int64_t rho_updates;     /* refactorisations triggered by the adaptive rho                     */
    double solve_ms;         /* device time, CUDA events (refactorisations included)               */
    int64_t kernel_launches;
} qpb200_proxqp_report;
void qpb200_proxqp_default_settings(qpb200_settings *s);
int qpb200_proxqp_solve(qpb200_handle *h, int64_t m_eq, const qpb200_settings *settings, double *x, double *y, double *z,
                        double *s, int32_t init_slack, qpb200_proxqp_report *report);

/* The operators of the path on their own (SparseArrays mul!, SolveQuadraticProgram.jl:85-89,
 * LinearSystemSolvers.jl:135,139,153-155): which = 0: y[n] = P x[n]; 1: y[m] = A x[n];
 * 2: y[n] = A' x[m]; 3: y[n] = (P + sigma I + rho A'A) x[n].  Host vectors in and out.          */
int qpb200_apply(qpb200_handle *h, int32_t which, const double *x, double *y);
/* Same operator, device-resident vectors, `reps` back-to-back launches timed with CUDA events; an
 * L2-flushing write precedes every launch when flush_l2 != 0.  ms_out = mean ms per launch.      */
int qpb200_time_apply(qpb200_handle *h, int32_t which, int32_t reps, int32_t flush_l2, double *ms_out);
/* Algorithmic bytes of one launch of `which` (SURVEY.md section 8(d) formula).                    */
int64_t qpb200_apply_bytes(qpb200_handle *h, int32_t which);

/* ---- batch of small dense QPs (MPC-style): replaces SolveQuadraticProgram! + a direct plugin --
 * P: batch blocks n x n column-major; A: batch blocks m x n column-major; q[batch*n];
 * l, u[batch*m]; X_inout[batch*n]; flags/iters[batch] (may be NULL).                              */
int qpb200_batch_create(qpb200_batch **out, int64_t batch, int64_t n, int64_t m,
                        const double *P, const double *A, const double *q, const double *l, const double *u,
                        const qpb200_settings *settings);
int qpb200_batch_solve(qpb200_batch *h, double *X_inout, int32_t *flags, int64_t *iters, qpb200_info *info);
/* MPC-style re-solve: new q[batch*n], l, u[batch*m] (any may be NULL) for the matrices already on the device;
 * the next qpb200_batch_solve uploads nothing but the start points.                                            */
int qpb200_batch_update_vectors(qpb200_batch *h, const double *q, const double *l, const double *u);
void qpb200_batch_destroy(qpb200_batch *h);
/* The MPC-style batch proper: every problem has the SAME P[n x n] and A[m x n] (column-major,
 * passed once), only q[batch*n], l, u[batch*m] differ.  One K^-1 for the batch; 16 problems at a time are the columns
 * of FP64 tensor-pipe GEMMs (n <= 64, m <= 96; adaptive_rho must be 0).  Same handle type: qpb200_batch_solve,
 * qpb200_batch_update_vectors and qpb200_batch_destroy apply.                                                       */
int qpb200_batch_create_shared(qpb200_batch **out, int64_t batch, int64_t n, int64_t m,
                               const double *P, const double *A, const double *q, const double *l, const double *u,
                               const qpb200_settings *settings);
/* A batch that is solved once (create + solve + destroy in one call, what SolveQuadraticProgramBatch does): the batch
 * is cut into chunks of settings.reserved_i[QPB200_RSV_BATCH_CHUNK] problems (0 = 4096) and chunk c + 1 is uploaded
 * while the kernel of chunk c runs; only two chunks are resident.  Same results as create + solve, bit for bit.
 * info->solve_ms = sum of the chunks' kernel times, info->setup_ms = wall time of the whole call.                 */
int qpb200_batch_solve_once(int64_t batch, int64_t n, int64_t m, const double *P, const double *A, const double *q,
                            const double *l, const double *u, const qpb200_settings *settings, double *X_inout,
                            int32_t *flags, int64_t *iters, qpb200_info *info);

/* ---- one large sparse QP row-partitioned over several GPUs, one rank (process or thread) per GPU.
 * Rank r owns rows [row_begin, row_end) of A (and of l, u, z, y) and the same-numbered share of P's
 * columns; the CSC arrays passed are that slice: A_slice is (row_end-row_begin) x n with LOCAL row
 * indices, P_slice is n x n holding only the columns [pcol_begin, pcol_end).  x, q are replicated.
 * nccl_unique_id: 128 bytes from qpb200_dist_unique_id on rank 0, distributed by the caller; 128 zero bytes
 * mean "reuse the communicator this process created for its previous distributed handle" (same rank, nranks
 * and device) -- ncclCommInitRank costs about a second, a solve often less.
 * settings.reserved_i[1] selects the collective: 0 = all-reduce inside the persistent kernel over NVLink peer
 * memory (cudaIpc) when every rank can map every peer, else NCCL; 1 = ncclAllReduce + host-driven kernel
 * segments; 2 = peer path required (error if unavailable).                                                  */
int qpb200_dist_unique_id(void *id128);
int qpb200_dist_create(qpb200_handle **out, int32_t rank, int32_t nranks, const void *nccl_unique_id,
                       int64_t n, int64_t m_local,
                       const int64_t *P_colptr, const int64_t *P_rowval, const double *P_nzval,
                       const int64_t *A_colptr, const int64_t *A_rowval, const double *A_nzval,
                       const double *q, const double *l_local, const double *u_local,
                       const qpb200_settings *settings, int32_t index_base);
/* The same with the partition done inside: every rank passes the WHOLE QP as the caller holds it
 * (P n x n, A m x n, q[n], l[m], u[m]); the library computes the nnz-balanced row blocks of A / column blocks of P,
 * cuts this rank's slice out on all host threads (P's columns without a copy) and proceeds as qpb200_dist_create.
 * qpb200_dist_rows returns the rows [row_begin, row_end) of A whose z, y this rank's qpb200_dist_solve returns.   */
int qpb200_dist_create_full(qpb200_handle **out, int32_t rank, int32_t nranks, const void *nccl_unique_id,
                            int64_t n, int64_t m,
                            const int64_t *P_colptr, const int64_t *P_rowval, const double *P_nzval,
                            const int64_t *A_colptr, const int64_t *A_rowval, const double *A_nzval,
                            const double *q, const double *l, const double *u,
                            const qpb200_settings *settings, int32_t index_base);
int qpb200_dist_rows(qpb200_handle *h, int64_t *row_begin, int64_t *row_end);
/* Collective: every rank calls it.  x_inout[n] (replicated), z_out/y_out[m_local] or NULL.        */
int qpb200_dist_solve(qpb200_handle *h, double *x_inout, double *z_out, double *y_out, qpb200_info *info);

/* ---- host-only introspection (no GPU needed; used by the CPU test-suite) --------------------------
 * Builds the tile plan the kernels consume for a CSR matrix given by rowptr[rows+1] (int32, 0-based):
 * tiles_out[4*i..] = {first row, #rows, first nnz, #nnz | flags} (flags: bit30 continues a long row,
 * bit29 the row continues), cta_begin_out[grid+1] = tile range of each CTA.  Returns the number of
 * tiles (> tiles_cap means tiles_out was too small), or an error < 0.  *lpr_out = lanes per row.      */
int64_t qpb200_debug_tile_plan(int32_t rows, const int32_t *rowptr, int32_t grid, int32_t *tiles_out,
                               int64_t tiles_cap, int32_t *cta_begin_out, int32_t *lpr_out);
int32_t qpb200_debug_tile_nnz(void);   /* kTileNnz */
/* The host part of qpb200_create with scaling off, without a device: the operator H = [P A'] (n x (n+m), CSR,
 * 0-based int32; rowmid[j] = where row j switches from P's entries to column j of A), diag(P) and the column
 * square sums of A (the Jacobi preconditioner's ingredients).  Output arrays sized n+1, n, nnzP+nnzA (x2), n, n;
 * any may be NULL.                                                                                         */
int qpb200_debug_assemble_h(int64_t n, int64_t m, const int64_t *P_colptr, const int64_t *P_rowval,
                            const double *P_nzval, const int64_t *A_colptr, const int64_t *A_rowval,
                            const double *A_nzval, int32_t index_base, int32_t *rowptr_out, int32_t *rowmid_out,
                            int32_t *col_out, double *val_out, double *diagP_out, double *colsqA_out);
/* The host part of qpb200_create with scaling on, without a device: CSC -> row-major copies -> `iters`
 * iterations of the equilibration (QPB200_RSV_SCALING_ITERS).  Writes D[n], E[m], *c and the scaled q[n],
 * and the scaled values of P and A back in CSC order (Pnzval_out[nnzP], Anzval_out[nnzA]; NULL to skip).  */
int qpb200_debug_equilibrate(int64_t n, int64_t m, const int64_t *P_colptr, const int64_t *P_rowval,
                             const double *P_nzval, const int64_t *A_colptr, const int64_t *A_rowval,
                             const double *A_nzval, const double *q, int32_t iters, int32_t index_base,
                             double *D_out, double *E_out, double *c_out, double *q_out, double *Pnzval_out,
                             double *Anzval_out);

/* The partition qpb200_dist_create_full uses, without a device: row_bounds_out[nranks+1] (rows of A),
 * col_bounds_out[nranks+1] (columns of P); returns nranks or an error < 0.  qpb200_debug_slice additionally cuts the
 * slice of `rank`: bounds4_out = {row_begin, row_end, pcol_begin, pcol_end}, *p_off_out = first non-zero of P used,
 * Pcolptr_out[n+1] relative to it, the row slice of A as CSC with local row indices (a_cap = capacity of the A arrays). */
int64_t qpb200_debug_partition(int64_t n, int64_t m, const int64_t *P_colptr, const int64_t *A_colptr,
                               const int64_t *A_rowval, int32_t index_base, int32_t nranks,
                               int64_t *row_bounds_out, int64_t *col_bounds_out);
int qpb200_debug_slice(int64_t n, int64_t m, const int64_t *P_colptr, const int64_t *A_colptr, const int64_t *A_rowval,
                       const double *A_nzval, int32_t index_base, int32_t rank, int32_t nranks, int64_t *bounds4_out,
                       int64_t *p_off_out, int64_t *Pcolptr_out, int64_t *Acolptr_out, int64_t *Arowval_out,
                       double *Anzval_out, int64_t a_cap);

#ifdef __cplusplus
}
#endif
#endif /* QPB200_H */
