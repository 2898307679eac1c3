// sparse_solver.h -- the object behind a qpb200_handle (single GPU, sparse QP).
#pragma once
#include "admm_kernels.cuh"
#include "polish_kernels.cuh"
#include "host_common.h"

namespace qpb {

constexpr int kDirectMaxN = 32768;        // dense n x n inverse: 8.6 GB at the limit

struct SparseSolver {
    int n = 0, m = 0, device = -1, num_sms = 0, grid = 0;
    int64_t nnzP = 0, nnzA = 0;
    int loader = 1;          // 0 LDG, 1 TMA (default), 2 TMA + software-pipelined gathers
    bool use_pre = true;
    // exact x~ step (settings.lin_solver = QPB200_LINSOLVE_CHOLESKY): dense -K^-1 [ldk x ldk] + sweep scratch
    bool direct = false, k_valid = false;
    double k_rho = 0.0;
    int ldk = 0;
    double *d_K = nullptr, *d_gjD = nullptr, *d_gjW = nullptr, *d_gjC = nullptr;
    int *d_gjStatus = nullptr;
    qpb200_settings settings{};
    SparseProblemDev prob{};
    AdmmInfoDev last_info{};
    DeviceArena arena;
    double *d_q = nullptr, *d_l = nullptr, *d_u = nullptr;
    double *d_dAA = nullptr, *d_rs = nullptr, *d_dAA_scaled = nullptr;   // column square sums of A: plain / weighted by rs
    double *scratch = nullptr, *flush_buf = nullptr;
    unsigned long long *sync_words = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double setup_ms = 0.0;
    // Ruiz equilibration (settings.reserved_i[QPB200_RSV_SCALING_ITERS] > 0): the device holds the scaled problem,
    // x = D x_s, z = z_s / E, y = E y_s / c; the kernel tests convergence on the unscaled residuals
    RuizScaling scaling;
    bool scaled = false;
    bool created = false;                  // settings_to_dev: later calls may not change what create fixed
    std::vector<double> h_l, h_u;          // the bounds as uploaded (scaled if the problem is), for update_vectors

    ~SparseSolver();
    int init(int64_t n, int64_t m, const int64_t *Pp, const int64_t *Pi, const double *Pv, const int64_t *Ap,
             const int64_t *Ai, const double *Av, const double *q, const double *l, const double *u,
             const qpb200_settings &s, int32_t base);
    int settings_to_dev(const qpb200_settings &s);
    int reset_state(const double *x0_host);
    int launch_admm();
    // solution polish (settings.reserved_i[QPB200_RSV_POLISH]): buffers allocated at the first use
    PolishDev pol{};
    bool pol_ready = false;
    long long pol_out[3] = {0, 0, 0};
    int polish();                                  // one cooperative launch on `stream` after the ADMM loop
    // the reference's second solver (ProxQP.jl) on an exact-solve handle holding [A; C]: proxqp_kernels.cuh
    int solve_proxqp(int64_t m_eq, const qpb200_settings &ps, double *x, double *y, double *z, double *s, bool init_slack,
                     qpb200_proxqp_report *report);
    int refactor(double rho, int64_t *launches);   // build K for rho and invert it in place
    bool one_reduction() const;   // which arrangement of the (P)CG recurrence admm_kernel runs (QPB200_RSV_CG_RECURRENCE)
    int solve(double *x_inout, double *z_out, double *y_out, qpb200_info *info);
    int apply(int which, const double *x_host, double *y_host);
    int apply_device(int which, const double *x, double *y);
    int time_apply(int which, int reps, int flush_l2, double *ms_out);
    int update_vectors(const double *q, const double *l, const double *u);
    int set_rho_scale(const double *rs);            // m positive factors or nullptr (scalar rho)
    int64_t spmv_bytes(int which) const;
    int64_t solve_bytes() const;
};

int prep_tile_kernel(const void *kernel, int *blocks_per_sm);   // shared-memory opt-in + L1 split of a tile-engine kernel

struct DistContext;                       // dist_solver.cu
void dist_destroy(DistContext *d);

}  // namespace qpb

// the object behind the opaque C handle
struct qpb200_handle {
    qpb::SparseSolver solver;
    qpb::DistContext *dist = nullptr;     // non-null for handles made by qpb200_dist_create
    ~qpb200_handle() { qpb::dist_destroy(dist); }
};
