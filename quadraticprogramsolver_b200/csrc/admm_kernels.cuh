// admm_kernels.cuh -- the persistent ADMM kernel (whole solve on the device) and the stand-alone
// SpMV kernels built on spmv_core.cuh.
//
// Reference path restated on the GPU (file:line in /root/reference):
//   SolveQuadraticProgram!  SolveQuadraticProgram.jl:14-76     outer loop, relaxation, clip, dual update
//   CheckConvergence        SolveQuadraticProgram.jl:79-112    residual inf-norms, adaptive rho, flags
//   LinOpCg!                LinearSystemSolvers.jl:145-186     rhs, matrix-free K, z~ = A x~
//   IterativeSolvers.cg!    (third party, v0.9.x algorithm)    CGIterable / PCGIterable
//
// Data layout: the operator is held as two tiled CSR matrices,
//     H = [P  A']   n x (n+m)   (row j = row j of P followed by column j of A, columns shifted by n)
//     A             m x n
// and the vectors that H gathers from are stored as contiguous pairs [n-part ; m-part]:
//     XY = [x ; y]      XG = [x~ ; g]  with g = rho (z~ - z) + y      UT = [u ; rho A u]
// so that  K u = H*UT + sigma u,   b - K x~ = sigma (x - x~) - q - H*XG,   [Px ; A'y] = split sums of H*XY.
#pragma once
#include "spmv_core.cuh"
#include "direct_kernels.cuh"

namespace qpb {

struct AdmmSettingsDev {
    long long max_iter;
    long long check_every;
    long long pcg_max_iter;
    double eps_abs, eps_rel, rho, sigma, alpha, rho_factor, pcg_eps, pcg_rel_eps;
    int adaptive_rho;
};

struct AdmmInfoDev {
    int conv_flag;
    int pad;
    long long iterations;
    double rho_final, res_prim, res_dual;
    long long rho_updates, pcg_iters_total, pcg_maxed;
    long long n_h_passes, n_a_passes;   // matrix passes executed (for the roofline accounting)
};

struct SparseProblemDev {
    int n, m;
    CsrTiled H, A;
    const double *q, *l, *u;    // problem vectors
    const double *dP, *dAA;     // diag(P), column square sums of A (Jacobi)
    double *XY, *XG, *UT;       // vector pairs, n + m each
    double *z, *zt;             // m
    double *r, *c, *zp, *dinv;  // n
    double *wv;                 // n: w = K z of the one-reduction PCG
    double normQ;
    // optional (Ruiz equilibration): D, 1/(c D) [n] and 1/E [m] turn the norms of CheckConvergence back into
    // those of the unscaled problem; nullptr = the problem is solved as given (the reference's behaviour)
    const double *Dv, *Dinvc, *Einv;
    // optional per-constraint step size (qpb200_set_rho_scale; not in the reference): rho_i = rho * rs[i].
    // nullptr = one scalar rho as in SolveQuadraticProgram.jl:16; dAA then holds sum_i rs[i] A_ij^2
    const double *rs;
    // direct x~ step (settings.lin_solver = QPB200_LINSOLVE_CHOLESKY, direct_kernels.cuh): Kneg = -(P + sigma I +
    // A' diag(rho_i) A)^-1 for the rho in rho0, dense row-major, leading dimension ldk; nullptr on the CG path
    const double *Kneg;
    int ldk;
    // where this launch starts: a fresh solve has iter0 = 0, rho0 = rhorho0 = settings.rho, resume_changed = 0.  The
    // direct path leaves the kernel when the rho trigger fires (conv_flag 0 = "refactorise"), the host rebuilds Kneg
    // and re-enters with the iteration count, the adopted rho and resume_changed = 1 (g depends on rho)
    long long iter0;
    double rho0, rhorho0;
    int resume_changed;
    GridSync gs;
    AdmmSettingsDev s;
    AdmmInfoDev *info;
};

// Book-keeping of a solve that only thread 0 of a CTA touches: kept in shared memory, not in registers (the tile loop
// needs them: with these as loop-carried registers ptxas spilled the epilogues' accumulators inside the tile loops).
struct AdmmCounters {
    long long rho_updates, pcg_total, pcg_maxed, n_h, n_a;
    double res_prim, res_dual;
};

__device__ __forceinline__ double clamp_julia(double x, double lo, double hi) {
    return x > hi ? hi : (x < lo ? lo : x);   // Julia's clamp: NaN passes through
}

// =============================================================================================
// The statements of the outer iteration that every arrangement of the solve shares (admm_kernel here, the segments of
// admm_dist_kernel, admm_peer_sliced_kernel): ONE copy of the reference's update and stop-test arithmetic.
// =============================================================================================
// SolveQuadraticProgram.jl:59-61 for constraint row i, given z~_i = (A x~)_i: relaxation, clip, dual update, and
// g_i = rho_i (z~_i - z_i) + y_i for the next right-hand side.  Returns |z_new - z_old| (the dz of :105).
__device__ __forceinline__ double admm_row_update(const SparseProblemDev &p, double *y, double *g, int i, double zt_i,
                                                  double alpha, double alpha1, double rho_i, double rho1_i) {
    const double z_old = p.z[i], y_old = y[i];
    const double zr = alpha * zt_i + alpha1 * z_old;
    const double z_new = clamp_julia(zr + rho1_i * y_old, p.l[i], p.u[i]);   // :60
    const double y_new = y_old + rho_i * (zr - z_new);                       // :61
    p.z[i] = z_new;
    y[i] = y_new;
    p.zt[i] = zt_i;
    g[i] = rho_i * (zt_i - z_new) + y_new;
    return fabs(z_new - z_old);
}

// SolveQuadraticProgram.jl:57 for variable j: x = alpha x~ + (1 - alpha) x.  Returns |x_new - x_old| (the dx of :105).
__device__ __forceinline__ double admm_x_relax(double *x, const double *xt, int j, double alpha, double alpha1) {
    const double x_old = x[j];
    const double x_new = alpha * xt[j] + alpha1 * x_old;
    x[j] = x_new;
    return fabs(x_new - x_old);
}

// CheckConvergence, primal side (:85-86,90): (A x)_i against z_i -> res = |Ax - z|, mx = max(|Ax|, |z|); e_i = 1 unless
// the problem was equilibrated (then the norms are those of the unscaled QP)
__device__ __forceinline__ void admm_prim_norms(double &res, double &mx, double ax_i, double z_i, double e_i) {
    res = nanmax(res, fabs(ax_i - z_i) * e_i);
    mx = nanmax(mx, fabs(ax_i) * e_i);
    mx = nanmax(mx, fabs(z_i) * e_i);
}

// CheckConvergence, dual side (:87-89,91): res = |Px + q + A'y|, mx = max(|Px|, |A'y|) (|q| joins after the reduction)
__device__ __forceinline__ void admm_dual_norms(double &res, double &mx, double px_j, double q_j, double aty_j, double d_j) {
    res = nanmax(res, fabs(px_j + q_j + aty_j) * d_j);
    mx = nanmax(mx, fabs(px_j) * d_j);
    mx = nanmax(mx, fabs(aty_j) * d_j);
}

// CheckConvergence from the reduced norms (:92-107): adaptive-rho proposal (clamped, NaN passes through as in Julia)
// and the two stop tests in the reference's order -- convAdmm overrides convPrimDual.
__device__ __forceinline__ void admm_stop_test(const AdmmSettingsDev &s, double rho, double dx, double dz, double res_prim,
                                               double res_dual, double max_prim, double max_dual, double &rhorho,
                                               int &conv_flag) {
    if (s.adaptive_rho) {                                             // :92-96
        const double num = res_prim * max_dual, den = res_dual * max_prim;
        rhorho = clamp_julia(rho * sqrt(num / den), 1e-3, 1e6);
    }
    const double eps_prim = s.eps_abs + s.eps_rel * max_prim;         // :99
    const double eps_dual = s.eps_abs + s.eps_rel * max_dual;         // :100
    if ((res_prim < eps_prim) && (res_dual < eps_dual)) conv_flag = 3;     // :102
    const double eps_admm = fmin(s.eps_abs, s.eps_rel) * 1e-2;             // :34
    if ((dx <= eps_admm) && (dz <= eps_admm)) conv_flag = 2;               // :105 (overrides)
}

// =============================================================================================
// Stand-alone SpMV (operator unit tests, SpMV roofline measurement).
//   mode 0: y[rows] = M x            mode 1 (split): y0 = M[:, :split] x[:split], y1 = M[:, split:] x[split:]
// =============================================================================================
template <int TMA, bool SPLIT>
__global__ void __launch_bounds__(kThreads, kMinCtas) spmv_kernel(CsrTiled M, const double *x, double *y0, double *y1) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SpmvSmem &sm = *reinterpret_cast<SpmvSmem *>(smem_raw);
    PipeState ps;
    spmv_smem_init(sm, ps);
    auto epi = [&](int row, double s0, double s1) {
        y0[row] = s0;
        if (SPLIT) y1[row] = s1;
    };
    spmv_tiles<TMA, SPLIT>(M, x, sm, ps, epi);
}

// =============================================================================================
// The persistent ADMM kernel.  Cooperative launch, grid = co-resident CTAs (<= 2 per SM).
// Every scalar (rho, alpha_cg, beta, residuals, flags) is recomputed identically by every thread
// from bit-identical all-reduced values, so control flow is uniform across the grid.
// =============================================================================================
// CGV = false: the recurrence of IterativeSolvers' CGIterable / PCGIterable as written (two reductions and four
//   grid barriers per iteration).
// CGV = true (default, settings.reserved_i[QPB200_RSV_CG_RECURRENCE] = 0): the Chronopoulos-Gear arrangement of the
//   SAME preconditioned CG -- gamma = r.z, delta = z.(K z) and |r|^2 come out of ONE reduction fused into the H pass,
//   beta = gamma / gamma_prev, alpha = gamma / (delta - beta gamma / alpha_prev), and p = z + beta p, s = w + beta s
//   (= K p by linearity), x~ += alpha p, r -= alpha s, z = Pl \ r are one vector pass: three grid barriers and three
//   phases per iteration instead of four and four.  Same iterates in exact arithmetic, same stopping rule on the same
//   recurrence residual; one more operator application per inner solve (the K z of the final residual is unused).
// DIRECT = true: the x~ step is exact (the reference's LaLdl!/QDLdl!/FacLdl! plugins, LinearSystemSolvers.jl:16-107):
//   x~ += K^-1 (b - K x~) with the dense inverse of direct_kernels.cuh; a rho change ends the launch (see iter0 above).
template <int TMA, bool PRE, bool CGV, bool DIRECT = false>
__global__ void __launch_bounds__(kThreads, kMinCtas) admm_kernel(SparseProblemDev p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SpmvSmem &sm = *reinterpret_cast<SpmvSmem *>(smem_raw);
    __shared__ AdmmCounters ctr;
    PipeState ps;
    spmv_smem_init(sm, ps);
    SyncState st;
    st.epoch = 0;
    if (threadIdx.x == 0) {
        ctr.rho_updates = ctr.pcg_total = ctr.pcg_maxed = ctr.n_h = ctr.n_a = 0;
        ctr.res_prim = ctr.res_dual = nan("");
    }
    auto count = [&](long long &c, long long by) { if (threadIdx.x == 0) c += by; };

    const int n = p.n, m = p.m;
    const int gtid = blockIdx.x * kThreads + threadIdx.x;
    const int gstride = gridDim.x * kThreads;
    double *const x = p.XY, *const y = p.XY + n;
    double *const xt = p.XG, *const g = p.XG + n;
    double *const u = p.UT, *const t = p.UT + n;
    double *const zpv = PRE ? p.zp : p.r;   // un-preconditioned: "z" of PCG is r itself

    double rho = p.rho0, rho1 = 1.0 / rho;                          // SolveQuadraticProgram.jl:30 (rho0 = settings.rho on a fresh solve)
    // p.s.alpha, 1 - p.s.alpha (:31), p.s.sigma, p.s.pcg_rel_eps are used straight out of the parameter block (constant-bank operands)
    double rhorho = p.rhorho0;                                      // :43
    int conv_flag = 1;                                              // :33 convNumItr
    bool dinv_ready = false;
    // rho and 1/rho of constraint i (the scalars of SolveQuadraticProgram.jl:30 unless a scale vector was set)
    auto rho_of = [&](int i) { return p.rs ? rho * p.rs[i] : rho; };
    auto rho1_of = [&](int i) { return p.rs ? 1.0 / (rho * p.rs[i]) : rho1; };

    long long ii = 0;
    bool refactor = false;
    for (ii = p.iter0 + 1; ii <= p.s.max_iter; ++ii) {              // :45
        // ---- rho trigger (:46-52) -> "refactorisation": Jacobi diagonal and g depend on rho
        bool changed = DIRECT && ii == p.iter0 + 1 && p.resume_changed;
        if (p.s.adaptive_rho && ((rhorho * p.s.rho_factor < rho) || (rhorho > p.s.rho_factor * rho))) {
            if (DIRECT) {                                           // the host refactorises K for rhorho and re-enters at ii
                refactor = true;
                break;
            }
            rho = rhorho;
            rho1 = 1.0 / rho;
            changed = true;
            count(ctr.rho_updates, 1);
        }
        if (changed || !dinv_ready) {
            if (PRE)
                for (int j = gtid; j < n; j += gstride) p.dinv[j] = 1.0 / (p.dP[j] + p.s.sigma + rho * p.dAA[j]);
            if (changed)
                for (int i = gtid; i < m; i += gstride) g[i] = rho_of(i) * (p.zt[i] - p.z[i]) + y[i];
            dinv_ready = true;
            grid_barrier(p.gs, st);
        }

        // ---- [P1] r0 = b - K x~  with b = p.s.sigma x - q + A'(rho z - y)   (LinearSystemSolvers.jl:178-180
        //      and the first MV product of cg!); u = Pl \ r0
        double acc[2] = {0.0, 0.0};
        {
            auto epi = [&](int j, double s0, double) {
                const double rj = p.s.sigma * (x[j] - xt[j]) - p.q[j] - s0;
                p.r[j] = rj;
                const double zj = PRE ? p.dinv[j] * rj : rj;
                if (PRE && !CGV) p.zp[j] = zj;
                u[j] = zj;
                acc[0] += rj * rj;
                acc[1] += rj * zj;
            };
            spmv_tiles<TMA, false>(p.H, p.XG, sm, ps, epi);
            count(ctr.n_h, 1);
        }
        grid_barrier_reduce<2, false>(p.gs, st, acc, sm.red, sm.bcast);
        double residual = sqrt(acc[0]);
        double rz = acc[1];
        const double tol = fmax(p.s.pcg_rel_eps * residual, p.s.pcg_eps);     // cg_iterator!: max(p.s.pcg_rel_eps*|r0|, abstol)

        long long k = 0;
        if (DIRECT) {
            // ---- exact solve as one refinement step from the previous x~: x~ += K^-1 r0 (r0 from [P1])
            dense_symv_sub(p.Kneg, p.ldk, n, p.r, xt);
            grid_barrier(p.gs, st);
        } else if (CGV) {
            // ---- one-reduction PCG: u holds z = Pl \ r (the vector A and H gather from), p.zp the search direction,
            //      p.c holds s = K p, p.wv holds w = K z
            double gam = 0.0, a_cg = 0.0;
            bool first = true;
            while (k < p.s.pcg_max_iter && !(residual <= tol)) {
                {   // t = rho * A z
                    auto epi = [&](int i, double s0, double) { t[i] = rho_of(i) * s0; };
                    spmv_tiles<TMA, false>(p.A, u, sm, ps, epi);
                    count(ctr.n_a, 1);
                }
                grid_barrier(p.gs, st);
                double d3[2] = {0.0, 0.0};                             // gamma = r.z, delta = z.w
                {   // w = P z + A' t + p.s.sigma z      (LinearSystemSolvers.jl:152-157)
                    auto epi = [&](int j, double s0, double) {
                        const double zj = u[j];
                        const double wj = s0 + p.s.sigma * zj;
                        p.wv[j] = wj;
                        d3[0] += p.r[j] * zj;
                        d3[1] += zj * wj;
                    };
                    spmv_tiles<TMA, false>(p.H, p.UT, sm, ps, epi);
                    count(ctr.n_h, 1);
                }
                grid_barrier_reduce<2, false>(p.gs, st, d3, sm.red, sm.bcast);
                double beta = 0.0;
                if (first) {
                    if (!(d3[1] > 0.0)) break;                         // breakdown guard (K is SPD)
                    a_cg = d3[0] / d3[1];
                } else {
                    beta = d3[0] / gam;
                    const double den = d3[1] - beta * d3[0] / a_cg;
                    if (!(den > 0.0)) break;
                    a_cg = d3[0] / den;
                }
                gam = d3[0];
                // p = z + beta p ; s = w + beta s ; x~ += a p ; r -= a s ; z = Pl \ r ; |r|^2
                double rr[1] = {0.0};
                for (int j = gtid; j < n; j += gstride) {
                    double pj = u[j], sj = p.wv[j];
                    if (!first) {                                      // (stale p, s of the previous solve are never read)
                        pj += beta * p.zp[j];
                        sj += beta * p.c[j];
                    }
                    p.zp[j] = pj;
                    p.c[j] = sj;
                    xt[j] += a_cg * pj;
                    const double rj = p.r[j] - a_cg * sj;
                    p.r[j] = rj;
                    u[j] = PRE ? p.dinv[j] * rj : rj;
                    rr[0] += rj * rj;
                }
                first = false;
                grid_barrier_reduce<1, false>(p.gs, st, rr, sm.red, sm.bcast);
                residual = sqrt(rr[0]);
                ++k;
            }
        } else {
        // ---- PCG loop (CGIterable / PCGIterable)
        while (k < p.s.pcg_max_iter && !(residual <= tol)) {
            // [S2] t = rho * A u
            {
                auto epi = [&](int i, double s0, double) { t[i] = rho_of(i) * s0; };
                spmv_tiles<TMA, false>(p.A, u, sm, ps, epi);
                count(ctr.n_a, 1);
            }
            grid_barrier(p.gs, st);
            // [S3] c = P u + rho A'(A u) + p.s.sigma u ; u.c      (LinearSystemSolvers.jl:152-157)
            double uc[1] = {0.0};
            {
                auto epi = [&](int j, double s0, double) {
                    const double uj = u[j];
                    const double cj = s0 + p.s.sigma * uj;
                    p.c[j] = cj;
                    uc[0] += uj * cj;
                };
                spmv_tiles<TMA, false>(p.H, p.UT, sm, ps, epi);
                count(ctr.n_h, 1);
            }
            grid_barrier_reduce<1, false>(p.gs, st, uc, sm.red, sm.bcast);
            if (!(uc[0] > 0.0)) break;                               // breakdown guard (K is SPD)
            const double a_cg = rz / uc[0];
            // [S4] x~ += a u ; r -= a c ; z = Pl \ r ; |r|^2, r.z
            double acc2[2] = {0.0, 0.0};
            for (int j = gtid; j < n; j += gstride) {
                xt[j] += a_cg * u[j];
                const double rj = p.r[j] - a_cg * p.c[j];
                p.r[j] = rj;
                const double zj = PRE ? p.dinv[j] * rj : rj;
                if (PRE) p.zp[j] = zj;
                acc2[0] += rj * rj;
                acc2[1] += rj * zj;
            }
            grid_barrier_reduce<2, false>(p.gs, st, acc2, sm.red, sm.bcast);
            residual = sqrt(acc2[0]);
            const double rz_new = acc2[1];
            ++k;
            if (k < p.s.pcg_max_iter && !(residual <= tol)) {
                // [S1] u = z + beta u
                const double beta = rz_new / rz;
                for (int j = gtid; j < n; j += gstride) u[j] = zpv[j] + beta * u[j];
                grid_barrier(p.gs, st);
            }
            rz = rz_new;
        }
        }
        count(ctr.pcg_total, k);
        if (k >= p.s.pcg_max_iter && !(residual <= tol)) count(ctr.pcg_maxed, 1);

        // ---- [Upd] z~ = A x~ (LinearSystemSolvers.jl:183) fused with SolveQuadraticProgram.jl:56-61
        const bool do_check = (ii % p.s.check_every) == 0;           // :63
        double nrm[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};          // dx dz rp nAx nZ rd|nPx|nAty packed below
        {
            auto epi = [&](int i, double s0, double) {
                const double dz_i = admm_row_update(p, y, g, i, s0, p.s.alpha, 1.0 - p.s.alpha, rho_of(i), rho1_of(i));
                const double ei = (p.Einv && do_check) ? p.Einv[i] : 1.0;
                nrm[1] = nanmax(nrm[1], dz_i * ei);
            };
            spmv_tiles<TMA, false>(p.A, xt, sm, ps, epi);
            count(ctr.n_a, 1);
        }
        for (int j = gtid; j < n; j += gstride) {
            const double dx_j = admm_x_relax(x, xt, j, p.s.alpha, 1.0 - p.s.alpha);   // :57
            const double dj = (p.Dv && do_check) ? p.Dv[j] : 1.0;
            nrm[0] = nanmax(nrm[0], dx_j * dj);
        }
        grid_barrier(p.gs, st);

        if (do_check) {
            // ---- [Chk] CheckConvergence (:79-112): A x, P x, A' y and their inf-norms in two passes
            //      (ei, dj = 1 unless the problem was equilibrated: then the norms are those of the unscaled QP)
            {
                auto epi = [&](int i, double s0, double) {
                    admm_prim_norms(nrm[2], nrm[3], s0, p.z[i], p.Einv ? p.Einv[i] : 1.0);
                };
                spmv_tiles<TMA, false>(p.A, x, sm, ps, epi);
                count(ctr.n_a, 1);
            }
            {
                auto epi = [&](int j, double s0, double s1) {
                    admm_dual_norms(nrm[4], nrm[5], s0, p.q[j], s1, p.Dinvc ? p.Dinvc[j] : 1.0);
                };
                spmv_tiles<TMA, true>(p.H, p.XY, sm, ps, epi);
                count(ctr.n_h, 1);
            }
            grid_barrier_reduce<7, true>(p.gs, st, nrm, sm.red, sm.bcast);
            const double dx = nrm[0], dz = nrm[1];
            const double res_prim = nrm[2], res_dual = nrm[4];
            if (threadIdx.x == 0) { ctr.res_prim = res_prim; ctr.res_dual = res_dual; }
            const double max_prim = nrm[3];
            const double max_dual = nanmax(nrm[5], p.normQ);
            admm_stop_test(p.s, rho, dx, dz, res_prim, res_dual, max_prim, max_dual, rhorho, conv_flag);
            if (conv_flag != 1) break;                                // :66
        }
    }
    if (ii > p.s.max_iter) ii = p.s.max_iter;
    if (refactor) {                                                 // iterations completed so far; rho to adopt
        conv_flag = 0;
        ii -= 1;
        rho = rhorho;
    }

    if (blockIdx.x == 0 && threadIdx.x == 0) {
        AdmmInfoDev &o = *p.info;
        o.conv_flag = conv_flag;
        o.iterations = ii;
        o.rho_final = rho;
        o.res_prim = ctr.res_prim;
        o.res_dual = ctr.res_dual;
        o.rho_updates = ctr.rho_updates;
        o.pcg_iters_total = ctr.pcg_total;
        o.pcg_maxed = ctr.pcg_maxed;
        o.n_h_passes = ctr.n_h;
        o.n_a_passes = ctr.n_a;
    }
}

}  // namespace qpb
