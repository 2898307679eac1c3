// proxqp_kernels.cuh -- the reference's second solver on the same kernels:
//     SolveQuadraticProgram!(sQpProb::ProxQP; ...)    /root/reference/ProxQP.jl:118-173
//     CalculateRhs!, UpdateX!/S!/Y!/Z!                :207-247
//     CheckConvergence!                               :250-296
//     UpdateDecomposition! (refactor on a rho change) :191-205  -> direct_solver.cu (dense inverse, DMMA trailing updates)
//   min 0.5 x'Px + q'x  s.t.  A x = b,  C x <= d;  state x, y (equality duals), z >= 0 (inequality duals), s >= 0 (slack).
//
// The handle holds the stacked constraint matrix Abar = [A; C] (rows < m_eq are equalities) exactly as the ADMM path
// holds its A, so M = P + sigma I + rho (A'A + C'C) is the K of the exact-solve path and the tile matrices
// H = [P Abar'] and Abar are reused unchanged.  One iteration (three grid barriers, like the ADMM direct path):
//   [H pass]  r = -q - H [x; g]           with g = rho Abar x - w,  w = [rho b - y ; rho (d - s) - z]   (= rhs - M x)
//   [dense]   x += M^-1 r                                                            (UpdateX!, as one refinement step)
//   [A pass]  c = Abar x fused with UpdateS!, UpdateY!, UpdateZ! (same two-step roundings as :229-246), new w and g
// CheckConvergence! every numItrConv iterations: one Abar pass and two split H passes (the reference takes the norms
// of A'y and C'z separately), one fused 4-value max reduction.  As in the reference there is no early exit (:157 is
// commented out there): all numIterations run, the report carries the last converged check.
// A rho change ends the launch; the host refactorises and re-enters at the next iteration (direct path protocol).
#pragma once
#include "admm_kernels.cuh"

namespace qpb {

struct ProxInfoDev {
    int status;                  // 0 = finished, 1 = refactorise for `rho` and re-enter at iterations_done + 1
    int converged;               // convFlag of the last check
    long long iterations_done;
    long long conv_iter;         // iteration of the last converged check (0 = none yet)
    double rho, res_prim, res_dual;
    long long rho_updates;
};

struct ProxDev {
    int m_eq;                    // rows [0, m_eq) of Abar are A x = b, the rest C x <= d; b and d are p.u
    int init_s;                  // first launch only: s = max(d - C x, 0) from the start point (ProxQP.jl:109)
    double tau;                  // adaptive-rho threshold (tau = 10)
    ProxInfoDev carry;           // state of the report when re-entering
    ProxInfoDev *info;
};

template <int TMA>
__global__ void __launch_bounds__(kThreads, kMinCtas) proxqp_kernel(SparseProblemDev p, ProxDev d) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SpmvSmem &sm = *reinterpret_cast<SpmvSmem *>(smem_raw);
    PipeState ps;
    spmv_smem_init(sm, ps);
    SyncState st;
    st.epoch = 0;
    const int n = p.n, m = p.m, meq = d.m_eq;
    const int gtid = blockIdx.x * kThreads + threadIdx.x;
    const int gstride = gridDim.x * kThreads;
    double *const x = p.XG, *const g = p.XG + n;       // the pair H gathers from
    double *const D = p.zt;                            // duals [y; z]
    double *const sv = p.z;                            // slack (entries of equality rows stay 0)
    const double *const bd = p.u;                      // [b; d]
    double rho = p.rho0, rho1 = 1.0 / rho;
    int converged = d.carry.converged;
    long long conv_iter = d.carry.conv_iter, rho_updates = d.carry.rho_updates;
    double res_prim = d.carry.res_prim, res_dual = d.carry.res_dual;
    int status = 0;

    if (p.iter0 == 0 || p.resume_changed) {
        // g for the current rho from the current x, s, y, z (start of the solve / after a refactorisation)
        auto epi = [&](int i, double cx, double) {
            if (i >= meq && d.init_s && p.iter0 == 0) sv[i] = fmax(bd[i] - cx, 0.0);
            const double w = i < meq ? rho * bd[i] - D[i] : rho * (bd[i] - sv[i]) - D[i];
            g[i] = rho * cx - w;
        };
        spmv_tiles<TMA, false>(p.A, x, sm, ps, epi);
        grid_barrier(p.gs, st);
    }
    long long ii = p.iter0;
    for (ii = p.iter0 + 1; ii <= p.s.max_iter; ++ii) {                 // ProxQP.jl:135
        {   // r = rhs - M x = -q - (P x + Abar' g)                      (CalculateRhs! :207-218)
            auto epi = [&](int j, double s0, double) { p.r[j] = -p.q[j] - s0; };
            spmv_tiles<TMA, false>(p.H, p.XG, sm, ps, epi);
        }
        grid_barrier(p.gs, st);
        dense_symv_sub(p.Kneg, p.ldk, n, p.r, x);                       // UpdateX! :220-224
        grid_barrier(p.gs, st);
        {   // UpdateS! :226-232, UpdateY! :234-239, UpdateZ! :241-247 on c = Abar x, then w and g for the next rhs
            auto epi = [&](int i, double cx, double) {
                double w;
                if (i < meq) {
                    const double yv = (D[i] - rho * bd[i]) + rho * cx;
                    D[i] = yv;
                    w = rho * bd[i] - yv;
                } else {
                    const double di = bd[i], z_old = D[i];
                    const double s_new = fmax((di - rho1 * z_old) - cx, 0.0);
                    const double z_new = fmax((z_old + rho * (s_new - di)) + rho * cx, 0.0);
                    sv[i] = s_new;
                    D[i] = z_new;
                    w = rho * (di - s_new) - z_new;
                }
                g[i] = rho * cx - w;
            };
            spmv_tiles<TMA, false>(p.A, x, sm, ps, epi);
        }
        grid_barrier(p.gs, st);
        if (ii % p.s.check_every != 0) continue;                        // :151

        // ---- CheckConvergence! (:250-296)
        double nrm[4] = {0.0, 0.0, 0.0, 0.0};                           // rp, maxNormPrim, rd, maxNormDual
        for (int j = gtid; j < n; j += gstride) p.XY[j] = x[j];
        for (int i = gtid; i < m; i += gstride) p.XY[n + i] = i < meq ? D[i] : 0.0;
        grid_barrier(p.gs, st);
        {
            auto epi = [&](int i, double cx, double) {
                const double bi = bd[i];
                if (i < meq) {
                    nrm[0] = nanmax(nrm[0], fabs(cx - bi));             // |A x - b|
                } else {
                    nrm[0] = nanmax(nrm[0], fabs((cx - bi) + sv[i]));   // |C x - d + s|
                    nrm[1] = nanmax(nrm[1], fabs(sv[i]));
                }
                nrm[1] = nanmax(nrm[1], nanmax(fabs(cx), fabs(bi)));
            };
            spmv_tiles<TMA, false>(p.A, p.XY, sm, ps, epi);
        }
        {
            auto epi = [&](int j, double s0, double s1) {
                p.c[j] = s0;                                            // P x
                p.zp[j] = s1;                                           // A' y
                nrm[3] = nanmax(nrm[3], nanmax(fabs(s0), fabs(s1)));
            };
            spmv_tiles<TMA, true>(p.H, p.XY, sm, ps, epi);
        }
        grid_barrier(p.gs, st);
        for (int i = gtid; i < m; i += gstride) p.XY[n + i] = i < meq ? 0.0 : D[i];
        grid_barrier(p.gs, st);
        {
            auto epi = [&](int j, double, double s1) {                  // s1 = C' z
                const double qj = p.q[j];
                nrm[2] = nanmax(nrm[2], fabs(((p.c[j] + p.zp[j]) + s1) + qj));
                nrm[3] = nanmax(nrm[3], nanmax(fabs(s1), fabs(qj)));
            };
            spmv_tiles<TMA, true>(p.H, p.XY, sm, ps, epi);
        }
        grid_barrier_reduce<4, true>(p.gs, st, nrm, sm.red, sm.bcast);
        res_prim = nrm[0];
        res_dual = nrm[2];
        bool updated = false;
        if (p.s.adaptive_rho) {                                         // :274-283
            const double ratio = (nrm[0] * nrm[3]) / (nrm[2] * nrm[1]);
            if ((ratio > d.tau) || (1.0 / ratio > d.tau)) {
                updated = true;
                rho = clamp_julia(rho * sqrt(sqrt(ratio)), 1e-5, 1e5);
                rho1 = 1.0 / rho;
                rho_updates += 1;
            }
        }
        converged = ((nrm[0] < p.s.eps_abs + p.s.eps_rel * nrm[1]) && (nrm[2] < p.s.eps_abs + p.s.eps_rel * nrm[3])) ? 1 : 0;
        if (converged) conv_iter = ii;                                  // :156
        if (updated && ii < p.s.max_iter) {                             // the host refactorises M for the new rho (:159-165)
            status = 1;
            break;
        }
    }
    if (status == 0) ii = p.s.max_iter > p.iter0 ? p.s.max_iter : p.iter0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        ProxInfoDev &o = *d.info;
        o.status = status;
        o.converged = converged;
        o.iterations_done = ii;
        o.conv_iter = conv_iter;
        o.rho = rho;
        o.res_prim = res_prim;
        o.res_dual = res_dual;
        o.rho_updates = rho_updates;
    }
}

}  // namespace qpb
