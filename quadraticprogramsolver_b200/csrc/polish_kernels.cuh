// polish_kernels.cuh -- solution polish after the ADMM loop, one cooperative launch.
//
// Reference: the Julia driver reserves numItrPolish, delta, eps_minres, numItrMinres and never uses them
// (SolveQuadraticProgram.jl:16-17); the algorithm is the MATLAB twin's, SolveQuadraticProgram.m:289-325:
//     active rows from the multipliers;  g = [-q; l_L; u_U];  K = [P A_act'; A_act 0];  KK = K + blkdiag(delta I, -delta I)
//     numPolishItr rounds:  tt = minres(KK, g - K t, eps_minres, numItrMinres, x0 = tt);  first failure stops;  t += tt
//     x = t[1:n] only if the last minres call converged.
// One deliberate change: a row is active only if its multiplier exceeds its distance to the bound
// (OSQP's rule: lower <=> z_i - l_i < -y_i, upper <=> u_i - z_i < y_i) instead of MATLAB's sign(y_i), which files
// rounding noise around 0 under whichever bound it points at.  oracle/qp_oracle.py::polish_solution is the restatement
// this kernel is tested against, statement for statement.
//
// B200 design: the reduced KKT system is kept at FULL size n + m.  Inactive rows carry the equation -delta nu_i = 0,
// so their entries stay exactly zero, and the operator is two independent masked passes of the tile engine
// -- A (bottom block: mask_i (A v_x)_i - delta v_nu) and H = [P A'] (top block: P v_x + A' v_nu + delta v_x), both
// gathering from the same contiguous pair [v_x; v_nu] -- with MINRES' dot products fused into their epilogues.
// One MINRES iteration = [A pass | H pass] -> reduce(alfa) -> vector pass -> reduce(beta) -> vector pass -> barrier;
// every scalar of the Lanczos / QR recurrences is computed by thread 0 of every CTA from bit-identical reduced
// values (uniform control flow, no host round trips, reproducible).
#pragma once
#include "admm_kernels.cuh"

namespace qpb {

struct PolishDev {
    double *T, *TT, *G, *B;      // n + m each: accumulated solution, minres iterate (correction), g, right-hand side
    double *V, *Y[3], *W[3];     // n + m each: Lanczos vector, the three most recent unnormalised ones, direction vectors
    double *mask;                // m: 1 on active rows
    double delta, tol;
    long long polish_iter, minres_iter;
    long long *out;              // [0] status (1 applied, 2 minres failed), [1] minres iterations, [2] active rows
};

struct MinresScalars {           // shared memory: written by thread 0, read by all after a bar.sync
    double oldb, beta, dbar, epsln, phibar, cs, sn;
    double c_r1, c_r2;           // coefficients of the two vector passes
    double oldeps, dlt, denom, phi, inv_beta;
    int done;
    long long total_its;
};

template <int TMA>
__global__ void __launch_bounds__(kThreads, kMinCtas) polish_kernel(SparseProblemDev p, PolishDev d) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SpmvSmem &sm = *reinterpret_cast<SpmvSmem *>(smem_raw);
    __shared__ MinresScalars ms;
    PipeState ps;
    spmv_smem_init(sm, ps);
    SyncState st;
    st.epoch = 0;
    const int n = p.n, m = p.m, N = n + m;
    const int gtid = blockIdx.x * kThreads + threadIdx.x;
    const int gstride = gridDim.x * kThreads;
    const double *y_dual = p.XY + n;

    // ---- active sets, g, t = tt = 0
    double cnt[1] = {0.0};
    for (int i = gtid; i < m; i += gstride) {
        const double zi = p.z[i], yi = y_dual[i], li = p.l[i], ui = p.u[i];
        double gi = 0.0, a = 0.0;
        if (zi - li < -yi) { a = 1.0; gi = li; }
        else if (ui - zi < yi) { a = 1.0; gi = ui; }
        d.mask[i] = a;
        d.G[n + i] = gi;
        cnt[0] += a;
    }
    for (int j = gtid; j < n; j += gstride) d.G[j] = -p.q[j];
    for (int e = gtid; e < N; e += gstride) { d.T[e] = 0.0; d.TT[e] = 0.0; }
    grid_barrier_reduce<1, false>(p.gs, st, cnt, sm.red, sm.bcast);
    if (blockIdx.x == 0 && threadIdx.x == 0) d.out[2] = (long long)cnt[0];
    if (threadIdx.x == 0) ms.total_its = 0;           // book-keeping of thread 0 lives in shared memory, not in registers
    int status = 2;                                   // until a round converges
    for (long long jj = 0; jj < d.polish_iter; ++jj) {
        // ---- b = g - K t (no regularisation);  Y0 = b - KK tt;  |b|^2, |Y0|^2
        double acc[2] = {0.0, 0.0};
        {
            auto epi = [&](int i, double s0, double) {
                const double bi = d.mask[i] != 0.0 ? d.G[n + i] - s0 : 0.0;
                d.B[n + i] = bi;
                acc[0] += bi * bi;
            };
            spmv_tiles<TMA, false>(p.A, d.T, sm, ps, epi);
        }
        {
            auto epi = [&](int j, double s0, double) {
                const double bj = d.G[j] - s0;
                d.B[j] = bj;
                acc[0] += bj * bj;
            };
            spmv_tiles<TMA, false>(p.H, d.T, sm, ps, epi);
        }
        {
            auto epi = [&](int i, double s0, double) {
                const double r = d.B[n + i] - (d.mask[i] * s0 - d.delta * d.TT[n + i]);
                d.Y[0][n + i] = r;
                acc[1] += r * r;
            };
            spmv_tiles<TMA, false>(p.A, d.TT, sm, ps, epi);
        }
        {
            auto epi = [&](int j, double s0, double) {
                const double r = d.B[j] - (s0 + d.delta * d.TT[j]);
                d.Y[0][j] = r;
                acc[1] += r * r;
            };
            spmv_tiles<TMA, false>(p.H, d.TT, sm, ps, epi);
        }
        grid_barrier_reduce<2, false>(p.gs, st, acc, sm.red, sm.bcast);
        const double tolb = d.tol * sqrt(acc[0]);
        const double beta1 = sqrt(acc[1]);
        bool converged = beta1 <= tolb;
        if (!converged) {
            if (threadIdx.x == 0) {
                ms.oldb = 0.0; ms.beta = beta1; ms.dbar = 0.0; ms.epsln = 0.0; ms.phibar = beta1; ms.cs = -1.0; ms.sn = 0.0;
                ms.done = 0;
            }
            const double ib = 1.0 / beta1;
            for (int e = gtid; e < N; e += gstride) {
                d.V[e] = d.Y[0][e] * ib;
                d.W[0][e] = 0.0; d.W[1][e] = 0.0; d.W[2][e] = 0.0;
            }
            grid_barrier(p.gs, st);
            long long k = 1;
            for (; k <= d.minres_iter; ++k) {
                double *Yk = d.Y[k % 3], *Y1 = d.Y[(k + 2) % 3], *Y2 = d.Y[(k + 1) % 3];   // k, k-1, k-2
                double *Wk = d.W[k % 3], *W1 = d.W[(k + 2) % 3], *W2 = d.W[(k + 1) % 3];
                const double c_old = (k >= 2) ? ms.beta / ms.oldb : 0.0;
                // ---- y = KK v - (beta/oldb) r1 ;  alfa = v . y
                double al[1] = {0.0};
                {
                    auto epi = [&](int i, double s0, double) {
                        const double vi = d.V[n + i];
                        double yv = d.mask[i] * s0 - d.delta * vi;
                        if (k >= 2) yv -= c_old * Y2[n + i];
                        Yk[n + i] = yv;
                        al[0] += vi * yv;
                    };
                    spmv_tiles<TMA, false>(p.A, d.V, sm, ps, epi);
                }
                {
                    auto epi = [&](int j, double s0, double) {
                        const double vj = d.V[j];
                        double yv = s0 + d.delta * vj;
                        if (k >= 2) yv -= c_old * Y2[j];
                        Yk[j] = yv;
                        al[0] += vj * yv;
                    };
                    spmv_tiles<TMA, false>(p.H, d.V, sm, ps, epi);
                }
                grid_barrier_reduce<1, false>(p.gs, st, al, sm.red, sm.bcast);
                const double alfa = al[0];
                // ---- y -= (alfa/beta) r2 ;  |y|^2
                const double c2 = alfa / ms.beta;
                double bb[1] = {0.0};
                for (int e = gtid; e < N; e += gstride) {
                    const double yv = Yk[e] - c2 * Y1[e];
                    Yk[e] = yv;
                    bb[0] += yv * yv;
                }
                grid_barrier_reduce<1, false>(p.gs, st, bb, sm.red, sm.bcast);
                if (threadIdx.x == 0) {               // Lanczos / QR scalars (Paige & Saunders), same order as the oracle
                    const double beta_new = sqrt(bb[0]);
                    ms.oldb = ms.beta;
                    ms.beta = beta_new;
                    ms.oldeps = ms.epsln;
                    ms.dlt = ms.cs * ms.dbar + ms.sn * alfa;
                    const double gbar = ms.sn * ms.dbar - ms.cs * alfa;
                    ms.epsln = ms.sn * beta_new;
                    ms.dbar = -ms.cs * beta_new;
                    const double gamma = fmax(sqrt(gbar * gbar + beta_new * beta_new), 2.220446049250313e-16);
                    ms.cs = gbar / gamma;
                    ms.sn = beta_new / gamma;
                    ms.phi = ms.cs * ms.phibar;
                    ms.phibar = ms.sn * ms.phibar;
                    ms.denom = 1.0 / gamma;
                    ms.inv_beta = beta_new != 0.0 ? 1.0 / beta_new : 0.0;
                    ms.done = (ms.phibar <= tolb || beta_new == 0.0) ? 1 : 0;
                }
                __syncthreads();
                const double oldeps = ms.oldeps, dlt = ms.dlt, denom = ms.denom, phi = ms.phi, ibn = ms.inv_beta;
                const int done = ms.done;
                // ---- w = (v - oldeps w1 - delta w2) / gamma ;  x += phi w ;  v = y / beta
                for (int e = gtid; e < N; e += gstride) {
                    const double w = (d.V[e] - oldeps * W2[e] - dlt * W1[e]) * denom;
                    Wk[e] = w;
                    d.TT[e] += phi * w;
                    d.V[e] = Yk[e] * ibn;
                }
                grid_barrier(p.gs, st);
                if (done) { converged = true; break; }
            }
            if (threadIdx.x == 0) ms.total_its += (k <= d.minres_iter) ? k : d.minres_iter;
        }
        if (!converged) { status = 2; break; }
        status = 1;
        for (int e = gtid; e < N; e += gstride) d.T[e] += d.TT[e];
        grid_barrier(p.gs, st);
    }
    if (d.polish_iter <= 0) status = 2;
    if (status == 1)
        for (int j = gtid; j < n; j += gstride) p.XY[j] = d.T[j];   // x = t[1:n]  (SolveQuadraticProgram.m:322-325)
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        d.out[0] = status;
        d.out[1] = ms.total_its;
    }
}

}  // namespace qpb
