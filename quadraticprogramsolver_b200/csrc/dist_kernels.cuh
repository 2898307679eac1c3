// dist_kernels.cuh -- the ADMM iteration of ONE large sparse QP row-partitioned over R GPUs.
//
// Rank r owns rows I_r of A (hence z, y, l, u, z~, g on I_r) and the columns J_r of P; x, q, x~ and the
// PCG vectors are replicated.  With H_r = [P[:, J_r]  A_r'] (n rows) the operator splits as
//     K u = sum_r H_r [u ; rho A_r u] + sigma u ,
// so every application needs exactly one all-reduce(sum) of an n-vector (ncclAllReduce over NVLink,
// issued by the host between two "segments" of this kernel); the residual inf-norms need one
// all-reduce(max) of four scalars per convergence check.  Because NCCL delivers bit-identical sums to
// all ranks and every replicated update is deterministic, all ranks take identical branches with no
// further communication (BASELINE.json north_star: "PCG dot products and A'y partial sums combined by
// NCCL allreduce").  Reference lines: same map as admm_kernels.cuh.
#pragma once
#include "admm_kernels.cuh"

namespace qpb {

struct DistState {            // device-resident control block, mirrored to pinned host memory
    unsigned long long epoch; // grid-barrier epoch carried across launches
    long long ii;             // ADMM iteration (1-based, of the iteration in progress)
    long long k;              // CG iteration within the current solve
    int cont;                 // CG continues (needs another all-reduce + step)
    int conv_flag;
    double rho, rho1, rhorho;
    double residual, tol, rz;
    double res_prim, res_dual;
    double lmax[4];           // dx, dz (local rows), |Ax - z| (local), max(|Ax|, |z|) (local)
    long long rho_updates, pcg_total, pcg_maxed, n_h, n_a;
};

enum DistSeg { kSegBegin = 0, kSegPcgInit = 1, kSegPcgStep = 2, kSegUpdate = 3, kSegCheck = 4 };

struct DistBuffers {
    DistState *state;
    double *wbuf;     // n      : partial / reduced  H_r * v
    double *wbuf2;    // 2 n    : [P x partial ; A' y partial]
};

template <int TMA, bool PRE>
__global__ void __launch_bounds__(kThreads, kMinCtas) admm_dist_kernel(SparseProblemDev p, DistBuffers d, int seg, int do_check) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SpmvSmem &sm = *reinterpret_cast<SpmvSmem *>(smem_raw);
    PipeState ps;
    spmv_smem_init(sm, ps);
    DistState &S = *d.state;
    SyncState st;
    st.epoch = S.epoch;

    const int n = p.n, m = p.m;
    const int gtid = blockIdx.x * kThreads + threadIdx.x;
    const int gstride = gridDim.x * kThreads;
    double *const x = p.XY, *const y = p.XY + n;
    double *const xt = p.XG, *const g = p.XG + n;
    double *const u = p.UT, *const t = p.UT + n;
    double *const zpv = PRE ? p.zp : p.r;
    const double sigma = p.s.sigma;
    const double alpha = p.s.alpha, alpha1 = 1.0 - alpha;

    // every thread reads the control block before the first barrier of this launch
    double rho = S.rho, rho1 = S.rho1, rhorho = S.rhorho;
    double residual = S.residual, tol = S.tol, rz = S.rz;
    long long ii = S.ii, k = S.k;
    int cont = S.cont, conv_flag = S.conv_flag;
    long long rho_updates = S.rho_updates, pcg_total = S.pcg_total, pcg_maxed = S.pcg_maxed, n_h = S.n_h, n_a = S.n_a;
    double res_prim = S.res_prim, res_dual = S.res_dual;
    double lmax[4] = {S.lmax[0], S.lmax[1], S.lmax[2], S.lmax[3]};
    __syncthreads();
    // a CG step enqueued speculatively after convergence: nothing to do, nothing to write (uniform exit)
    if (seg == kSegPcgStep && !cont) return;
    // NOTE: the control block is rewritten only at the very end, after at least one grid barrier
    // whenever any value changed (segments without a barrier write back what they read, plus counters
    // that only block 0 / thread 0 touches).
    bool had_barrier = false;

    auto spmv_A_t = [&]() {      // [S2] t = rho * A_r u
        auto epi = [&](int i, double s0, double) { t[i] = rho * s0; };
        spmv_tiles<TMA, false>(p.A, u, sm, ps, epi);
        ++n_a;
    };
    auto spmv_H_partial = [&](const double *pair) {   // wbuf = H_r * pair
        auto epi = [&](int j, double s0, double) { d.wbuf[j] = s0; };
        spmv_tiles<TMA, false>(p.H, pair, sm, ps, epi);
        ++n_h;
    };

    if (seg == kSegBegin) {
        ++ii;
        bool changed = false;
        if (p.s.adaptive_rho && ((rhorho * p.s.rho_factor < rho) || (rhorho > p.s.rho_factor * rho))) {
            rho = rhorho;
            rho1 = 1.0 / rho;
            changed = true;
            ++rho_updates;
        }
        if (changed || ii == 1) {
            if (PRE)
                for (int j = gtid; j < n; j += gstride) p.dinv[j] = 1.0 / (p.dP[j] + sigma + rho * p.dAA[j]);
            if (changed)
                for (int i = gtid; i < m; i += gstride) g[i] = rho * (p.zt[i] - p.z[i]) + y[i];
            grid_barrier(p.gs, st);
            had_barrier = true;
        }
        spmv_H_partial(p.XG);
    } else if (seg == kSegPcgInit) {
        // wbuf now holds sum_r H_r [x~ ; g_r]
        double acc[2] = {0.0, 0.0};
        for (int j = gtid; j < n; j += gstride) {
            const double rj = sigma * (x[j] - xt[j]) - p.q[j] - d.wbuf[j];
            p.r[j] = rj;
            const double zj = PRE ? p.dinv[j] * rj : rj;
            if (PRE) p.zp[j] = zj;
            u[j] = zj;
            acc[0] += rj * rj;
            acc[1] += rj * zj;
        }
        grid_barrier_reduce<2, false>(p.gs, st, acc, sm.red, sm.bcast);
        had_barrier = true;
        residual = sqrt(acc[0]);
        rz = acc[1];
        tol = fmax(p.s.pcg_rel_eps * residual, p.s.pcg_eps);
        k = 0;
        cont = (k < p.s.pcg_max_iter && !(residual <= tol)) ? 1 : 0;
        if (cont) {
            spmv_A_t();
            grid_barrier(p.gs, st);
            spmv_H_partial(p.UT);
        }
    } else if (seg == kSegPcgStep) {
        // wbuf = sum_r H_r [u ; rho A_r u]
        double uc[1] = {0.0};
        for (int j = gtid; j < n; j += gstride) {
            const double uj = u[j];
            const double cj = d.wbuf[j] + sigma * uj;
            p.c[j] = cj;
            uc[0] += uj * cj;
        }
        grid_barrier_reduce<1, false>(p.gs, st, uc, sm.red, sm.bcast);
        had_barrier = true;
        if (!(uc[0] > 0.0)) {
            cont = 0;
        } else {
            const double a_cg = rz / uc[0];
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
            cont = (k < p.s.pcg_max_iter && !(residual <= tol)) ? 1 : 0;
            if (cont) {
                const double beta = rz_new / rz;
                for (int j = gtid; j < n; j += gstride) u[j] = zpv[j] + beta * u[j];
                grid_barrier(p.gs, st);
                spmv_A_t();
                grid_barrier(p.gs, st);
                spmv_H_partial(p.UT);
            }
            rz = rz_new;
        }
    } else if (seg == kSegUpdate) {
        pcg_total += k;
        if (k >= p.s.pcg_max_iter && !(residual <= tol)) ++pcg_maxed;
        double nrm[4] = {0.0, 0.0, 0.0, 0.0};
        {
            auto epi = [&](int i, double s0, double) {
                nrm[1] = nanmax(nrm[1], admm_row_update(p, y, g, i, s0, alpha, alpha1, rho, rho1));
            };
            spmv_tiles<TMA, false>(p.A, xt, sm, ps, epi);
            ++n_a;
        }
        for (int j = gtid; j < n; j += gstride) nrm[0] = nanmax(nrm[0], admm_x_relax(x, xt, j, alpha, alpha1));
        if (do_check) {
            grid_barrier(p.gs, st);
            {
                auto epi = [&](int i, double s0, double) { admm_prim_norms(nrm[2], nrm[3], s0, p.z[i], 1.0); };
                spmv_tiles<TMA, false>(p.A, x, sm, ps, epi);
                ++n_a;
            }
            {
                auto epi = [&](int j, double s0, double s1) {
                    d.wbuf2[j] = s0;        // (P[:, J_r] x[J_r])_j
                    d.wbuf2[n + j] = s1;    // (A_r' y_r)_j
                };
                spmv_tiles<TMA, true>(p.H, p.XY, sm, ps, epi);
                ++n_h;
            }
            grid_barrier_reduce<4, true>(p.gs, st, nrm, sm.red, sm.bcast);
            had_barrier = true;
            lmax[0] = nrm[0]; lmax[1] = nrm[1]; lmax[2] = nrm[2]; lmax[3] = nrm[3];
        }
    } else if (seg == kSegCheck) {
        // wbuf2 = [P x ; A' y] (all-reduced), S.lmax = all-reduced (max) local norms
        double nrm[2] = {0.0, 0.0};
        for (int j = gtid; j < n; j += gstride) admm_dual_norms(nrm[0], nrm[1], d.wbuf2[j], p.q[j], d.wbuf2[n + j], 1.0);
        grid_barrier_reduce<2, true>(p.gs, st, nrm, sm.red, sm.bcast);
        had_barrier = true;
        const double dx = lmax[0], dz = lmax[1];
        res_prim = lmax[2];
        res_dual = nrm[0];
        const double max_prim = lmax[3];
        const double max_dual = nanmax(nrm[1], p.normQ);
        admm_stop_test(p.s, rho, dx, dz, res_prim, res_dual, max_prim, max_dual, rhorho, conv_flag);
    }

    // write the control block back.  Segments that changed a value every block must re-read next time
    // have passed at least one grid barrier, so nobody is still reading the old block.
    if (!had_barrier) {
        grid_barrier(p.gs, st);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        S.epoch = st.epoch;
        S.ii = ii; S.k = k; S.cont = cont; S.conv_flag = conv_flag;
        S.rho = rho; S.rho1 = rho1; S.rhorho = rhorho;
        S.residual = residual; S.tol = tol; S.rz = rz;
        S.res_prim = res_prim; S.res_dual = res_dual;
        S.lmax[0] = lmax[0]; S.lmax[1] = lmax[1]; S.lmax[2] = lmax[2]; S.lmax[3] = lmax[3];
        S.rho_updates = rho_updates; S.pcg_total = pcg_total; S.pcg_maxed = pcg_maxed; S.n_h = n_h; S.n_a = n_a;
    }
}

}  // namespace qpb
