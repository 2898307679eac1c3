// peer_kernels.cuh -- ONE persistent kernel per GPU for a QP row-partitioned over R GPUs: the SpMV passes,
// the vector updates AND the all-reduces run inside the same cooperative launch; the collectives are
// plain loads/stores on NVLink peer memory (cudaIpc-mapped), with no NCCL call and no host round trip in
// the loop.  (dist_kernels.cuh + ncclAllReduce is the baseline this replaces.)
//
// Operator split (same as dist_kernels.cuh): rank r owns rows I_r of A and columns J_r of P,
//     K u = sum_r H_r [u ; rho A_r u] + sigma u,   H_r = [P[:, J_r]  A_r'].
// In-kernel all-reduce of an n-vector ("two-shot", deterministic):
//   1. every rank writes its partial H_r v into its own peer-visible buffer      -> system barrier
//   2. rank r sums slice r of all R partials in rank order (remote reads over NVLink) and PUSHES the sums
//      into every rank's result buffer (remote writes)                                 -> system barrier
// Every element is reduced by exactly one rank, so all ranks receive bit-identical vectors and branch
// identically.  The system barrier is the grid barrier whose last-arriving CTA exchanges an epoch flag with
// the peers (st.release.sys / ld.acquire.sys on peer memory) before releasing the local CTAs.
// One process per GPU launches its kernel at the same time; kernels on DIFFERENT GPUs spin on each other,
// never kernels sharing a GPU.
#pragma once
#include "dist_kernels.cuh"

namespace qpb {

constexpr int kMaxPeers = 8;

struct PeerDev {
    int rank, nranks;
    double *region[kMaxPeers];       // peer-visible region of every rank (own entry = local pointer)
    // offsets in doubles inside a region
    long long off_wpart, off_wred;   // n each
    long long off_w2part, off_w2red; // 2n each
    long long off_lmax;              // nranks * 4
    long long off_flags;             // kMaxPeers unsigned long long
    long long off_cta;               // 2 (parity) * (grid_max * 4 + kMaxPeers * 4): CTA partials + per-GPU totals of the dots
    int grid_max;
    AdmmInfoDev *info;
    unsigned long long *dbg;         // 16 phase timers in ns (block 0 / thread 0), printed with QPB200_TIMING
    double *wslice;                  // n doubles (local): w = K z of the Chronopoulos-Gear variant
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct XState {
    unsigned long long xepoch;
};

// grid barrier + cross-GPU barrier in one: only the last-arriving CTA talks to the peers, and it does so with
// one lane per peer (signal + poll in parallel: the handshake costs one NVLink round trip for any R).
// Memory ordering: every CTA releases at gpu scope (that waits for its outstanding stores, local or peer, to be
// performed); system-scope visibility is established once, cumulatively, by the last arriver's
// fence.acq_rel.sys before it raises the peers' flags.  (A system-scope fence in every CTA costs +7.7 us per
// barrier at 592 CTAs: measured.)
// NV > 0: a sum all-reduce of NV doubles rides on the barrier.  Every CTA deposits its partial before arriving;
// the last arriver adds the G partials in a fixed order (deterministic whoever arrives last), pushes the GPU
// total into slot `rank` of every rank's table, and after the barrier everybody adds the R totals in rank
// order -> bit-identical results on all ranks, ONE barrier for a cross-GPU dot product.
template <int NV>
__device__ __forceinline__ void sys_barrier_impl(const GridSync &gs, SyncState &st, const PeerDev &pd, XState &xs,
                                                 double *v, SpmvSmem &sm, unsigned &parity) {
    static_assert(NV <= 4, "slot width");
    double *loc = nullptr, *tot = nullptr;
    if (NV > 0) {
        parity ^= 1u;
        const long long per = (long long)pd.grid_max * 4 + kMaxPeers * 4;
        loc = pd.region[pd.rank] + pd.off_cta + (long long)parity * per;   // [grid_max][4] CTA partials
        tot = loc + (long long)pd.grid_max * 4;                            // [kMaxPeers][4] GPU totals
        if (threadIdx.x == 0)
            for (int i = 0; i < NV; ++i) loc[blockIdx.x * 4 + i] = v[i];
    }
    st.epoch += 1;
    xs.xepoch += 1;
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        unsigned long long old = 0;
        if (lane == 0) {
            fence_acq_rel_gpu();
            old = atomicAdd(gs.count, 1ULL);
        }
        old = __shfl_sync(0xffffffffu, old, 0);
        if (old == st.epoch * gridDim.x - 1ULL) {
            if (lane == 0) fence_acq_rel_gpu();     // one fence per warp: a fence executed by 32 lanes is 32 fences
            __syncwarp();
            if (NV > 0) {
                for (int i = 0; i < NV; ++i) {
                    // this warp is alone on the critical path: keep 8 independent L2 loads in flight per lane
                    double x = 0.0;
                    for (unsigned jb = lane; jb < gridDim.x; jb += 32 * 8) {
                        double tl[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const unsigned j = jb + 32 * e;
                            tl[e] = j < gridDim.x ? __ldcg(loc + j * 4 + i) : 0.0;
                        }
#pragma unroll
                        for (int e = 0; e < 8; ++e) x += tl[e];
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
                    x = __shfl_sync(0xffffffffu, x, 0);
                    if (lane < pd.nranks) {
                        const long long off = (long long)(tot - pd.region[pd.rank]) + pd.rank * 4 + i;
                        pd.region[lane][off] = x;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) asm volatile("fence.acq_rel.sys;" ::: "memory");
            __syncwarp();
            if (lane < pd.nranks && lane != pd.rank) {
                st_release_sys(reinterpret_cast<unsigned long long *>(pd.region[lane] + pd.off_flags) + pd.rank, xs.xepoch);
                const unsigned long long *mine = reinterpret_cast<const unsigned long long *>(pd.region[pd.rank] + pd.off_flags) + lane;
                const long long t0 = clock64();
                while (ld_acquire_sys(mine) < xs.xepoch) {
                    if (clock64() - t0 > 60000000000LL) __trap();   // ~30 s: a peer is gone; fail loudly, do not hang
                }
            }
            __syncwarp();
            if (lane == 0) st_release_gpu(gs.flag, st.epoch);
        } else if (lane == 0) {
            while (ld_acquire_gpu(gs.flag) < st.epoch) {
            }
        }
        if (lane == 0) fence_acq_rel_gpu();   // acquire + L1 invalidate for the phase that follows
    }
    __syncthreads();
    if (NV > 0) {
        for (int i = 0; i < NV; ++i) {
            double x = 0.0;
            for (int q = 0; q < pd.nranks; ++q) x += __ldcg(tot + q * 4 + i);
            v[i] = x;
        }
    }
}

template <bool REMOTE_WRITES>
__device__ __forceinline__ void sys_barrier(const GridSync &gs, SyncState &st, const PeerDev &pd, XState &xs) {
    SpmvSmem *none = nullptr;
    unsigned dummy = 0;
    sys_barrier_impl<0>(gs, st, pd, xs, nullptr, *none, dummy);
}

// all-reduce(sum) of `len` doubles: partial at off_part in every region -> result at off_red in every region
__device__ __forceinline__ void peer_allreduce(const GridSync &gs, SyncState &st, const PeerDev &pd, XState &xs,
                                               long long off_part, long long off_red, int len) {
    sys_barrier<false>(gs, st, pd, xs);               // all partials written (locally) and visible
    const int R = pd.nranks;
    const int j0 = (int)((long long)len * pd.rank / R), j1 = (int)((long long)len * (pd.rank + 1) / R);
    const double *src[kMaxPeers];
    double *dst[kMaxPeers];
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q) {
        src[q] = pd.region[q < R ? q : 0] + off_part;
        dst[q] = pd.region[q < R ? q : 0] + off_red;
    }
    constexpr int U = 4;     // independent remote loads in flight per thread
    const int stride = gridDim.x * kThreads;
    for (int jb = j0 + blockIdx.x * kThreads + threadIdx.x; jb < j1; jb += U * stride) {
        double s[U];
#pragma unroll
        for (int e = 0; e < U; ++e) s[e] = 0.0;
#pragma unroll
        for (int q = 0; q < kMaxPeers; ++q)          // rank order: deterministic
            if (q < R) {
#pragma unroll
                for (int e = 0; e < U; ++e) {
                    const int j = jb + e * stride;
                    if (j < j1) s[e] += __ldcg(src[q] + j);
                }
            }
#pragma unroll
        for (int q = 0; q < kMaxPeers; ++q)          // push to everyone (incl. self)
            if (q < R) {
#pragma unroll
                for (int e = 0; e < U; ++e) {
                    const int j = jb + e * stride;
                    if (j < j1) dst[q][j] = s[e];
                }
            }
    }
    sys_barrier<true>(gs, st, pd, xs);                // all slices delivered everywhere
}

// all-reduce(max) of 4 doubles per rank: every rank pushes its 4 values into slot `rank` of every region
__device__ __forceinline__ void peer_allmax4(const GridSync &gs, SyncState &st, const PeerDev &pd, XState &xs, double (&v)[4]) {
    if (blockIdx.x == 0 && threadIdx.x < 4)
        for (int q = 0; q < pd.nranks; ++q) pd.region[q][pd.off_lmax + 4 * pd.rank + threadIdx.x] = v[threadIdx.x];
    sys_barrier<true>(gs, st, pd, xs);
    const double *slots = pd.region[pd.rank] + pd.off_lmax;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double x = 0.0;
        for (int q = 0; q < pd.nranks; ++q) x = nanmax(x, __ldcg(slots + 4 * q + i));
        v[i] = x;
    }
}

// =====================================================================================================
// Sliced variant: the CG vectors are NOT replicated.  Rank r owns slice S_r = [n r / R, n (r+1) / R) of
// x~, r, z, c and is the only one to update it; u (gathered by A_r and H_r) and x~ (gathered once per ADMM
// iteration) are kept coherent by pushing the owner's slice into every rank's copy (all-gather by remote
// stores).  The reduce-scatter of H_r [u ; rho A_r u] is fused with c = w + sigma u and the u.c partial sums;
// dot products are reduced across GPUs by pushing every CTA's partial into every rank's slot table and
// summing the R x G partials in the same fixed order everywhere after ONE system barrier.
// Per CG iteration: 1 grid barrier + 4 system barriers, vector traffic 1/R of the replicated variant.
// =====================================================================================================
template <int NV>
__device__ __forceinline__ void peer_barrier_sum(const GridSync &gs, SyncState &st, const PeerDev &pd, XState &xs,
                                                 double (&v)[NV], SpmvSmem &sm, unsigned &parity) {
    block_reduce<NV, false>(v, sm.red);            // thread 0 holds the CTA partial
    sys_barrier_impl<NV>(gs, st, pd, xs, v, sm, parity);
}

// CGV = true: Chronopoulos-Gear arrangement of the same PCG (one fused reduction per iteration, no all-gather of
// the search direction): see the loop below.
template <int TMA, bool PRE, bool CGV>
__global__ void __launch_bounds__(kThreads, kMinCtas) admm_peer_sliced_kernel(SparseProblemDev p, PeerDev pd) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SpmvSmem &sm = *reinterpret_cast<SpmvSmem *>(smem_raw);
    PipeState ps;
    spmv_smem_init(sm, ps);
    SyncState st;
    st.epoch = 0;
    XState xs;
    xs.xepoch = 0;
    unsigned parity = 0;

    const int n = p.n, m = p.m, R = pd.nranks;
    const int gtid = blockIdx.x * kThreads + threadIdx.x;
    const int gstride = gridDim.x * kThreads;
    const int s0 = (int)((long long)n * pd.rank / R), s1 = (int)((long long)n * (pd.rank + 1) / R);   // my slice
    double *const x = p.XY, *const y = p.XY + n;
    double *const xt = p.XG, *const g = p.XG + n;          // p.XG / p.UT live in the peer-visible region
    double *const u = p.UT, *const t = p.UT + n;
    double *const zpv = PRE ? p.zp : p.r;
    double *const wpart = pd.region[pd.rank] + pd.off_wpart;
    double *const w2part = pd.region[pd.rank] + pd.off_w2part;
    const double *const w2red = pd.region[pd.rank] + pd.off_w2red;
    // remote views of u and x~ (same offset in every region)
    const long long off_u = (long long)(p.UT - pd.region[pd.rank]), off_xt = (long long)(p.XG - pd.region[pd.rank]);
    const double *wsrc[kMaxPeers];
    double *udst[kMaxPeers], *xtdst[kMaxPeers];
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q) {
        wsrc[q] = pd.region[q < R ? q : 0] + pd.off_wpart;
        udst[q] = pd.region[q < R ? q : 0] + off_u;
        xtdst[q] = pd.region[q < R ? q : 0] + off_xt;
    }

    double rho = p.s.rho, rho1 = 1.0 / rho;
    const double alpha = p.s.alpha, alpha1 = 1.0 - alpha;
    const double sigma = p.s.sigma;
    const double eps_admm = fmin(p.s.eps_abs, p.s.eps_rel) * 1e-2;
    double rhorho = rho;
    int conv_flag = 1;
    long long rho_updates = 0, pcg_total = 0, pcg_maxed = 0, n_h = 0, n_a = 0;
    double res_prim = nan(""), res_dual = nan("");
    bool dinv_ready = false;

    unsigned long long t_last = gtimer();
    auto tick = [&](int slot) {
        if (pd.dbg && blockIdx.x == 0 && threadIdx.x == 0) {
            const unsigned long long now = gtimer();
            pd.dbg[slot] += now - t_last;
            t_last = now;
        }
    };
    auto spmv_A_t = [&]() {
        auto epi = [&](int i, double s0_, double) { t[i] = rho * s0_; };
        spmv_tiles<TMA, false>(p.A, u, sm, ps, epi);
        ++n_a;
    };
    auto spmv_H_partial = [&](const double *pair) {
        auto epi = [&](int j, double s0_, double) { wpart[j] = s0_; };
        spmv_tiles<TMA, false>(p.H, pair, sm, ps, epi);
        ++n_h;
    };
    auto reduced_w = [&](int j) {          // sum of all ranks' partials in rank order (deterministic)
        double w = 0.0;
#pragma unroll
        for (int q = 0; q < kMaxPeers; ++q)
            if (q < R) w += __ldcg(wsrc[q] + j);
        return w;
    };

    long long ii = 0;
    for (ii = 1; ii <= p.s.max_iter; ++ii) {
        bool changed = false;
        if (p.s.adaptive_rho && ((rhorho * p.s.rho_factor < rho) || (rhorho > p.s.rho_factor * rho))) {
            rho = rhorho;
            rho1 = 1.0 / rho;
            changed = true;
            ++rho_updates;
        }
        if (changed || !dinv_ready) {
            if (PRE)
                for (int j = s0 + gtid; j < s1; j += gstride) p.dinv[j] = 1.0 / (p.dP[j] + sigma + rho * p.dAA[j]);
            if (changed)
                for (int i = gtid; i < m; i += gstride) g[i] = rho * (p.zt[i] - p.z[i]) + y[i];
            dinv_ready = true;
            grid_barrier(p.gs, st);
        }
        // ---- r0 on my slice: r = sigma (x - x~) - q - sum_q H_q [x~ ; g_q];  u = z = Pl \ r, pushed to everyone
        spmv_H_partial(p.XG);
        sys_barrier<false>(p.gs, st, pd, xs);
        double acc[2] = {0.0, 0.0};
        for (int j = s0 + gtid; j < s1; j += gstride) {
            const double rj = sigma * (x[j] - xt[j]) - p.q[j] - reduced_w(j);
            p.r[j] = rj;
            const double zj = PRE ? p.dinv[j] * rj : rj;
            if (PRE && !CGV) p.zp[j] = zj;
#pragma unroll
            for (int q = 0; q < kMaxPeers; ++q)
                if (q < R) udst[q][j] = zj;
            acc[0] += rj * rj;
            acc[1] += rj * zj;
        }
        double residual, tol;
        long long k = 0;
        if (CGV) {
            // ---- Chronopoulos-Gear PCG: u holds z = Pl \ r (the only gathered CG vector), p.zp holds the search
            //      direction, p.c holds s = K p, pd.wslice holds w = K z; gamma = r.z, delta = z.w and |r|^2 come out of
            //      ONE reduction per iteration.  Same iterates as the standard recurrence in exact arithmetic; one
            //      extra operator application per solve (the w of the converged residual is not used).
            sys_barrier<true>(p.gs, st, pd, xs);                       // z slices published
            double gam = 0.0, a_cg = 0.0;
            bool first = true;
            for (;;) {
                tick(6);
                spmv_A_t();
                grid_barrier(p.gs, st);
                tick(0);
                spmv_H_partial(p.UT);
                tick(1);
                sys_barrier<false>(p.gs, st, pd, xs);
                double d3[3] = {0.0, 0.0, 0.0};                        // r.z, z.w, r.r on my slice
                for (int j = s0 + gtid; j < s1; j += gstride) {
                    const double zj = u[j], rj = p.r[j];
                    const double wj = reduced_w(j) + sigma * zj;
                    pd.wslice[j] = wj;
                    d3[0] += rj * zj;
                    d3[1] += zj * wj;
                    d3[2] += rj * rj;
                }
                peer_barrier_sum<3>(p.gs, st, pd, xs, d3, sm, parity);
                tick(2);
                residual = sqrt(d3[2]);
                if (first) tol = fmax(p.s.pcg_rel_eps * residual, p.s.pcg_eps);
                if (!first) ++k;
                if (!(k < p.s.pcg_max_iter && !(residual <= tol))) break;
                double beta = 0.0;
                if (first) {
                    a_cg = d3[0] / d3[1];
                } else {
                    beta = d3[0] / gam;
                    const double den = d3[1] - beta * d3[0] / a_cg;
                    if (!(den > 0.0)) break;
                    a_cg = d3[0] / den;
                }
                if (first && !(d3[1] > 0.0)) break;
                gam = d3[0];
                first = false;
                // p = z + beta p ; s = w + beta s ; x~ += a p ; r -= a s ; z = Pl \ r -> pushed to everyone
                for (int j = s0 + gtid; j < s1; j += gstride) {
                    const double pj = u[j] + beta * p.zp[j];
                    const double sj = pd.wslice[j] + beta * p.c[j];
                    p.zp[j] = pj;
                    p.c[j] = sj;
                    xt[j] += a_cg * pj;
                    const double rj = p.r[j] - a_cg * sj;
                    p.r[j] = rj;
                    const double zj = PRE ? p.dinv[j] * rj : rj;
#pragma unroll
                    for (int q = 0; q < kMaxPeers; ++q)
                        if (q < R) udst[q][j] = zj;
                }
                sys_barrier<true>(p.gs, st, pd, xs);
                tick(4);
            }
        } else {
        peer_barrier_sum<2>(p.gs, st, pd, xs, acc, sm, parity);      // also publishes the u slices
        residual = sqrt(acc[0]);
        double rz = acc[1];
        tol = fmax(p.s.pcg_rel_eps * residual, p.s.pcg_eps);
        while (k < p.s.pcg_max_iter && !(residual <= tol)) {
            tick(6);
            spmv_A_t();
            grid_barrier(p.gs, st);
            tick(0);
            spmv_H_partial(p.UT);
            tick(1);
            sys_barrier<false>(p.gs, st, pd, xs);
            // reduce-scatter fused with c = w + sigma u and u.c (my slice only)
            double uc[1] = {0.0};
            for (int j = s0 + gtid; j < s1; j += gstride) {
                const double uj = u[j];
                const double cj = reduced_w(j) + sigma * uj;
                p.c[j] = cj;
                uc[0] += uj * cj;
            }
            peer_barrier_sum<1>(p.gs, st, pd, xs, uc, sm, parity);
            tick(2);
            if (!(uc[0] > 0.0)) break;
            const double a_cg = rz / uc[0];
            double acc2[2] = {0.0, 0.0};
            for (int j = s0 + gtid; j < s1; j += gstride) {
                xt[j] += a_cg * u[j];
                const double rj = p.r[j] - a_cg * p.c[j];
                p.r[j] = rj;
                const double zj = PRE ? p.dinv[j] * rj : rj;
                if (PRE) p.zp[j] = zj;
                acc2[0] += rj * rj;
                acc2[1] += rj * zj;
            }
            peer_barrier_sum<2>(p.gs, st, pd, xs, acc2, sm, parity);
            residual = sqrt(acc2[0]);
            const double rz_new = acc2[1];
            ++k;
            tick(4);
            if (k < p.s.pcg_max_iter && !(residual <= tol)) {
                const double beta = rz_new / rz;
                for (int j = s0 + gtid; j < s1; j += gstride) {
                    const double un = zpv[j] + beta * u[j];
#pragma unroll
                    for (int q = 0; q < kMaxPeers; ++q)
                        if (q < R) udst[q][j] = un;
                }
                sys_barrier<true>(p.gs, st, pd, xs);
            }
            rz = rz_new;
            tick(5);
        }
        }
        pcg_total += k;
        if (k >= p.s.pcg_max_iter && !(residual <= tol)) ++pcg_maxed;
        // ---- all-gather x~ (every rank pushes its slice), then the row-local update
        for (int j = s0 + gtid; j < s1; j += gstride) {
            const double v = xt[j];
#pragma unroll
            for (int q = 0; q < kMaxPeers; ++q)
                if (q < R && q != pd.rank) xtdst[q][j] = v;
        }
        sys_barrier<true>(p.gs, st, pd, xs);

        const bool do_check = (ii % p.s.check_every) == 0;
        double nrm[4] = {0.0, 0.0, 0.0, 0.0};
        {
            auto epi = [&](int i, double s0_, double) {
                const double zt_i = s0_;
                const double z_old = p.z[i], y_old = y[i];
                const double zr = alpha * zt_i + alpha1 * z_old;
                const double z_new = clamp_julia(zr + rho1 * y_old, p.l[i], p.u[i]);
                const double y_new = y_old + rho * (zr - z_new);
                p.z[i] = z_new;
                y[i] = y_new;
                p.zt[i] = zt_i;
                g[i] = rho * (zt_i - z_new) + y_new;
                nrm[1] = nanmax(nrm[1], fabs(z_new - z_old));
            };
            spmv_tiles<TMA, false>(p.A, xt, sm, ps, epi);
            ++n_a;
        }
        for (int j = gtid; j < n; j += gstride) {
            const double x_old = x[j];
            const double x_new = alpha * xt[j] + alpha1 * x_old;
            x[j] = x_new;
            nrm[0] = nanmax(nrm[0], fabs(x_new - x_old));
        }
        grid_barrier(p.gs, st);
        if (do_check) {
            {
                auto epi = [&](int i, double s0_, double) {
                    const double zi = p.z[i];
                    nrm[2] = nanmax(nrm[2], fabs(s0_ - zi));
                    nrm[3] = nanmax(nrm[3], fabs(s0_));
                    nrm[3] = nanmax(nrm[3], fabs(zi));
                };
                spmv_tiles<TMA, false>(p.A, x, sm, ps, epi);
                ++n_a;
            }
            {
                auto epi = [&](int j, double s0_, double s1_) {
                    w2part[j] = s0_;
                    w2part[n + j] = s1_;
                };
                spmv_tiles<TMA, true>(p.H, p.XY, sm, ps, epi);
                ++n_h;
            }
            grid_barrier_reduce<4, true>(p.gs, st, nrm, sm.red, sm.bcast);
            peer_allmax4(p.gs, st, pd, xs, nrm);
            peer_allreduce(p.gs, st, pd, xs, pd.off_w2part, pd.off_w2red, 2 * n);
            double nd[2] = {0.0, 0.0};
            for (int j = gtid; j < n; j += gstride) {
                const double px = w2red[j], aty = w2red[n + j];
                nd[0] = nanmax(nd[0], fabs(px + p.q[j] + aty));
                nd[1] = nanmax(nd[1], fabs(px));
                nd[1] = nanmax(nd[1], fabs(aty));
            }
            grid_barrier_reduce<2, true>(p.gs, st, nd, sm.red, sm.bcast);
            const double dx = nrm[0], dz = nrm[1];
            res_prim = nrm[2];
            res_dual = nd[0];
            const double max_prim = nrm[3];
            const double max_dual = nanmax(nd[1], p.normQ);
            if (p.s.adaptive_rho) {
                const double num = res_prim * max_dual, den = res_dual * max_prim;
                rhorho = clamp_julia(rho * sqrt(num / den), 1e-3, 1e6);
            }
            if ((res_prim < p.s.eps_abs + p.s.eps_rel * max_prim) && (res_dual < p.s.eps_abs + p.s.eps_rel * max_dual)) conv_flag = 3;
            if ((dx <= eps_admm) && (dz <= eps_admm)) conv_flag = 2;
            if (conv_flag != 1) break;
        }
    }
    if (ii > p.s.max_iter) ii = p.s.max_iter;

    if (blockIdx.x == 0 && threadIdx.x == 0) {
        AdmmInfoDev &o = *pd.info;
        o.conv_flag = conv_flag;
        o.iterations = ii;
        o.rho_final = rho;
        o.res_prim = res_prim;
        o.res_dual = res_dual;
        o.rho_updates = rho_updates;
        o.pcg_iters_total = pcg_total;
        o.pcg_maxed = pcg_maxed;
        o.n_h_passes = n_h;
        o.n_a_passes = n_a;
    }
}

}  // namespace qpb
