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
    long long off_wrecv;             // nranks * sstride: receive buffer of the push-style reduce-scatter
    long long sstride;               // doubles between the partials of two source ranks (>= longest slice)
    int sb[kMaxPeers + 1];           // slice bounds: rank q owns [sb[q], sb[q+1]) of the n-vectors
    long long off_w2part, off_w2red; // 2n each
    long long off_lmax;              // nranks * 4
    long long off_flags;             // kMaxPeers unsigned long long
    long long off_cta;               // 2 (parity) * (grid_max * 4 + kMaxPeers * 4): CTA partials + per-GPU totals of the dots
    int grid_max;
    AdmmInfoDev *info;
    unsigned long long *dbg;         // 16 phase timers in ns (block 0 / thread 0), printed with QPB200_TIMING
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
// The kernel.  The CG vectors are NOT replicated: rank r owns slice S_r = [sb[r], sb[r+1]) of x~, r, p, s and is
// the only one to update it; z = Pl \ r (the one vector A_r and H_r gather from) and x~ (gathered once per ADMM
// iteration) are kept coherent by pushing the owner's slice into every rank's copy (all-gather by remote stores).
//
// One CG iteration (one-reduction arrangement of the PCG, see admm_kernels.cuh) costs ONE grid barrier and TWO
// system barriers:
//   A pass    t = rho A_r z                                   (rows of this rank)          | grid barrier
//   H pass    w_r = H_r [z ; t]; the epilogue PUSHES row j of w_r straight into the receive buffer of the rank that
//             owns j (push-style reduce-scatter riding on the pass: the 7/8 n remote stores are on the wire while the
//             next tiles are computed) and accumulates z . w_r over ALL rows               | system barrier + sum<1>
//             => delta = z . K z = sum_r z . w_r + sigma z . z is known to everybody without the reduced w
//   slice     w = sum_q recv[q][j] + sigma z (R LOCAL reads, rank order => deterministic), p = z + beta p,
//             s = w + beta s, x~ += alpha p, r -= alpha s, z' = Pl \ r pushed into every rank's copy,
//             gamma' = r . z', |r|^2, z' . z' on the slice                                 | system barrier + sum<3>
// (round 1: pull-style reduce-scatter -- R dependent remote loads per element, pure NVLink latency -- and three
//  system barriers per iteration: 59 + 21 of the 122 us of an iteration on 8 GPUs, profiles/r1c_dist8_cg.txt.)
// Scalars are summed in rank order from per-GPU totals that every rank holds bit-identically, so all ranks branch
// identically.
// =====================================================================================================
template <int NV>
__device__ __forceinline__ void peer_barrier_sum(const GridSync &gs, SyncState &st, const PeerDev &pd, XState &xs,
                                                 double (&v)[NV], SpmvSmem &sm, unsigned &parity) {
    block_reduce<NV, false>(v, sm.red);            // thread 0 holds the CTA partial
    sys_barrier_impl<NV>(gs, st, pd, xs, v, sm, parity);
}

template <int TMA, bool PRE>
__global__ void __launch_bounds__(kThreads, kMinCtas) admm_peer_sliced_kernel(SparseProblemDev p, PeerDev pd) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SpmvSmem &sm = *reinterpret_cast<SpmvSmem *>(smem_raw);
    __shared__ AdmmCounters ctr;
    PipeState ps;
    spmv_smem_init(sm, ps);
    SyncState st;
    st.epoch = 0;
    XState xs;
    xs.xepoch = 0;
    unsigned parity = 0;
    if (threadIdx.x == 0) {
        ctr.rho_updates = ctr.pcg_total = ctr.pcg_maxed = ctr.n_h = ctr.n_a = 0;
        ctr.res_prim = ctr.res_dual = nan("");
    }
    auto count = [&](long long &c, long long by) { if (threadIdx.x == 0) c += by; };

    const int n = p.n, m = p.m, R = pd.nranks;
    const int gtid = blockIdx.x * kThreads + threadIdx.x;
    const int gstride = gridDim.x * kThreads;
    const int s0 = pd.sb[pd.rank], s1 = pd.sb[pd.rank + 1];   // my slice
    double *const x = p.XY, *const y = p.XY + n;
    double *const xt = p.XG, *const g = p.XG + n;          // p.XG / p.UT live in the peer-visible region
    double *const u = p.UT, *const t = p.UT + n;
    double *const w2part = pd.region[pd.rank] + pd.off_w2part;
    const double *const w2red = pd.region[pd.rank] + pd.off_w2red;
    const double *const recv = pd.region[pd.rank] + pd.off_wrecv;   // [R][sstride]: row j of rank q's partial at q sstride + j - s0
    // remote views of z (= u) and x~: same offset in every region
    const long long off_u = (long long)(p.UT - pd.region[pd.rank]), off_xt = (long long)(p.XG - pd.region[pd.rank]);

    double rho = p.s.rho, rho1 = 1.0 / rho;
    double rhorho = rho;
    int conv_flag = 1;
    bool dinv_ready = false;

    unsigned long long t_last = gtimer();
    auto tick = [&](int slot) {
        if (pd.dbg && blockIdx.x == 0 && threadIdx.x == 0) {
            const unsigned long long now = gtimer();
            pd.dbg[slot] += now - t_last;
            t_last = now;
        }
    };
    // H pass with the push-style reduce-scatter in its epilogue; returns this thread's share of sum_j pair[j] * (H_r pair)_j
    auto spmv_H_push = [&](const double *pair, double &zw) {
        auto epi = [&](int j, double sum, double) {
            int q = 0;
#pragma unroll
            for (int b = 1; b < kMaxPeers; ++b) q += (b < R && j >= pd.sb[b]) ? 1 : 0;
            pd.region[q][pd.off_wrecv + (long long)pd.rank * pd.sstride + (j - pd.sb[q])] = sum;
            zw += pair[j] * sum;
        };
        spmv_tiles<TMA, false>(p.H, pair, sm, ps, epi);
        count(ctr.n_h, 1);
    };
    auto reduced_w = [&](int j) {          // sum of all ranks' partials in rank order (deterministic); local reads
        double w = 0.0;
#pragma unroll
        for (int q = 0; q < kMaxPeers; ++q)
            if (q < R) w += __ldcg(recv + (long long)q * pd.sstride + (j - s0));
        return w;
    };
    auto push_all = [&](long long off, int j, double v) {   // element j of a replicated vector, into every rank's copy
#pragma unroll
        for (int q = 0; q < kMaxPeers; ++q)
            if (q < R) pd.region[q][off + j] = v;
    };

    long long ii = 0;
    for (ii = 1; ii <= p.s.max_iter; ++ii) {
        bool changed = false;
        if (p.s.adaptive_rho && ((rhorho * p.s.rho_factor < rho) || (rhorho > p.s.rho_factor * rho))) {
            rho = rhorho;
            rho1 = 1.0 / rho;
            changed = true;
            count(ctr.rho_updates, 1);
        }
        if (changed || !dinv_ready) {
            if (PRE)
                for (int j = s0 + gtid; j < s1; j += gstride) p.dinv[j] = 1.0 / (p.dP[j] + p.s.sigma + rho * p.dAA[j]);
            if (changed)
                for (int i = gtid; i < m; i += gstride) g[i] = rho * (p.zt[i] - p.z[i]) + y[i];
            dinv_ready = true;
            grid_barrier(p.gs, st);
        }
        // ---- r0 on my slice: r = sigma (x - x~) - q - sum_q H_q [x~ ; g_q];  z = Pl \ r pushed to everyone
        double unused = 0.0;
        spmv_H_push(p.XG, unused);
        sys_barrier<true>(p.gs, st, pd, xs);
        double d3[3] = {0.0, 0.0, 0.0};                                // gamma = r.z, |r|^2, z.z on my slice
        for (int j = s0 + gtid; j < s1; j += gstride) {
            const double rj = p.s.sigma * (x[j] - xt[j]) - p.q[j] - reduced_w(j);
            p.r[j] = rj;
            const double zj = PRE ? p.dinv[j] * rj : rj;
            push_all(off_u, j, zj);
            d3[0] += rj * zj;
            d3[1] += rj * rj;
            d3[2] += zj * zj;
        }
        peer_barrier_sum<3>(p.gs, st, pd, xs, d3, sm, parity);         // also publishes the z slices
        double residual = sqrt(d3[1]);
        const double tol = fmax(p.s.pcg_rel_eps * residual, p.s.pcg_eps);
        long long k = 0;
        double gam_prev = 0.0, a_cg = 0.0;
        bool first = true;
        while (k < p.s.pcg_max_iter && !(residual <= tol)) {
            tick(5);
            {   // t = rho A_r z
                auto epi = [&](int i, double sum, double) { t[i] = rho * sum; };
                spmv_tiles<TMA, false>(p.A, u, sm, ps, epi);
                count(ctr.n_a, 1);
            }
            grid_barrier(p.gs, st);
            tick(0);
            double zw[1] = {0.0};
            spmv_H_push(p.UT, zw[0]);
            tick(1);
            peer_barrier_sum<1>(p.gs, st, pd, xs, zw, sm, parity);     // partials delivered; z . sum_r w_r known
            tick(2);
            const double gam = d3[0];
            const double delta = zw[0] + p.s.sigma * d3[2];            // z . (P + rho A'A + sigma I) z
            double beta = 0.0;
            if (first) {
                if (!(delta > 0.0)) break;                             // breakdown guard (K is SPD)
                a_cg = gam / delta;
            } else {
                beta = gam / gam_prev;
                const double den = delta - beta * gam / a_cg;
                if (!(den > 0.0)) break;
                a_cg = gam / den;
            }
            gam_prev = gam;
            d3[0] = d3[1] = d3[2] = 0.0;
            for (int j = s0 + gtid; j < s1; j += gstride) {
                const double zj = u[j];
                double pj = zj, sj = reduced_w(j) + p.s.sigma * zj;
                if (!first) {                                          // (stale p, s of the previous solve are never read)
                    pj += beta * p.zp[j];
                    sj += beta * p.c[j];
                }
                p.zp[j] = pj;
                p.c[j] = sj;
                xt[j] += a_cg * pj;
                const double rj = p.r[j] - a_cg * sj;
                p.r[j] = rj;
                const double zn = PRE ? p.dinv[j] * rj : rj;
                push_all(off_u, j, zn);
                d3[0] += rj * zn;
                d3[1] += rj * rj;
                d3[2] += zn * zn;
            }
            first = false;
            tick(3);
            peer_barrier_sum<3>(p.gs, st, pd, xs, d3, sm, parity);
            tick(4);
            residual = sqrt(d3[1]);
            ++k;
        }
        count(ctr.pcg_total, k);
        if (k >= p.s.pcg_max_iter && !(residual <= tol)) count(ctr.pcg_maxed, 1);
        // ---- all-gather x~ (every rank pushes its slice), then the row-local update
        for (int j = s0 + gtid; j < s1; j += gstride) {
            const double v = xt[j];
#pragma unroll
            for (int q = 0; q < kMaxPeers; ++q)
                if (q < R && q != pd.rank) pd.region[q][off_xt + j] = v;
        }
        sys_barrier<true>(p.gs, st, pd, xs);

        const bool do_check = (ii % p.s.check_every) == 0;
        double nrm[4] = {0.0, 0.0, 0.0, 0.0};
        {
            auto epi = [&](int i, double sum, double) {
                nrm[1] = nanmax(nrm[1], admm_row_update(p, y, g, i, sum, p.s.alpha, 1.0 - p.s.alpha, rho, rho1));
            };
            spmv_tiles<TMA, false>(p.A, xt, sm, ps, epi);
            count(ctr.n_a, 1);
        }
        for (int j = gtid; j < n; j += gstride)
            nrm[0] = nanmax(nrm[0], admm_x_relax(x, xt, j, p.s.alpha, 1.0 - p.s.alpha));
        grid_barrier(p.gs, st);
        if (do_check) {
            {
                auto epi = [&](int i, double sum, double) { admm_prim_norms(nrm[2], nrm[3], sum, p.z[i], 1.0); };
                spmv_tiles<TMA, false>(p.A, x, sm, ps, epi);
                count(ctr.n_a, 1);
            }
            {
                auto epi = [&](int j, double sum0, double sum1) {
                    w2part[j] = sum0;
                    w2part[n + j] = sum1;
                };
                spmv_tiles<TMA, true>(p.H, p.XY, sm, ps, epi);
                count(ctr.n_h, 1);
            }
            grid_barrier_reduce<4, true>(p.gs, st, nrm, sm.red, sm.bcast);
            peer_allmax4(p.gs, st, pd, xs, nrm);
            peer_allreduce(p.gs, st, pd, xs, pd.off_w2part, pd.off_w2red, 2 * n);
            double nd[2] = {0.0, 0.0};
            for (int j = gtid; j < n; j += gstride) admm_dual_norms(nd[0], nd[1], w2red[j], p.q[j], w2red[n + j], 1.0);
            grid_barrier_reduce<2, true>(p.gs, st, nd, sm.red, sm.bcast);
            const double dx = nrm[0], dz = nrm[1];
            const double res_prim = nrm[2], res_dual = nd[0];
            if (threadIdx.x == 0) { ctr.res_prim = res_prim; ctr.res_dual = res_dual; }
            const double max_prim = nrm[3];
            const double max_dual = nanmax(nd[1], p.normQ);
            admm_stop_test(p.s, rho, dx, dz, res_prim, res_dual, max_prim, max_dual, rhorho, conv_flag);
            if (conv_flag != 1) break;
        }
    }
    if (ii > p.s.max_iter) ii = p.s.max_iter;

    if (blockIdx.x == 0 && threadIdx.x == 0) {
        AdmmInfoDev &o = *pd.info;
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
