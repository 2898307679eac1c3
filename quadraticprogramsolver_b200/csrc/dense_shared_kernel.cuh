// dense_shared_kernel.cuh -- a batch of small dense QPs that SHARE P and A (MPC-style: one plant model, many initial
// states / references: only q, l, u differ), SURVEY.md 8(f) row 3: "one factor, many right-hand sides -> turns
// cfg3's matrix-vector products into tensor-core GEMMs".
//
// Replaces SolveQuadraticProgram! (SolveQuadraticProgram.jl:14-112) with a direct plugin (LinearSystemSolvers.jl:16-107,
// reduced form K x~ = sigma x - q + A'(rho z - y), z~ = A x~) for every problem of the batch, rho fixed (adptRho = false,
// the reference's default: with one rho there is one K = P + sigma I + rho A'A for the whole batch).
//
// B200 design: K^-1 is formed ONCE (dense_shared_factor_kernel: the DMMA SYRK + blocked Cholesky + in-place inverse
// of dense_kernel.cuh on one CTA).  A CTA of 8 warps then keeps A, K^-1 and P in shared memory and iterates a TILE
// of 16 problems as the 16 columns of three FP64 tensor-pipe GEMMs per ADMM iteration (mma.sync.m8n8k4.f64 -> DMMA):
//     RHS[64 x 16] = A'[64 x mp] W[mp x 16]       then  rhs = sigma x - q + .        (LinearSystemSolvers.jl:37-38)
//     X~ [64 x 16] = K^-1[64 x 64] RHS[64 x 16]   then  x = alpha x~ + (1 - alpha) x (SolveQuadraticProgram.jl:57)
//     Z~ [mp x 16] = A[mp x 64] X~[64 x 16]       then  z, y update, w = rho z - y   (:59-61)
// The iterates x, q, z, y, l, u live in REGISTERS in the accumulator-fragment layout of the GEMM that produces them
// (the element-wise updates are GEMM epilogues); only the B operands (W, RHS, X~) pass through shared memory.
// Columns converge independently: every check_every iterations CheckConvergence (:79-112) runs for all 16 columns
// as three more GEMMs (A X, A' Y, P X) with column-wise max reductions; a finished column is written out and its
// slot refilled from the batch's work queue, so the tile never idles behind its slowest problem.
// Bit-reproducible (fixed tile ownership and reduction order); results agree with the per-problem-factor kernel of
// dense_kernel.cuh to rounding (different summation order inside the products).
#pragma once
#include "dense_kernel.cuh"

namespace qpb {

constexpr int kShThreads = 256;
constexpr int kShWarps = 8;
constexpr int kShNc = 16;        // problems (columns) per tile
constexpr int kShLda = 68;       // leading dimension of the A-operand matrices (== 4 mod 16: conflict-free fragments)
constexpr int kShLdb = 20;       // leading dimension of the B-operand tiles  (== 4 mod 16)

struct DenseSharedParams {
    int batch, n, m, mp;         // mp = m rounded up to a multiple of 8 (<= 128)
    const double *P, *A;         // ONE n x n and ONE m x n matrix, column-major
    const double *Kinv;          // 64 x 64 row-major, written by dense_shared_factor_kernel
    const double *q, *l, *u;     // batch x n, batch x m
    double *X;                   // batch x n, start points in, solutions out
    int *flags;
    long long *iters;
    unsigned long long *totals;  // [0] ADMM iterations summed over the batch
    unsigned int *queue;         // next problem index
    AdmmSettingsDev s;
};

static size_t dense_shared_smem_bytes(int mp) {
    // A (mp x 68), K^-1 and P (64 x 68 each), W and Y (mp x 20 each), RHS and X~ (64 x 20 each), reduction scratch
    return sizeof(double) * ((size_t)mp * kShLda + 2 * kDN * kShLda + 2 * (size_t)mp * kShLdb + 2 * kDN * kShLdb +
                             (16 + 24) * 8 * 3 + 4 * kShNc) + sizeof(int) * 8 * kShNc;
}

// K^-1 for the whole batch: one CTA of kDThreads threads, the factorisation code of dense_kernel.cuh.
__global__ void __launch_bounds__(kDThreads) dense_shared_factor_kernel(const double *Pg, const double *Ag, int n, int m, int mp4,
                                                                        double rho, double sigma, double *Kinv, int *fail) {
    extern __shared__ __align__(16) unsigned char raw[];
    const DenseSmem sm = carve(raw, mp4);
    const int lda = mp4 + 2;
    for (int idx = threadIdx.x; idx < mp4 * kDN; idx += kDThreads) {
        const int i = idx % mp4, j = idx / mp4;
        sm.As[i + lda * j] = (i < m && j < n) ? Ag[i + (size_t)m * j] : 0.0;
    }
    __syncthreads();
    build_K(sm, mp4, Pg, n, rho, sigma);
    const bool ok = chol_blocked(sm);
    trtri_lower(sm);
    lauum_lower_and_mirror(sm);
    __syncthreads();
    for (int e = threadIdx.x; e < kDN * kDN; e += kDThreads) Kinv[e] = sm.Lp[pidx(e / kDN, e % kDN)];
    if (threadIdx.x == 0 && !ok) *fail = 1;
}

// max over the 8 lanes that hold the same columns of an accumulator fragment (they differ in lane >> 2)
__device__ __forceinline__ double frag_colmax(double v) {
    v = nanmax(v, __shfl_xor_sync(0xffffffffu, v, 4));
    v = nanmax(v, __shfl_xor_sync(0xffffffffu, v, 8));
    v = nanmax(v, __shfl_xor_sync(0xffffffffu, v, 16));
    return v;
}

__global__ void __launch_bounds__(kShThreads, 1) dense_shared_kernel(DenseSharedParams p) {
    extern __shared__ __align__(16) unsigned char raw[];
    const int n = p.n, m = p.m, mp = p.mp;
    double *Arm = reinterpret_cast<double *>(raw);            // A row-major [mp][68]
    double *Kin = Arm + (size_t)mp * kShLda;                   // K^-1 [64][68]
    double *Prm = Kin + kDN * kShLda;                          // P    [64][68]
    double *Wsm = Prm + kDN * kShLda;                          // W    [mp][20]
    double *Ysm = Wsm + (size_t)mp * kShLdb;                   // Y    [mp][20]   (checks only)
    double *Rsm = Ysm + (size_t)mp * kShLdb;                   // RHS  [64][20]   (X at the checks)
    double *Xtm = Rsm + kDN * kShLdb;                          // X~   [64][20]
    double *red1 = Xtm + kDN * kShLdb;                         // [16 tiles][8 cols][3]: dx, rd, max(|Px|,|A'y|)
    double *red3 = red1 + 16 * 8 * 3;                          // [24 tiles][8 cols][3]: dz, rp, max(|Ax|,|z|)
    double *normQ = red3 + 24 * 8 * 3;                         // [16]
    double *colres = normQ + kShNc;                            // [16][3] spare
    int *colb = reinterpret_cast<int *>(colres + 3 * kShNc);   // problem in each column, -1 = empty
    int *colstart = colb + kShNc;                              // iteration at which the column was (re)filled
    int *colflag = colstart + kShNc;                           // 0 = running, else the ConvergenceFlag it finished with
    int *colnew = colflag + kShNc;                             // problem to load into the column at this refill, -2 = keep
    int *ctl = colnew + kShNc;                                 // [0] live columns, [1] next iteration cap
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g8 = lane >> 2, t4 = lane & 3;
    const double alpha = p.s.alpha, alpha1 = 1.0 - alpha, sigma = p.s.sigma, rho = p.s.rho, rho1 = 1.0 / rho;
    const double eps_admm = fmin(p.s.eps_abs, p.s.eps_rel) * 1e-2;
    const int ksA = mp / 4;

    // ---- the shared operands
    for (int e = tid; e < mp * kDN; e += kShThreads) {
        const int i = e % mp, j = e / mp;
        Arm[i * kShLda + j] = (i < m && j < n) ? __ldg(p.A + i + (size_t)m * j) : 0.0;
    }
    for (int e = tid; e < kDN * kDN; e += kShThreads) {
        const int i = e % kDN, j = e / kDN;
        Prm[i * kShLda + j] = (i < n && j < n) ? __ldg(p.P + i + (size_t)n * j) : 0.0;
        Kin[(e / kDN) * kShLda + (e % kDN)] = __ldg(p.Kinv + e);
    }
    if (tid < kShNc) { colb[tid] = -1; colstart[tid] = 0; colflag[tid] = 0; colnew[tid] = -1; normQ[tid] = 0.0; }
    if (tid == 0) { ctl[0] = 0; ctl[1] = 0x7fffffff; }

    // ---- register-resident iterates in accumulator-fragment layout
    // type-1 tiles (64 x 16 results): warp w owns rows 8w.. of both column halves: element (8w + g8, 8 nt + 2 t4 + e)
    // type-3 tiles (mp x 16 results): 2 (mp/8) tiles dealt round-robin, tile T -> (mt = T >> 1, nt = T & 1)
    double x[2][2], qv[2][2], dxv[2][2];
    double z[3][2], y[3][2], lo[3][2], hi[3][2], dzv[3][2];
    const int ntiles3 = 2 * (mp / 8);
    const int row1 = 8 * warp + g8;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int e = 0; e < 2; ++e) x[a][e] = qv[a][e] = dxv[a][e] = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int e = 0; e < 2; ++e) z[a][e] = y[a][e] = lo[a][e] = hi[a][e] = dzv[a][e] = 0.0;
    for (int e = tid; e < mp * kShLdb; e += kShThreads) Wsm[e] = 0.0;

    unsigned long long tot_iters = 0;
    int g = 0;                              // iterations this CTA has run
    bool refill = true;                     // the first pass only fills the columns
    for (;;) {
        __syncthreads();
        if (refill) {
            // ---- columns: retire the finished ones, draw new problems from the queue
            if (tid < kShNc) {
                const int c = tid;
                int nb = -2;                                    // keep
                if (colb[c] < 0 || colflag[c] != 0) {
                    if (colb[c] >= 0) {
                        p.flags[colb[c]] = colflag[c];
                        p.iters[colb[c]] = (long long)colres[3 * c];
                    }
                    const unsigned int b = atomicAdd(p.queue, 1u);
                    nb = b < (unsigned int)p.batch ? (int)b : -1;
                }
                colnew[c] = nb;
            }
            __syncthreads();
            // load the new problems into the fragments
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = 8 * a + 2 * t4 + e;
                    const int nb = colnew[c];
                    if (nb == -2) continue;
                    x[a][e] = (nb >= 0 && row1 < n) ? p.X[(size_t)nb * n + row1] : 0.0;   // (the old x went out when it finished)
                    qv[a][e] = (nb >= 0 && row1 < n) ? p.q[(size_t)nb * n + row1] : 0.0;
                }
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int T = warp + kShWarps * a;
                if (T >= ntiles3) continue;
                const int r = 8 * (T >> 1) + g8;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = 8 * (T & 1) + 2 * t4 + e;
                    const int nb = colnew[c];
                    if (nb == -2) continue;
                    z[a][e] = y[a][e] = 0.0;
                    lo[a][e] = (nb >= 0 && r < m) ? p.l[(size_t)nb * m + r] : 0.0;
                    hi[a][e] = (nb >= 0 && r < m) ? p.u[(size_t)nb * m + r] : 0.0;
                    Wsm[r * kShLdb + c] = 0.0;
                }
            }
            __syncthreads();
            if (tid < kShNc) {
                const int c = tid, nb = colnew[c];
                if (nb != -2) {
                    colb[c] = nb;
                    colstart[c] = g;
                    colflag[c] = 0;
                    double nq = 0.0;
                    if (nb >= 0)
                        for (int j = 0; j < n; ++j) nq = nanmax(nq, fabs(p.q[(size_t)nb * n + j]));
                    normQ[c] = nq;
                }
            }
            __syncthreads();
            if (tid == 0) {
                int live = 0, cap = 0x7fffffff;
                for (int c = 0; c < kShNc; ++c)
                    if (colb[c] >= 0) {
                        ++live;
                        const long long cc = (long long)colstart[c] + p.s.max_iter;
                        if (cc < cap) cap = (int)cc;
                    }
                ctl[0] = live;
                ctl[1] = cap;
            }
            __syncthreads();
            refill = false;
            if (ctl[0] == 0) break;
            if (p.s.max_iter <= 0) {                            // nothing to iterate: every column is at its cap already
                if (tid < kShNc && colb[tid] >= 0) { colflag[tid] = 1; colres[3 * tid] = 0.0; }
                refill = true;
                continue;
            }
        }
        ++g;
        // ---- [1] RHS = sigma x - q + A' W
        {
            double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            for (int ks = 0; ks < ksA; ++ks) {
                const double a = Arm[(4 * ks + t4) * kShLda + row1];
                const double b0 = Wsm[(4 * ks + t4) * kShLdb + g8], b1 = Wsm[(4 * ks + t4) * kShLdb + 8 + g8];
                dmma8x8x4(acc[0], a, b0);
                dmma8x8x4(acc[1], a, b1);
            }
#pragma unroll
            for (int a = 0; a < 2; ++a)
                *reinterpret_cast<double2 *>(Rsm + row1 * kShLdb + 8 * a + 2 * t4) =
                    make_double2(sigma * x[a][0] - qv[a][0] + acc[a][0], sigma * x[a][1] - qv[a][1] + acc[a][1]);
        }
        __syncthreads();
        // ---- [2] X~ = K^-1 RHS ;  x = alpha x~ + (1 - alpha) x
        {
            double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll 4
            for (int ks = 0; ks < kDN / 4; ++ks) {
                const double a = Kin[row1 * kShLda + 4 * ks + t4];
                const double b0 = Rsm[(4 * ks + t4) * kShLdb + g8], b1 = Rsm[(4 * ks + t4) * kShLdb + 8 + g8];
                dmma8x8x4(acc[0], a, b0);
                dmma8x8x4(acc[1], a, b1);
            }
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                *reinterpret_cast<double2 *>(Xtm + row1 * kShLdb + 8 * a + 2 * t4) = make_double2(acc[a][0], acc[a][1]);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double x_old = x[a][e];
                    const double x_new = alpha * acc[a][e] + alpha1 * x_old;        // SolveQuadraticProgram.jl:57
                    x[a][e] = x_new;
                    dxv[a][e] = fabs(x_new - x_old);
                }
            }
        }
        __syncthreads();
        // ---- [3] Z~ = A X~ ;  z, y update ;  W = rho z - y
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int T = warp + kShWarps * a;
            if (T >= ntiles3) continue;
            const int r = 8 * (T >> 1) + g8, cb = 8 * (T & 1);
            double acc[2] = {0.0, 0.0};
#pragma unroll 4
            for (int ks = 0; ks < kDN / 4; ++ks)
                dmma8x8x4(acc, Arm[r * kShLda + 4 * ks + t4], Xtm[(4 * ks + t4) * kShLdb + cb + g8]);
            double w2[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double z_old = z[a][e], y_old = y[a][e];
                const double zr = alpha * acc[e] + alpha1 * z_old;
                const double z_new = clamp_julia(zr + rho1 * y_old, lo[a][e], hi[a][e]);   // :60
                const double y_new = y_old + rho * (zr - z_new);                           // :61
                z[a][e] = z_new;
                y[a][e] = y_new;
                dzv[a][e] = fabs(z_new - z_old);
                w2[e] = rho * z_new - y_new;
            }
            *reinterpret_cast<double2 *>(Wsm + r * kShLdb + cb + 2 * t4) = make_double2(w2[0], w2[1]);
        }
        const bool at_check = (g % (int)p.s.check_every) == 0;
        const bool at_cap = g >= ctl[1];
        if (!at_check && !at_cap) continue;

        // ---- CheckConvergence (:79-112) for all columns: A X, A' Y, P X as GEMMs, column-wise max norms
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 2; ++a)
            *reinterpret_cast<double2 *>(Rsm + row1 * kShLdb + 8 * a + 2 * t4) = make_double2(x[a][0], x[a][1]);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int T = warp + kShWarps * a;
            if (T >= ntiles3) continue;
            *reinterpret_cast<double2 *>(Ysm + (8 * (T >> 1) + g8) * kShLdb + 8 * (T & 1) + 2 * t4) = make_double2(y[a][0], y[a][1]);
        }
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int T = warp + kShWarps * a;
            if (T >= ntiles3) continue;
            const int r = 8 * (T >> 1) + g8, cb = 8 * (T & 1);
            double acc[2] = {0.0, 0.0};
#pragma unroll 4
            for (int ks = 0; ks < kDN / 4; ++ks)
                dmma8x8x4(acc, Arm[r * kShLda + 4 * ks + t4], Rsm[(4 * ks + t4) * kShLdb + cb + g8]);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double v0 = frag_colmax(dzv[a][e]);
                const double v1 = frag_colmax(fabs(acc[e] - z[a][e]));                        // |Ax - z|
                const double v2 = frag_colmax(nanmax(fabs(acc[e]), fabs(z[a][e])));           // max(|Ax|, |z|)
                if (g8 == 0) {
                    double *o = red3 + (T * 8 + 2 * t4 + e) * 3;
                    o[0] = v0; o[1] = v1; o[2] = v2;
                }
            }
        }
        {
            double aty[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, px[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            for (int ks = 0; ks < ksA; ++ks) {
                const double a = Arm[(4 * ks + t4) * kShLda + row1];
                dmma8x8x4(aty[0], a, Ysm[(4 * ks + t4) * kShLdb + g8]);
                dmma8x8x4(aty[1], a, Ysm[(4 * ks + t4) * kShLdb + 8 + g8]);
            }
#pragma unroll 4
            for (int ks = 0; ks < kDN / 4; ++ks) {
                const double a = Prm[row1 * kShLda + 4 * ks + t4];
                dmma8x8x4(px[0], a, Rsm[(4 * ks + t4) * kShLdb + g8]);
                dmma8x8x4(px[1], a, Rsm[(4 * ks + t4) * kShLdb + 8 + g8]);
            }
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double v0 = frag_colmax(dxv[a][e]);
                    const double v1 = frag_colmax(fabs(px[a][e] + qv[a][e] + aty[a][e]));      // |Px + q + A'y|
                    const double v2 = frag_colmax(nanmax(fabs(px[a][e]), fabs(aty[a][e])));
                    if (g8 == 0) {
                        double *o = red1 + ((2 * warp + a) * 8 + 2 * t4 + e) * 3;
                        o[0] = v0; o[1] = v1; o[2] = v2;
                    }
                }
        }
        __syncthreads();
        if (tid < kShNc && colb[tid] >= 0 && colflag[tid] == 0) {
            const int c = tid, nt = c >> 3, cc = c & 7;
            const int its = g - colstart[c];
            int flag = 0;
            if (at_check && (its % (int)p.s.check_every) == 0) {
                double dx = 0.0, rd = 0.0, nd = 0.0, dz = 0.0, rp = 0.0, np_ = 0.0;
                for (int w = 0; w < kShWarps; ++w) {
                    const double *o = red1 + ((2 * w + nt) * 8 + cc) * 3;
                    dx = nanmax(dx, o[0]); rd = nanmax(rd, o[1]); nd = nanmax(nd, o[2]);
                }
                for (int mt = 0; mt < mp / 8; ++mt) {
                    const double *o = red3 + ((2 * mt + nt) * 8 + cc) * 3;
                    dz = nanmax(dz, o[0]); rp = nanmax(rp, o[1]); np_ = nanmax(np_, o[2]);
                }
                const double max_dual = nanmax(nd, normQ[c]);
                flag = 1;
                if ((rp < p.s.eps_abs + p.s.eps_rel * np_) && (rd < p.s.eps_abs + p.s.eps_rel * max_dual)) flag = 3;   // :102
                if ((dx <= eps_admm) && (dz <= eps_admm)) flag = 2;                                                   // :105
                if (flag == 1) flag = 0;                        // keep iterating
            }
            if (flag == 0 && its >= p.s.max_iter) flag = 1;     // convNumItr
            colnew[c] = flag != 0 ? 1 : 0;                      // finished in this block: its x goes out below
            if (flag != 0) {
                colflag[c] = flag;
                colres[3 * c] = (double)its;
            }
        } else if (tid < kShNc) {
            colnew[tid] = 0;
        }
        __syncthreads();
        // the solution of a column is its x at the iteration it finished (a column that met its iteration cap between two
        // check points keeps iterating until its slot is refilled at the next one)
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = 8 * a + 2 * t4 + e;
                if (colnew[c] == 1 && row1 < n) p.X[(size_t)colb[c] * n + row1] = x[a][e];
            }
        // (Ysm / Rsm were scratch; W itself is untouched)  decide whether to refill
        {
            bool any = false;
            for (int c = 0; c < kShNc; ++c) any = any || (colb[c] >= 0 && colflag[c] != 0);
            // a finished column keeps iterating harmlessly until the next check point, where its slot is refilled
            refill = any && at_check;
            if (any && !at_check && tid == 0) {
                int cap = 0x7fffffff;
                for (int c = 0; c < kShNc; ++c)
                    if (colb[c] >= 0 && colflag[c] == 0) {
                        const long long c2 = (long long)colstart[c] + p.s.max_iter;
                        if (c2 < cap) cap = (int)c2;
                    }
                ctl[1] = cap;
            }
        }
        if (refill && tid < kShNc && colb[tid] >= 0 && colflag[tid] != 0) tot_iters += (unsigned long long)colres[3 * tid];
    }
    // (columns that finished at a cap between check points were counted when their slot was refilled)
    if (tid < kShNc && tot_iters) atomicAdd(p.totals, tot_iters);
}

}  // namespace qpb
