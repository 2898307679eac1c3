// dense_batch.cu -- batch of small dense QPs (MPC-style, configs[2]: 65 536 x (n = 64, m = 96)).
//
// Replaces SolveQuadraticProgram! (SolveQuadraticProgram.jl:14-112) driven by a *direct* plugin
// (LaLdl / QDLdl / FacLdl, LinearSystemSolvers.jl:16-107).  Eliminating nu from the KKT system those
// plugins factor gives the reduced system  (P + sigma I + rho A'A) x~ = sigma x - q + A'(rho z - y),
// z~ = A x~  (compare :37-40 with :134-139), which is what is factored here; a rho change triggers the
// same full refactorisation the reference does (:30-32, :61-63, :93-95).
//
// One CTA (128 threads) owns one QP at a time; A (m x 64, column-major), the packed factor and all
// vectors live in shared memory for the whole solve:
//   K = P + sigma I + rho A'A   SYRK on the FP64 tensor pipe (mma.sync m8n8k4 f64 -> SASS DMMA)
//   K = L L'                    blocked right-looking Cholesky, 8-wide panels, DMMA trailing update
//   Linv = L^-1                 in-place triangular inverse (packed storage)
//   x~ = Linv' (Linv rhs)       two triangular matrix-vector products per ADMM iteration
// The per-iteration work (2 GEMVs with A, 2 with Linv) runs on the FP64 FMA pipe out of shared memory.
#include <chrono>
#include <cmath>
#include <cstring>
#include <new>

#include "admm_kernels.cuh"
#include "host_common.h"

namespace qpb {

constexpr int kDN = 64;          // n padded to 64
constexpr int kDThreads = 128;
constexpr int kDPacked = kDN * (kDN + 1) / 2;   // 2080

struct DenseBatchParams {
    int batch, n, m, mp;         // mp = m rounded up to a multiple of 4
    const double *P, *A, *q, *l, *u;
    double *X;
    int *flags;
    long long *iters;
    int *factor_fail;            // set to 1 if any pivot was not positive
    unsigned long long *totals;  // [0] iterations, [1] rho updates
    unsigned int *queue;         // next problem index (zeroed before every launch)
    int blocked_chol;
    AdmmSettingsDev s;
};

__device__ __forceinline__ int pidx(int i, int j) { return i * (i + 1) / 2 + j; }   // j <= i

__device__ __forceinline__ void dmma8x8x4(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d[0]), "+d"(d[1])
                 : "d"(a), "d"(b));
}

struct DenseSmem {
    double *As;    // mp * 64   (i + mp * j)
    double *Lp;    // packed lower 64 x 64
    double *x, *xt, *rhs, *q, *tt, *colb;   // 64 each
    double *part;  // 128
    double *z, *y, *w, *l, *u;              // mp each
    double *red;   // 64
};

__device__ __forceinline__ DenseSmem carve(unsigned char *raw, int mp) {
    DenseSmem s;
    double *p = reinterpret_cast<double *>(raw);
    s.As = p; p += (size_t)mp * kDN;
    s.Lp = p; p += kDPacked + 16;
    s.x = p; p += kDN;
    s.xt = p; p += kDN;
    s.rhs = p; p += kDN;
    s.q = p; p += kDN;
    s.tt = p; p += kDN;
    s.colb = p; p += kDN;
    s.part = p; p += 2 * kDN;
    s.z = p; p += mp;
    s.y = p; p += mp;
    s.w = p; p += mp;
    s.l = p; p += mp;
    s.u = p; p += mp;
    s.red = p; p += 64;
    return s;
}

static size_t dense_smem_bytes(int mp) {
    return sizeof(double) * ((size_t)mp * kDN + kDPacked + 16 + 6 * kDN + 2 * kDN + 5 * (size_t)mp + 64);
}

// ---- K = P + sigma I + rho A'A (lower, packed) via DMMA ----------------------------------------
// Warp w owns the 8-row tiles w and 7-w of the lower triangle (9 tiles each: balanced).
__device__ __forceinline__ void build_K(const DenseSmem &sm, int mp, const double *Pg, int n, double rho, double sigma) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int rtA = warp, rtB = 7 - warp;
    double accA[4][2], accB[8][2];
#pragma unroll
    for (int c = 0; c < 4; ++c) accA[c][0] = accA[c][1] = 0.0;
#pragma unroll
    for (int c = 0; c < 8; ++c) accB[c][0] = accB[c][1] = 0.0;
    const double *As = sm.As;
    for (int kk = 0; kk < mp; kk += 4) {
        // fragment of column tile ct: element (k = kk + t, column 8 ct + g) -- serves as the A operand
        // (row-major A'[r][k]) of row tile ct and as the B operand (col-major A[k][c]) of column tile ct
        double f[8];
#pragma unroll
        for (int ct = 0; ct < 8; ++ct) f[ct] = As[(kk + t) + mp * (8 * ct + g)];
        const double fa = As[(kk + t) + mp * (8 * rtA + g)];
        const double fb = As[(kk + t) + mp * (8 * rtB + g)];
#pragma unroll
        for (int ct = 0; ct < 4; ++ct)
            if (ct <= rtA) dmma8x8x4(accA[ct], fa, f[ct]);
#pragma unroll
        for (int ct = 0; ct < 8; ++ct)
            if (ct <= rtB) dmma8x8x4(accB[ct], fb, f[ct]);
    }
    // C fragment: lane holds (row 8 rt + g, cols 8 ct + 2t, +1)
#pragma unroll
    for (int ct = 0; ct < 4; ++ct)
        if (ct <= rtA) {
            const int i = 8 * rtA + g;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = 8 * ct + 2 * t + e;
                if (j <= i) {
                    const double pij = (i < n && j < n) ? __ldg(Pg + i + (size_t)n * j) : 0.0;
                    sm.Lp[pidx(i, j)] = pij + rho * accA[ct][e] + (i == j ? (i < n ? sigma : 1.0) : 0.0);
                }
            }
        }
#pragma unroll
    for (int ct = 0; ct < 8; ++ct)
        if (ct <= rtB) {
            const int i = 8 * rtB + g;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = 8 * ct + 2 * t + e;
                if (j <= i) {
                    const double pij = (i < n && j < n) ? __ldg(Pg + i + (size_t)n * j) : 0.0;
                    sm.Lp[pidx(i, j)] = pij + rho * accB[ct][e] + (i == j ? (i < n ? sigma : 1.0) : 0.0);
                }
            }
        }
    __syncthreads();
}

// ---- unblocked right-looking Cholesky on the packed lower triangle (reference implementation) ----
__device__ __forceinline__ bool chol_unblocked(const DenseSmem &sm, int j0, int j1) {
    bool ok = true;
    double *Lp = sm.Lp, *colb = sm.colb;
    for (int j = j0; j < j1; ++j) {
        __syncthreads();
        double djj = Lp[pidx(j, j)];
        if (!(djj > 0.0)) { ok = false; djj = 1.0; }
        const double ljj = sqrt(djj), inv = 1.0 / ljj;
        const int i = j + 1 + threadIdx.x;
        if (i < kDN) {
            const double v = Lp[pidx(i, j)] * inv;
            colb[i] = v;
            Lp[pidx(i, j)] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) Lp[pidx(j, j)] = ljj;
        const int ti = threadIdx.x & 63, tk = threadIdx.x >> 6;
        if (ti > j) {
            const double ci = colb[ti];
            for (int k = j + 1 + tk; k <= ti; k += 2) Lp[pidx(ti, k)] -= ci * colb[k];
        }
    }
    __syncthreads();
    return ok;
}

// ---- blocked Cholesky: 8-wide panels, DMMA trailing update ---------------------------------------
// For panel p (columns 8p .. 8p+7): (1) factor the panel's columns with the unblocked kernel restricted
// to updates *inside* the panel, (2) trailing update of all tiles (rt, ct), p < ct <= rt, with
// C -= Lpanel(rt) Lpanel(ct)' on the tensor pipe (two k-steps of 4).
__device__ __forceinline__ bool chol_blocked(const DenseSmem &sm) {
    bool ok = true;
    double *Lp = sm.Lp, *colb = sm.colb;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int p = 0; p < 8; ++p) {
        const int c0 = 8 * p, c1 = c0 + 8;
        // (1) panel factorisation: columns c0..c1-1, all rows below; updates only columns < c1
        for (int j = c0; j < c1; ++j) {
            __syncthreads();
            double djj = Lp[pidx(j, j)];
            if (!(djj > 0.0)) { ok = false; djj = 1.0; }
            const double ljj = sqrt(djj), inv = 1.0 / ljj;
            const int i = j + 1 + threadIdx.x;
            if (i < kDN) {
                const double v = Lp[pidx(i, j)] * inv;
                colb[i] = v;
                Lp[pidx(i, j)] = v;
            }
            __syncthreads();
            if (threadIdx.x == 0) Lp[pidx(j, j)] = ljj;
            // rows i > j, columns k in (j, min(i, c1-1)]
            const int ti = threadIdx.x & 63, tk = threadIdx.x >> 6;
            if (ti > j) {
                const double ci = colb[ti];
                const int kend = ti < c1 - 1 ? ti : c1 - 1;
                for (int k = j + 1 + tk; k <= kend; k += 2) Lp[pidx(ti, k)] -= ci * colb[k];
            }
        }
        __syncthreads();
        // (2) trailing update on the tensor pipe: tiles (rt, ct) with p < ct <= rt < 8
        const int nt = 7 - p;                       // trailing tile rows/cols
        const int ntiles = nt * (nt + 1) / 2;
        for (int tile = warp; tile < ntiles; tile += 4) {
            // tile -> (a, b) with b <= a < nt  (row-major lower enumeration)
            int a = 0;
            while ((a + 1) * (a + 2) / 2 <= tile) ++a;
            const int b = tile - a * (a + 1) / 2;
            const int rt = p + 1 + a, ct = p + 1 + b;
            const int i = 8 * rt + g;
            const int j = 8 * ct + 2 * t;
            double acc[2] = {0.0, 0.0};
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const double fa = Lp[pidx(8 * rt + g, c0 + 4 * ks + t)];
                const double fb = Lp[pidx(8 * ct + g, c0 + 4 * ks + t)];
                dmma8x8x4(acc, fa, fb);
            }
            if (j <= i) Lp[pidx(i, j)] -= acc[0];
            if (j + 1 <= i) Lp[pidx(i, j + 1)] -= acc[1];
        }
    }
    __syncthreads();
    return ok;
}

// ---- in-place inverse of the packed lower-triangular factor (LAPACK dtrti2, lower, non-unit) ------
__device__ __forceinline__ void trtri_packed(const DenseSmem &sm) {
    double *Lp = sm.Lp, *colb = sm.colb, *part = sm.part;
    for (int j = kDN - 1; j >= 0; --j) {
        __syncthreads();
        const double ajj = 1.0 / Lp[pidx(j, j)];
        if ((int)threadIdx.x < kDN - 1 - j) colb[j + 1 + threadIdx.x] = Lp[pidx(j + 1 + threadIdx.x, j)];
        __syncthreads();
        if (threadIdx.x == 0) Lp[pidx(j, j)] = ajj;
        const int i = threadIdx.x & 63, th = threadIdx.x >> 6;
        double s = 0.0;
        if (i > j) {
            const int len = i - j, kmid = j + 1 + len / 2;
            const int ka = th == 0 ? j + 1 : kmid, kb = th == 0 ? kmid : i + 1;
            const double *row = Lp + pidx(i, 0);
            for (int k = ka; k < kb; ++k) s += row[k] * colb[k];
        }
        part[threadIdx.x] = s;
        __syncthreads();
        if ((int)threadIdx.x < kDN && (int)threadIdx.x > j)
            Lp[pidx(threadIdx.x, j)] = -ajj * (part[threadIdx.x] + part[threadIdx.x + 64]);
    }
    __syncthreads();
}

// Mapping of the per-iteration matrix-vector products: output o = tid >> 1 is computed by the two adjacent
// lanes h = tid & 1 (each one half of the sum) and combined with one shuffle -- no shared-memory round trip.
//
// s_o = sum_i A[i, o] v_i (o = 0..63): lane h sums rows [h*hlen, (h+1)*hlen), starting at a lane-dependent
// offset so that the 16 lanes of a shared-memory phase hit 16 different banks although the leading
// dimension mp is a multiple of 16.
__device__ __forceinline__ double at_times_v(const double *As, int mp, const double *v) {
    const int o = threadIdx.x >> 1, h = threadIdx.x & 1;
    const int hlen = mp >> 1;                 // mp is a multiple of 4
    const double *col = As + (size_t)mp * o + h * hlen;
    const double *vv = v + h * hlen;
    const int i0 = (threadIdx.x & 15) < hlen ? (threadIdx.x & 15) : 0;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int i = i0;
    for (; i + 4 <= hlen; i += 4) {
        s0 += col[i] * vv[i];
        s1 += col[i + 1] * vv[i + 1];
        s2 += col[i + 2] * vv[i + 2];
        s3 += col[i + 3] * vv[i + 3];
    }
    for (; i < hlen; ++i) s0 += col[i] * vv[i];
    for (i = 0; i < i0; ++i) s1 += col[i] * vv[i];
    double s = (s0 + s1) + (s2 + s3);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    return s;
}

// block-wide max of `nv` values held per thread (NaN-propagating), broadcast to all threads
template <int NV>
__device__ __forceinline__ void block_max(double (&v)[NV], double *red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x = nanmax(x, __shfl_xor_sync(0xffffffffu, x, o));
        if (lane == 0) red[warp * NV + i] = x;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double x = red[i];
#pragma unroll
        for (int w = 1; w < kDThreads / 32; ++w) x = nanmax(x, red[w * NV + i]);
        v[i] = x;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kDThreads) dense_batch_kernel(DenseBatchParams p) {
    extern __shared__ __align__(16) unsigned char raw[];
    const DenseSmem sm = carve(raw, p.mp);
    const int n = p.n, m = p.m, mp = p.mp;
    const int tid = threadIdx.x;
    const double alpha = p.s.alpha, alpha1 = 1.0 - alpha, sigma = p.s.sigma;
    const double eps_admm = fmin(p.s.eps_abs, p.s.eps_rel) * 1e-2;
    unsigned long long tot_iters = 0, tot_rho = 0;

    __shared__ int next_b;
    for (;;) {
        // dynamic work queue: iteration counts vary by 100x between problems, static striding leaves a long tail
        __syncthreads();
        if (tid == 0) next_b = (int)atomicAdd(p.queue, 1u);
        __syncthreads();
        const int b = next_b;
        if (b >= p.batch) break;
        const double *Pg = p.P + (size_t)b * n * n;
        const double *Ag = p.A + (size_t)b * m * n;
        __syncthreads();
        // ---- load the problem into shared memory (zero padded to mp x 64)
        for (int idx = tid; idx < mp * kDN; idx += kDThreads) {
            const int i = idx % mp, j = idx / mp;
            sm.As[idx] = (i < m && j < n) ? __ldg(Ag + i + (size_t)m * j) : 0.0;
        }
        if (tid < kDN) {
            sm.q[tid] = tid < n ? p.q[(size_t)b * n + tid] : 0.0;
            sm.x[tid] = tid < n ? p.X[(size_t)b * n + tid] : 0.0;
            sm.xt[tid] = 0.0;
        }
        for (int i = tid; i < mp; i += kDThreads) {
            sm.l[i] = i < m ? p.l[(size_t)b * m + i] : 0.0;
            sm.u[i] = i < m ? p.u[(size_t)b * m + i] : 0.0;
            sm.z[i] = 0.0;
            sm.y[i] = 0.0;
            sm.w[i] = 0.0;
        }
        double normQ = 0.0;
        {
            double v[1] = {tid < n ? fabs(p.q[(size_t)b * n + tid]) : 0.0};
            __syncthreads();
            block_max<1>(v, sm.red);
            normQ = v[0];
        }

        double rho = p.s.rho, rho1 = 1.0 / rho, rhorho = rho;
        int conv_flag = 1;
        bool need_factor = true, fact_ok = true;
        long long ii = 0;
        for (ii = 1; ii <= p.s.max_iter; ++ii) {
            // ---- rho trigger (SolveQuadraticProgram.jl:46-52) -> full refactorisation
            if (p.s.adaptive_rho && ((rhorho * p.s.rho_factor < rho) || (rhorho > p.s.rho_factor * rho))) {
                rho = rhorho;
                rho1 = 1.0 / rho;
                need_factor = true;
                ++tot_rho;
                for (int i = tid; i < mp; i += kDThreads) sm.w[i] = rho * sm.z[i] - sm.y[i];
            }
            if (need_factor) {
                __syncthreads();
                build_K(sm, mp, Pg, n, rho, sigma);
                const bool ok = p.blocked_chol ? chol_blocked(sm) : chol_unblocked(sm, 0, kDN);
                fact_ok = fact_ok && ok;
                trtri_packed(sm);
                need_factor = false;
            }
            __syncthreads();
            // ---- rhs = sigma x - q + A' w,  w = rho z - y      (LinearSystemSolvers.jl:37-38 reduced)
            const int o = tid >> 1, h = tid & 1;
            {
                const double s = at_times_v(sm.As, mp, sm.w);
                if (h == 0) sm.rhs[o] = sigma * sm.x[o] - sm.q[o] + s;
            }
            __syncthreads();
            // ---- t = Linv rhs   (row o: j = 0..o, split between the two lanes)
            {
                const int mid = (o + 1) >> 1;
                const int ja = h == 0 ? 0 : mid, jb = h == 0 ? mid : o + 1;
                const double *row = sm.Lp + pidx(o, 0);
                double s0 = 0.0, s1 = 0.0;
                int j = ja;
                for (; j + 2 <= jb; j += 2) {
                    s0 += row[j] * sm.rhs[j];
                    s1 += row[j + 1] * sm.rhs[j + 1];
                }
                if (j < jb) s0 += row[j] * sm.rhs[j];
                double s = s0 + s1;
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                if (h == 0) sm.tt[o] = s;
            }
            __syncthreads();
            // ---- x~ = Linv' t   (column o: k = o..63), then the x relaxation (:57)
            double dx = 0.0, dz = 0.0;
            {
                const int mid = o + ((kDN - o + 1) >> 1);
                const int ka = h == 0 ? o : mid, kb = h == 0 ? mid : kDN;
                double s0 = 0.0, s1 = 0.0;
                int k = ka;
                int idx = pidx(ka, o);
                for (; k + 2 <= kb; k += 2) {
                    s0 += sm.Lp[idx] * sm.tt[k];
                    s1 += sm.Lp[idx + k + 1] * sm.tt[k + 1];
                    idx += 2 * k + 3;
                }
                if (k < kb) s0 += sm.Lp[idx] * sm.tt[k];
                double s = s0 + s1;
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                if (h == 0) {
                    sm.xt[o] = s;
                    const double x_old = sm.x[o];
                    const double x_new = alpha * s + alpha1 * x_old;         // :57
                    sm.x[o] = x_new;
                    dx = fabs(x_new - x_old);
                }
            }
            __syncthreads();
            // ---- z~ = A x~, then the z / y update (:59-61), row-local
            if (tid < m) {
                const double *row = sm.As + tid;
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 4
                for (int j = 0; j < kDN; j += 4) {
                    s0 += row[(size_t)mp * j] * sm.xt[j];
                    s1 += row[(size_t)mp * (j + 1)] * sm.xt[j + 1];
                    s2 += row[(size_t)mp * (j + 2)] * sm.xt[j + 2];
                    s3 += row[(size_t)mp * (j + 3)] * sm.xt[j + 3];
                }
                const double zt = (s0 + s1) + (s2 + s3);
                const double z_old = sm.z[tid], y_old = sm.y[tid];
                const double zr = alpha * zt + alpha1 * z_old;
                const double z_new = clamp_julia(zr + rho1 * y_old, sm.l[tid], sm.u[tid]);
                const double y_new = y_old + rho * (zr - z_new);
                sm.z[tid] = z_new;
                sm.y[tid] = y_new;
                sm.w[tid] = rho * z_new - y_new;
                dz = fabs(z_new - z_old);
            }
            if (ii % p.s.check_every == 0) {
                // ---- CheckConvergence (:79-112)
                __syncthreads();
                double nr[6] = {dx, dz, 0.0, 0.0, 0.0, 0.0};   // dx dz rp max(|Ax|,|z|) rd max(|Px|,|A'y|)
                if (tid < m) {
                    const double *row = sm.As + tid;
                    double s0 = 0.0;
                    for (int j = 0; j < kDN; ++j) s0 += row[(size_t)mp * j] * sm.x[j];
                    const double zi = sm.z[tid];
                    nr[2] = fabs(s0 - zi);
                    nr[3] = nanmax(fabs(s0), fabs(zi));
                }
                const double aty = at_times_v(sm.As, mp, sm.y);
                double px = 0.0;
                if (o < n) {
                    const int ja = h == 0 ? 0 : (n >> 1), jb = h == 0 ? (n >> 1) : n;
                    for (int j = ja; j < jb; ++j) px += __ldg(Pg + o + (size_t)n * j) * sm.x[j];
                }
                px += __shfl_xor_sync(0xffffffffu, px, 1);
                if (h == 0) {
                    nr[4] = fabs(px + sm.q[o] + aty);
                    nr[5] = nanmax(fabs(px), fabs(aty));
                }
                block_max<6>(nr, sm.red);
                const double res_prim = nr[2], res_dual = nr[4];
                const double max_prim = nr[3], max_dual = nanmax(nr[5], normQ);
                if (p.s.adaptive_rho) {
                    const double num = res_prim * max_dual, den = res_dual * max_prim;
                    rhorho = clamp_julia(rho * sqrt(num / den), 1e-3, 1e6);
                }
                if ((res_prim < p.s.eps_abs + p.s.eps_rel * max_prim) && (res_dual < p.s.eps_abs + p.s.eps_rel * max_dual))
                    conv_flag = 3;
                if ((nr[0] <= eps_admm) && (nr[1] <= eps_admm)) conv_flag = 2;
                if (conv_flag != 1) break;
            }
        }
        if (ii > p.s.max_iter) ii = p.s.max_iter;
        __syncthreads();
        if (tid < n) p.X[(size_t)b * n + tid] = sm.x[tid];
        if (tid == 0) {
            if (p.flags) p.flags[b] = conv_flag;
            if (p.iters) p.iters[b] = ii;
            if (!fact_ok) *p.factor_fail = 1;
        }
        tot_iters += (unsigned long long)ii;
    }
    if (tid == 0) {
        atomicAdd(p.totals + 0, tot_iters);
        atomicAdd(p.totals + 1, tot_rho);
    }
}

struct DenseBatch {
    int64_t batch = 0;
    int n = 0, m = 0, mp = 0, device = -1, grid = 0;
    qpb200_settings settings{};
    DenseBatchParams prm{};
    DeviceArena arena;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double setup_ms = 0.0;
    size_t smem = 0;
    ~DenseBatch() {
        if (device >= 0) cudaSetDevice(device);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (stream) cudaStreamDestroy(stream);
        arena.release();
    }
};

}  // namespace qpb

struct qpb200_batch {
    qpb::DenseBatch b;
};

using namespace qpb;

extern "C" {

int qpb200_batch_create(qpb200_batch **out, int64_t batch, int64_t n, int64_t m, const double *P, const double *A,
                        const double *q, const double *l, const double *u, const qpb200_settings *settings) {
    if (!out) return fail(QPB200_ERR_ARG, "qpb200_batch_create: out is NULL");
    *out = nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    if (batch <= 0 || n <= 0 || n > kDN || m <= 0 || m > 128)
        return fail(QPB200_ERR_ARG, "qpb200_batch_create: need batch > 0, 0 < n <= 64, 0 < m <= 128 (got %lld, %lld, %lld)",
                    (long long)batch, (long long)n, (long long)m);
    if (batch >= (int64_t(1) << 31)) return fail(QPB200_ERR_ARG, "qpb200_batch_create: batch too large");
    if (!P || !A || !q || !l || !u) return fail(QPB200_ERR_ARG, "qpb200_batch_create: NULL array");
    qpb200_settings s;
    if (settings) s = *settings;
    else { qpb200_default_settings(&s); s.lin_solver = QPB200_LINSOLVE_CHOLESKY; }
    if (!(s.rho > 0.0) || !(s.sigma >= 0.0) || s.max_iter < 0 || s.check_every <= 0)
        return fail(QPB200_ERR_ARG, "settings: need rho > 0, sigma >= 0, max_iter >= 0, check_every > 0");
    if (s.lin_solver != QPB200_LINSOLVE_CHOLESKY)
        return fail(QPB200_ERR_ARG, "qpb200_batch_create: the dense batch path implements lin_solver = QPB200_LINSOLVE_CHOLESKY only");
    // value checks on a strided sample would miss entries: scan everything (memory-bound, ~GB/s)
    const size_t nP = (size_t)batch * n * n, nA = (size_t)batch * m * n;
    if (!all_finite(P, nP)) return fail(QPB200_ERR_NONFINITE, "P has a non-finite entry");
    if (!all_finite(A, nA)) return fail(QPB200_ERR_NONFINITE, "A has a non-finite entry");
    if (!all_finite(q, (size_t)batch * n)) return fail(QPB200_ERR_NONFINITE, "q has a non-finite entry");
    for (size_t i = 0; i < (size_t)batch * m; ++i)
        if (std::isnan(l[i]) || std::isnan(u[i]) || l[i] > u[i]) return fail(QPB200_ERR_NONFINITE, "bounds: need l <= u, not NaN (entry %zu)", i);
    int rc = check_device(s.device);
    if (rc) return rc;
    qpb200_batch *h = new (std::nothrow) qpb200_batch();
    if (!h) return fail(QPB200_ERR_ARG, "out of host memory");
    DenseBatch &B = h->b;
#define QPB_CUDA_H(call)                                                                                     \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) {                                                                             \
            delete h;                                                                                        \
            return fail(QPB200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                                    \
    } while (0)
    QPB_CUDA_H(cudaGetDevice(&B.device));
    B.batch = batch; B.n = (int)n; B.m = (int)m; B.mp = ((int)m + 3) & ~3;
    B.settings = s;
    double *dP, *dA, *dq, *dl, *du;
    QPB_CUDA_H(B.arena.alloc(&dP, nP));
    QPB_CUDA_H(B.arena.alloc(&dA, nA));
    QPB_CUDA_H(B.arena.alloc(&dq, (size_t)batch * n));
    QPB_CUDA_H(B.arena.alloc(&dl, (size_t)batch * m));
    QPB_CUDA_H(B.arena.alloc(&du, (size_t)batch * m));
    QPB_CUDA_H(B.arena.alloc(&B.prm.X, (size_t)batch * n));
    QPB_CUDA_H(B.arena.alloc(&B.prm.flags, (size_t)batch));
    QPB_CUDA_H(B.arena.alloc(&B.prm.iters, (size_t)batch));
    QPB_CUDA_H(B.arena.alloc(&B.prm.factor_fail, 1, true));
    QPB_CUDA_H(B.arena.alloc(&B.prm.totals, 4, true));
    QPB_CUDA_H(B.arena.alloc(&B.prm.queue, 4, true));
    QPB_CUDA_H(cudaStreamCreateWithFlags(&B.stream, cudaStreamNonBlocking));
    QPB_CUDA_H(cudaEventCreate(&B.ev0));
    QPB_CUDA_H(cudaEventCreate(&B.ev1));
    QPB_CUDA_H(cudaMemcpyAsync(dP, P, nP * sizeof(double), cudaMemcpyHostToDevice, B.stream));
    QPB_CUDA_H(cudaMemcpyAsync(dA, A, nA * sizeof(double), cudaMemcpyHostToDevice, B.stream));
    QPB_CUDA_H(cudaMemcpyAsync(dq, q, (size_t)batch * n * sizeof(double), cudaMemcpyHostToDevice, B.stream));
    QPB_CUDA_H(cudaMemcpyAsync(dl, l, (size_t)batch * m * sizeof(double), cudaMemcpyHostToDevice, B.stream));
    QPB_CUDA_H(cudaMemcpyAsync(du, u, (size_t)batch * m * sizeof(double), cudaMemcpyHostToDevice, B.stream));
    B.prm.batch = (int)batch; B.prm.n = B.n; B.prm.m = B.m; B.prm.mp = B.mp;
    B.prm.P = dP; B.prm.A = dA; B.prm.q = dq; B.prm.l = dl; B.prm.u = du;
    B.prm.blocked_chol = s.reserved_i[0] == 1 ? 0 : 1;     // reserved_i[0] = 1 selects the unblocked factor (A/B testing)
    AdmmSettingsDev &d = B.prm.s;
    d.max_iter = s.max_iter; d.check_every = s.check_every; d.pcg_max_iter = 0;
    d.eps_abs = s.eps_abs; d.eps_rel = s.eps_rel; d.rho = s.rho; d.sigma = s.sigma; d.alpha = s.alpha;
    d.rho_factor = s.rho_factor; d.pcg_eps = 0; d.pcg_rel_eps = 0; d.adaptive_rho = s.adaptive_rho;
    B.smem = dense_smem_bytes(B.mp);
    QPB_CUDA_H(cudaFuncSetAttribute(dense_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B.smem));
    int per_sm = 0;
    QPB_CUDA_H(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dense_batch_kernel, kDThreads, B.smem));
    if (per_sm < 1) { delete h; return fail(QPB200_ERR_CUDA, "dense batch kernel does not fit on an SM (smem %zu)", B.smem); }
    cudaDeviceProp prop;
    QPB_CUDA_H(cudaGetDeviceProperties(&prop, B.device));
    B.grid = (int)std::min<int64_t>(batch, (int64_t)prop.multiProcessorCount * per_sm);
    QPB_CUDA_H(cudaStreamSynchronize(B.stream));
    B.setup_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    *out = h;
    return QPB200_OK;
#undef QPB_CUDA_H
}

int qpb200_batch_solve(qpb200_batch *h, double *X_inout, int32_t *flags, int64_t *iters, qpb200_info *info) {
    if (!h || !X_inout) return fail(QPB200_ERR_ARG, "qpb200_batch_solve: NULL argument");
    DenseBatch &B = h->b;
    QPB_CUDA(cudaSetDevice(B.device));
    const size_t nx = (size_t)B.batch * B.n;
    QPB_CUDA(cudaMemcpyAsync(B.prm.X, X_inout, nx * sizeof(double), cudaMemcpyHostToDevice, B.stream));
    QPB_CUDA(cudaMemsetAsync(B.prm.totals, 0, 4 * sizeof(unsigned long long), B.stream));
    QPB_CUDA(cudaMemsetAsync(B.prm.factor_fail, 0, sizeof(int), B.stream));
    QPB_CUDA(cudaMemsetAsync(B.prm.queue, 0, 4 * sizeof(unsigned int), B.stream));
    QPB_CUDA(cudaEventRecord(B.ev0, B.stream));
    dense_batch_kernel<<<B.grid, kDThreads, B.smem, B.stream>>>(B.prm);
    QPB_CUDA(cudaGetLastError());
    QPB_CUDA(cudaEventRecord(B.ev1, B.stream));
    QPB_CUDA(cudaMemcpyAsync(X_inout, B.prm.X, nx * sizeof(double), cudaMemcpyDeviceToHost, B.stream));
    if (flags) QPB_CUDA(cudaMemcpyAsync(flags, B.prm.flags, (size_t)B.batch * sizeof(int), cudaMemcpyDeviceToHost, B.stream));
    if (iters) QPB_CUDA(cudaMemcpyAsync(iters, B.prm.iters, (size_t)B.batch * sizeof(long long), cudaMemcpyDeviceToHost, B.stream));
    unsigned long long tot[4];
    int ffail = 0;
    QPB_CUDA(cudaMemcpyAsync(tot, B.prm.totals, sizeof(tot), cudaMemcpyDeviceToHost, B.stream));
    QPB_CUDA(cudaMemcpyAsync(&ffail, B.prm.factor_fail, sizeof(int), cudaMemcpyDeviceToHost, B.stream));
    QPB_CUDA(cudaStreamSynchronize(B.stream));
    float ms = 0.f;
    QPB_CUDA(cudaEventElapsedTime(&ms, B.ev0, B.ev1));
    if (info) {
        std::memset(info, 0, sizeof(*info));
        info->conv_flag = 0;
        info->iterations = (int64_t)tot[0];
        info->rho_updates = (int64_t)tot[1];
        info->rho_final = B.settings.rho;
        info->res_prim = NAN;
        info->res_dual = NAN;
        info->solve_ms = ms;
        info->setup_ms = B.setup_ms;
        info->kernel_launches = 1;
    }
    if (ffail) return fail(QPB200_ERR_FACTOR, "Cholesky breakdown: a pivot of P + sigma I + rho A'A was not positive");
    return QPB200_OK;
}

void qpb200_batch_destroy(qpb200_batch *h) { delete h; }

}  // extern "C"
