// dense_kernel.cuh -- batch of small dense QPs (MPC-style, configs[2]: 65 536 x (n = 64, m = 96)).
//
// Replaces SolveQuadraticProgram! (SolveQuadraticProgram.jl:14-112) driven by a *direct* plugin
// (LaLdl / QDLdl / FacLdl, LinearSystemSolvers.jl:16-107).  Eliminating nu from the KKT system those
// plugins factor gives the reduced system  (P + sigma I + rho A'A) x~ = sigma x - q + A'(rho z - y),
// z~ = A x~  (compare :37-40 with :134-139), which is what is factored here; a rho change triggers the
// same full refactorisation the reference does (:30-32, :61-63, :93-95).
//
// One CTA (128 threads) owns one QP at a time (dynamic work queue); A, the factor and all vectors live in
// shared memory for the whole solve:
//   K = P + sigma I + rho A'A     SYRK on the FP64 tensor pipe (mma.sync m8n8k4 f64 -> SASS DMMA)
//   K = L L'                      blocked right-looking Cholesky, 8-wide panels, DMMA trailing update
//   K^-1 = L^-T L^-1              in-place triangular inverse (dtrti2) + in-place L'L product (dlauu2)
//   per ADMM iteration            rhs = sigma x - q + A'w ;  x~ = K^-1 rhs ;  z~ = A x~   -- three dense,
//                                 branch-free matrix-vector products with 128-bit shared-memory loads
// The explicit inverse has the same worst-case error order, cond(K) eps, as two products with L^-1, and ADMM
// re-corrects the x~ error every iteration; measured against exact-solve mode D: identical flags and
// iteration counts, |x - x_ref| <= 1e-9 (tests/test_gpu_dense_batch.py).
//
// Shared-memory layout: A is m x 64 column-major with leading dimension mp + 2, K / K^-1 is 64 x 64 row-major
// with leading dimension 66.  Both paddings are even (rows/columns stay 16-byte aligned for LDS.128) and shift
// consecutive columns/rows by one 16-byte bank group, so the 8 lanes of a 128-bit shared-memory phase hit 8
// different bank groups without any index skewing.
#pragma once
#include "admm_kernels.cuh"

namespace qpb {

constexpr int kDN = 64;          // n padded to 64
constexpr int kDThreads = 128;
constexpr int kLd = kDN + 2;     // leading dimension of the 64 x 64 factor / inverse

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

__device__ __forceinline__ int pidx(int i, int j) { return i * kLd + j; }

__device__ __forceinline__ void dmma8x8x4(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d[0]), "+d"(d[1])
                 : "d"(a), "d"(b));
}

struct DenseSmem {
    double *As;    // lda * 64   (i + lda * j), lda = mp + 2
    double *Lp;    // 64 x kLd row-major: K, then L, then L^-1, then K^-1 (full, symmetric)
    double *x, *xt, *rhs, *q, *colb;   // 64 each
    double *part;  // 128
    double *z, *y, *w, *l, *u;         // mp each
    double *red;   // 64
    double *part4; // 4 x 64: per-warp partial sums of A'w (register-resident A variant)
};

__device__ __forceinline__ DenseSmem carve(unsigned char *raw, int mp) {
    DenseSmem s;
    double *p = reinterpret_cast<double *>(raw);
    s.As = p; p += (size_t)(mp + 2) * kDN;
    s.Lp = p; p += kDN * kLd;
    s.x = p; p += kDN;
    s.xt = p; p += kDN;
    s.rhs = p; p += kDN;
    s.q = p; p += kDN;
    s.colb = p; p += kDN;
    s.part = p; p += 2 * kDN;
    s.z = p; p += mp;
    s.y = p; p += mp;
    s.w = p; p += mp;
    s.l = p; p += mp;
    s.u = p; p += mp;
    s.red = p; p += 64;
    s.part4 = p; p += 4 * kDN;
    return s;
}

static size_t dense_smem_bytes(int mp) {
    return sizeof(double) * ((size_t)(mp + 2) * kDN + kDN * kLd + 5 * kDN + 2 * kDN + 5 * (size_t)mp + 64 + 4 * kDN);
}

// ---- K = P + sigma I + rho A'A (lower triangle) via DMMA ----------------------------------------
// Warp w owns the 8-row tiles w and 7-w of the lower triangle (9 tiles each: balanced).
__device__ __forceinline__ void build_K(const DenseSmem &sm, int mp, const double *Pg, int n, double rho, double sigma) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int rtA = warp, rtB = 7 - warp;
    const int lda = mp + 2;
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
        for (int ct = 0; ct < 8; ++ct) f[ct] = As[(kk + t) + lda * (8 * ct + g)];
        const double fa = As[(kk + t) + lda * (8 * rtA + g)];
        const double fb = As[(kk + t) + lda * (8 * rtB + g)];
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

// ---- unblocked right-looking Cholesky on the lower triangle (A/B reference for the blocked one) ----
__device__ __forceinline__ bool chol_unblocked(const DenseSmem &sm) {
    bool ok = true;
    double *Lp = sm.Lp, *colb = sm.colb;
    for (int j = 0; j < kDN; ++j) {
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
// For panel p (columns 8p .. 8p+7): (1) factor the panel's columns, updates restricted to the panel,
// (2) trailing update of all tiles (rt, ct), p < ct <= rt, with C -= Lpanel(rt) Lpanel(ct)' on the tensor
// pipe (two k-steps of 4).
__device__ __forceinline__ bool chol_blocked(const DenseSmem &sm) {
    bool ok = true;
    double *Lp = sm.Lp, *colb = sm.colb;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int p = 0; p < 8; ++p) {
        const int c0 = 8 * p, c1 = c0 + 8;
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
        const int nt = 7 - p;                       // trailing tile rows/cols
        const int ntiles = nt * (nt + 1) / 2;
        for (int tile = warp; tile < ntiles; tile += 4) {
            int a = 0;                              // tile -> (a, b), b <= a < nt (row-major lower enumeration)
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

// ---- in-place inverse of the lower-triangular factor (LAPACK dtrti2, lower, non-unit) -------------
__device__ __forceinline__ void trtri_lower(const DenseSmem &sm) {
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

// ---- K^-1 = Linv' Linv in place (LAPACK dlauu2, lower), then mirrored to the full square ------------
// Row i of the result needs rows k >= i of Linv only, so rows are finished in ascending order:
//   Kinv[i][j] = sum_{k >= i} Linv[k][i] Linv[k][j],  j <= i.
__device__ __forceinline__ void lauum_lower_and_mirror(const DenseSmem &sm) {
    double *Lp = sm.Lp, *part = sm.part;
    for (int i = 0; i < kDN; ++i) {
        __syncthreads();
        const int j = threadIdx.x & 63, th = threadIdx.x >> 6;
        double s = 0.0;
        if (j <= i) {
            const int len = kDN - i, kmid = i + (len + 1) / 2;
            const int ka = th == 0 ? i : kmid, kb = th == 0 ? kmid : kDN;
            for (int k = ka; k < kb; ++k) s += Lp[pidx(k, i)] * Lp[pidx(k, j)];
        }
        part[threadIdx.x] = s;
        __syncthreads();                      // every read of row i (k = i terms) is done
        if ((int)threadIdx.x <= i) Lp[pidx(i, threadIdx.x)] = part[threadIdx.x] + part[threadIdx.x + 64];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < kDN * kDN; e += kDThreads) {
        const int i = e >> 6, j = e & 63;
        if (j > i) Lp[pidx(i, j)] = Lp[pidx(j, i)];
    }
    __syncthreads();
}

// ---- per-iteration matrix-vector products ---------------------------------------------------------
// Thread -> (output o, half h): o = (tid & 7) + 8 (tid >> 4), h = (tid >> 3) & 1.  The 8 lanes of a 128-bit
// shared-memory phase therefore work on 8 consecutive outputs with the same h (8 different bank groups,
// thanks to the padded leading dimensions), and the two halves of one output sit 8 lanes apart
// (combined with one __shfl_xor(.., 8)).
__device__ __forceinline__ int out_index() { return (threadIdx.x & 7) + 8 * (threadIdx.x >> 4); }
__device__ __forceinline__ int out_half() { return (threadIdx.x >> 3) & 1; }

// s_o = sum_i A[i, o] v_i,  o = 0..63   (both lanes of a pair return the full sum)
// MPC > 0: the padded row count is a compile-time constant (configs[2]: 96) -> fully unrolled, the 128-bit loads
// of a batch are all issued before the first FMA that needs them (the kernel is shared-memory-latency bound).
template <int MPC>
__device__ __forceinline__ double at_times_v(const double *As, int mp_rt, const double *v) {
    const int mp = MPC ? MPC : mp_rt;
    const int o = out_index(), h = out_half();
    const int hlen = mp >> 1;                                    // mp % 4 == 0 -> hlen even
    const double2 *col = reinterpret_cast<const double2 *>(As + (size_t)(mp + 2) * o + h * hlen);
    const double2 *vv = reinterpret_cast<const double2 *>(v + h * hlen);
    const int nd2 = hlen >> 1;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int c = 0;
    if (MPC) {
#pragma unroll
        for (int cb = 0; cb + 8 <= nd2; cb += 8) {
            double2 a[8], b[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { a[e] = col[cb + e]; b[e] = vv[cb + e]; }
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
                s0 += a[e].x * b[e].x;
                s1 += a[e].y * b[e].y;
                s2 += a[e + 1].x * b[e + 1].x;
                s3 += a[e + 1].y * b[e + 1].y;
            }
        }
        c = nd2 & ~7;
    }
    for (; c + 2 <= nd2; c += 2) {
        const double2 a0 = col[c], b0 = vv[c], a1 = col[c + 1], b1 = vv[c + 1];
        s0 += a0.x * b0.x;
        s1 += a0.y * b0.y;
        s2 += a1.x * b1.x;
        s3 += a1.y * b1.y;
    }
    if (c < nd2) {
        const double2 a0 = col[c], b0 = vv[c];
        s0 += a0.x * b0.x;
        s1 += a0.y * b0.y;
    }
    double s = (s0 + s1) + (s2 + s3);
    s += __shfl_xor_sync(0xffffffffu, s, 8);
    return s;
}

// s_o = sum_j Kinv[o][j] v_j,  o = 0..63
__device__ __forceinline__ double kinv_times_v(const double *Kf, const double *v) {
    const int o = out_index(), h = out_half();
    const double2 *row = reinterpret_cast<const double2 *>(Kf + pidx(o, 32 * h));
    const double2 *vv = reinterpret_cast<const double2 *>(v + 32 * h);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int cb = 0; cb < 16; cb += 8) {
        double2 a[8], b[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { a[e] = row[cb + e]; b[e] = vv[cb + e]; }
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
            s0 += a[e].x * b[e].x;
            s1 += a[e].y * b[e].y;
            s2 += a[e + 1].x * b[e + 1].x;
            s3 += a[e + 1].y * b[e + 1].y;
        }
    }
    double s = (s0 + s1) + (s2 + s3);
    s += __shfl_xor_sync(0xffffffffu, s, 8);
    return s;
}

// (A v)_i for the row pair (2p, 2p+1), p = out_index() < mp / 2: lane h sums columns [32 h, 32 h + 32); after
// the shuffle both lanes hold both sums; returns the sum of row 2p + h (so every thread owns ONE row).
template <int MPC>
__device__ __forceinline__ double a_times_v_row(const double *As, int mp_rt, const double *v, int &row) {
    const int mp = MPC ? MPC : mp_rt;
    const int p = out_index(), h = out_half();
    const int lda = mp + 2;
    row = 2 * p + h;
    double r0 = 0.0, r1 = 0.0;
    if (2 * p < mp) {
        const double *base = As + 2 * p + (size_t)lda * (32 * h);
        const double2 *vv = reinterpret_cast<const double2 *>(v + 32 * h);
        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll
        for (int cb = 0; cb < 16; cb += 4) {
            double2 xv[4], m0[4], m1[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                xv[e] = vv[cb + e];
                m0[e] = *reinterpret_cast<const double2 *>(base + (size_t)lda * (2 * (cb + e)));
                m1[e] = *reinterpret_cast<const double2 *>(base + (size_t)lda * (2 * (cb + e) + 1));
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                a0 += m0[e].x * xv[e].x;
                a1 += m0[e].y * xv[e].x;
                b0 += m1[e].x * xv[e].y;
                b1 += m1[e].y * xv[e].y;
            }
        }
        r0 = a0 + b0;
        r1 = a1 + b1;
    }
    r0 += __shfl_xor_sync(0xffffffffu, r0, 8);
    r1 += __shfl_xor_sync(0xffffffffu, r1, 8);
    return h == 0 ? r0 : r1;
}

// ---- register-resident A (compile-time shape mp = 96, the configs[2] shape) ---------------------------------
// The two products with A dominate the shared-memory traffic of an iteration (2 x 48 KB of the 131 KB); here every
// thread keeps a 6 x 8 block of A in registers for the whole solve: warp w owns rows [24 w, 24 w + 24), lane
// 8 g + c owns rows 24 w + 6 g + (0..5) and columns 8 c + (0..7).  A x~ reduces across the 8 lanes of a row group
// and A'w across the 4 row groups of a warp by butterfly reduce-scatters (13 FP64 shuffles per thread and
// iteration), the 4 warps' partial sums of A'w meet in shared memory.  Fixed summation order => reproducible.
struct RegA {
    double a[6][8];
};

__device__ __forceinline__ void rega_load(RegA &R, const double *As) {
    constexpr int lda = 96 + 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 3, c = lane & 7;
    const double *base = As + (24 * warp + 6 * g) + lda * (8 * c);
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int r = 0; r < 6; r += 2) {
            const double2 v = *reinterpret_cast<const double2 *>(base + r + lda * k);
            R.a[r][k] = v.x;
            R.a[r + 1][k] = v.y;
        }
}

// part4[64 w + j] = sum over the rows of warp w of A[i, j] v_i
__device__ __forceinline__ void rega_at_times_v(const RegA &R, const double *v, double *part4) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 3, c = lane & 7;
    const double2 *v2 = reinterpret_cast<const double2 *>(v + 24 * warp + 6 * g);
    double vr[6];
#pragma unroll
    for (int r = 0; r < 6; r += 2) {
        const double2 t = v2[r >> 1];
        vr[r] = t.x;
        vr[r + 1] = t.y;
    }
    double cp[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        double acc = 0.0;
#pragma unroll
        for (int r = 0; r < 6; ++r) acc += R.a[r][k] * vr[r];
        cp[k] = acc;
    }
    const bool hi2 = (g & 2) != 0, hi1 = (g & 1) != 0;
    double v4[4], s2[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double send = hi2 ? cp[i] : cp[i + 4], keep = hi2 ? cp[i + 4] : cp[i];
        v4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const double send = hi1 ? v4[i] : v4[i + 2], keep = hi1 ? v4[i + 2] : v4[i];
        s2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const int col = 8 * c + (hi2 ? 4 : 0) + (hi1 ? 2 : 0);
    *reinterpret_cast<double2 *>(part4 + kDN * warp + col) = make_double2(s2[0], s2[1]);
}

// (A v)_row for row = 24 w + 6 g + c, valid on the lanes with c < 6
__device__ __forceinline__ double rega_a_times_v(const RegA &R, const double *v, int &row, bool &valid) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 3, c = lane & 7;
    const double2 *v2 = reinterpret_cast<const double2 *>(v + 8 * c);
    double vc[8];
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
        const double2 t = v2[k >> 1];
        vc[k] = t.x;
        vc[k + 1] = t.y;
    }
    double rp[8];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += R.a[r][k] * vc[k];
        rp[r] = acc;
    }
    rp[6] = rp[7] = 0.0;
    const bool b4 = (c & 4) != 0, b2 = (c & 2) != 0, b1 = (c & 1) != 0;
    double v4[4], s2[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double send = b4 ? rp[i] : rp[i + 4], keep = b4 ? rp[i + 4] : rp[i];
        v4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const double send = b2 ? v4[i] : v4[i + 2], keep = b2 ? v4[i + 2] : v4[i];
        s2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    const double send = b1 ? s2[0] : s2[1], keep = b1 ? s2[1] : s2[0];
    const double out = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    row = 24 * warp + 6 * g + c;
    valid = c < 6;
    return out;
}

// K^-1 in registers as well: warp w owns rows [16 w, 16 w + 16), lane 8 g + c owns rows 16 w + 4 g + (0..3) and
// columns 8 c + (0..7); the row sums reduce across the 8 lanes of a row group (4 FP64 shuffles per thread).
struct RegK {
    double k[4][8];
};

__device__ __forceinline__ void regk_load(RegK &K, const double *Kf) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 3, c = lane & 7;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const double2 *row = reinterpret_cast<const double2 *>(Kf + pidx(16 * warp + 4 * g + r, 8 * c));
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const double2 v = row[e];
            K.k[r][2 * e] = v.x;
            K.k[r][2 * e + 1] = v.y;
        }
    }
}

// (K^-1 v)_row for row = 16 w + 4 g + (c >> 1); both lanes of a pair (c, c ^ 1) return it
__device__ __forceinline__ double regk_times_v(const RegK &K, const double *v, int &row) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 3, c = lane & 7;
    const double2 *v2 = reinterpret_cast<const double2 *>(v + 8 * c);
    double vc[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const double2 t = v2[e];
        vc[2 * e] = t.x;
        vc[2 * e + 1] = t.y;
    }
    double rp[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += K.k[r][j] * vc[j];
        rp[r] = acc;
    }
    const bool b4 = (c & 4) != 0, b2 = (c & 2) != 0;
    double s2[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const double send = b4 ? rp[i] : rp[i + 2], keep = b4 ? rp[i + 2] : rp[i];
        s2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    const double send = b2 ? s2[0] : s2[1], keep = b2 ? s2[1] : s2[0];
    double out = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    out += __shfl_xor_sync(0xffffffffu, out, 1);
    row = 16 * warp + 4 * g + (b4 ? 2 : 0) + (b2 ? 1 : 0);
    return out;
}

// block-wide max of NV values held per thread (NaN-propagating), broadcast to all threads
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

// REGA (requires MPC == 96): A lives in registers during the iterations (see RegA above); the factorisation and
// the convergence check keep using the shared-memory copy.
template <int MPC, int REG = 0>
__global__ void __launch_bounds__(kDThreads) dense_batch_kernel(DenseBatchParams p) {
    static_assert(REG == 0 || MPC == 96, "the register-resident variants are compiled for mp = 96 only");
    constexpr bool REGA = REG >= 1;   // A in registers
    constexpr bool REGK = REG >= 2;   // ... and K^-1 too
    extern __shared__ __align__(16) unsigned char raw[];
    const DenseSmem sm = carve(raw, p.mp);
    const int n = p.n, m = p.m, mp = MPC ? MPC : p.mp, lda = mp + 2;
    const int tid = threadIdx.x;
    const double alpha = p.s.alpha, alpha1 = 1.0 - alpha, sigma = p.s.sigma;
    const double eps_admm = fmin(p.s.eps_abs, p.s.eps_rel) * 1e-2;
    unsigned long long tot_iters = 0, tot_rho = 0;
    const int o = out_index(), h = out_half();

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
        // ---- load the problem into shared memory (zero padded to mp x 64, leading dimension mp + 2)
        for (int idx = tid; idx < mp * kDN; idx += kDThreads) {
            const int i = idx % mp, j = idx / mp;
            sm.As[i + lda * j] = (i < m && j < n) ? __ldg(Ag + i + (size_t)m * j) : 0.0;
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
        RegA R;
        RegK RK;
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
                const bool ok = p.blocked_chol ? chol_blocked(sm) : chol_unblocked(sm);
                fact_ok = fact_ok && ok;
                trtri_lower(sm);
                lauum_lower_and_mirror(sm);
                need_factor = false;
                if (REGA) rega_load(R, sm.As);   // (re)loaded here so that R is not live across the factorisation code
                if (REGK) {
                    __syncthreads();
                    regk_load(RK, sm.Lp);
                }
            }
            __syncthreads();
            // ---- rhs = sigma x - q + A' w,  w = rho z - y      (LinearSystemSolvers.jl:37-38 reduced)
            if (REGA) {
                rega_at_times_v(R, sm.w, sm.part4);
                __syncthreads();
                if (tid < kDN) {
                    const double s = (sm.part4[tid] + sm.part4[kDN + tid]) + (sm.part4[2 * kDN + tid] + sm.part4[3 * kDN + tid]);
                    sm.rhs[tid] = sigma * sm.x[tid] - sm.q[tid] + s;
                }
            } else {
                const double s = at_times_v<MPC>(sm.As, mp, sm.w);
                if (h == 0) sm.rhs[o] = sigma * sm.x[o] - sm.q[o] + s;
            }
            __syncthreads();
            // ---- x~ = K^-1 rhs, then the x relaxation (:57)
            double dx = 0.0, dz = 0.0;
            if (REGK) {
                int row;
                const double s = regk_times_v(RK, sm.rhs, row);
                if ((tid & 1) == 0) {
                    sm.xt[row] = s;
                    const double x_old = sm.x[row];
                    const double x_new = alpha * s + alpha1 * x_old;
                    sm.x[row] = x_new;
                    dx = fabs(x_new - x_old);
                }
            } else {
                const double s = kinv_times_v(sm.Lp, sm.rhs);
                if (h == 0) {
                    sm.xt[o] = s;
                    const double x_old = sm.x[o];
                    const double x_new = alpha * s + alpha1 * x_old;
                    sm.x[o] = x_new;
                    dx = fabs(x_new - x_old);
                }
            }
            __syncthreads();
            // ---- z~ = A x~, then the z / y update (:59-61), one row per thread
            {
                int row;
                bool valid = true;
                const double zt = REGA ? rega_a_times_v(R, sm.xt, row, valid) : a_times_v_row<MPC>(sm.As, mp, sm.xt, row);
                if (valid && row < m) {
                    const double z_old = sm.z[row], y_old = sm.y[row];
                    const double zr = alpha * zt + alpha1 * z_old;
                    const double z_new = clamp_julia(zr + rho1 * y_old, sm.l[row], sm.u[row]);   // :60
                    const double y_new = y_old + rho * (zr - z_new);                               // :61
                    sm.z[row] = z_new;
                    sm.y[row] = y_new;
                    sm.w[row] = rho * z_new - y_new;
                    dz = fabs(z_new - z_old);
                }
            }
            if (ii % p.s.check_every == 0) {
                // ---- CheckConvergence (:79-112)
                __syncthreads();
                double nr[6] = {dx, dz, 0.0, 0.0, 0.0, 0.0};   // dx dz rp max(|Ax|,|z|) rd max(|Px|,|A'y|)
                {
                    int row;
                    const double ax = a_times_v_row<MPC>(sm.As, mp, sm.x, row);
                    if (row < m) {
                        const double zi = sm.z[row];
                        nr[2] = fabs(ax - zi);
                        nr[3] = nanmax(fabs(ax), fabs(zi));
                    }
                }
                const double aty = at_times_v<MPC>(sm.As, mp, sm.y);
                double px = 0.0;
                if (o < n) {
                    const int ja = h == 0 ? 0 : (n >> 1), jb = h == 0 ? (n >> 1) : n;
                    for (int j = ja; j < jb; ++j) px += __ldg(Pg + o + (size_t)n * j) * sm.x[j];
                }
                px += __shfl_xor_sync(0xffffffffu, px, 8);
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

}  // namespace qpb
