// direct_solver.cu -- kernels and host driver of the dense "factorisation" behind the exact x~ step
// (settings.lin_solver = QPB200_LINSOLVE_CHOLESKY on qpb200_create); design notes in direct_kernels.cuh.
// Replaces the factorisation calls of the reference's direct plugins: ldlt / qdldl / ldl at init
// (LinearSystemSolvers.jl:18,49,81) and on every rho change (:30-32,61-63,93-95).
#include <cmath>
#include <cstring>
#include <mutex>

#include "direct_kernels.cuh"
#include "proxqp_kernels.cuh"
#include "host_common.h"
#include "sparse_solver.h"

namespace qpb {

__device__ __forceinline__ void dmma_8x8x4(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d[0]), "+d"(d[1])
                 : "d"(a), "d"(b));
}

// K = P + sigma I + A' diag(rho_i) A, dense row-major with leading dimension ld (a multiple of 64); rows / columns
// n..ld-1 are those of the identity so that the padded matrix stays SPD.  One warp per row j: row j of H = [P A'] is
// row j of P followed by column j of A; every entry (i, a_ij) of that column adds rho_i a_ij * (row i of A).
// The warp walks the column in order and its lanes own distinct columns of a row => fixed summation order.
__global__ void __launch_bounds__(256) build_k_kernel(CsrTiled H, CsrTiled A, int n, int ld, double sigma, double rho,
                                                      const double *rs, double *K) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int j = gw; j < ld; j += nw) {
        double *row = K + (size_t)j * ld;
        for (int c = lane; c < ld; c += 32) row[c] = (j >= n && c == j) ? 1.0 : 0.0;
        if (j >= n) continue;
        __syncwarp();
        for (int k = H.rowptr[j] + lane; k < H.rowmid[j]; k += 32) atomicAdd(row + H.col[k], H.val[k]);   // (duplicates allowed)
        __syncwarp();
        if (lane == 0) row[j] += sigma;
        __syncwarp();
        for (int k = H.rowmid[j]; k < H.rowptr[j + 1]; ++k) {
            const int i = H.col[k] - n;
            const double w = (rs ? rho * rs[i] : rho) * H.val[k];
            for (int e = A.rowptr[i] + lane; e < A.rowptr[i + 1]; e += 32) row[A.col[e]] += w * A.val[e];
            __syncwarp();
        }
    }
}

// D = K_kk^-1 of the 32 x 32 pivot block (unblocked sweep in shared memory); *status |= 1 on a non-positive pivot.
__global__ void __launch_bounds__(256) gj_pivot_kernel(const double *K, int ld, int kb, double *D, int *status) {
    __shared__ double a[kGjNb][kGjNb + 1];
    __shared__ double colp[kGjNb], rowp[kGjNb];
    const int t = threadIdx.x;
    const double *blk = K + (size_t)kb * kGjNb * ld + (size_t)kb * kGjNb;
    for (int e = t; e < kGjNb * kGjNb; e += 256) a[e >> 5][e & 31] = blk[(size_t)(e >> 5) * ld + (e & 31)];
    __syncthreads();
    for (int p = 0; p < kGjNb; ++p) {
        if (t < 32) colp[t] = a[t][p];
        else if (t < 64) rowp[t - 32] = a[p][t - 32];
        __syncthreads();
        const double piv = colp[p];
        if (t == 0 && !(piv > 0.0)) atomicOr(status, 1);
        const double d = 1.0 / piv;
        for (int e = t; e < kGjNb * kGjNb; e += 256) {
            const int i = e >> 5, j = e & 31;
            double v;
            if (i == p) v = (j == p) ? -d : rowp[j] * d;
            else if (j == p) v = colp[i] * d;
            else v = a[i][j] - colp[i] * d * rowp[j];
            a[i][j] = v;
        }
        __syncthreads();
    }
    for (int e = t; e < kGjNb * kGjNb; e += 256) D[e] = -a[e >> 5][e & 31];
}

// Panel of pivot block kb, 64 rows per CTA: C = K[:, kb] (old), W = C D.  Writes the scratch panels the trailing update
// reads (Wneg = -W and Cs = C, both ZERO on the pivot rows so that the update leaves the pivot rows / columns alone)
// and the new pivot column and row of K:  K[:, kb] = W, K[kb, :] = W', K[kb, kb] = -D.
__global__ void __launch_bounds__(256) gj_panel_kernel(double *K, int ld, int kb, const double *D, double *Wneg, double *Cs) {
    __shared__ double Ds[kGjNb][kGjNb + 1];
    __shared__ double Cr[kGjTile][kGjNb + 1];
    const int t = threadIdx.x;
    const int r0 = blockIdx.x * kGjTile, k0 = kb * kGjNb;
    for (int e = t; e < kGjNb * kGjNb; e += 256) Ds[e >> 5][e & 31] = D[e];
    for (int e = t; e < kGjTile * kGjNb; e += 256) Cr[e >> 5][e & 31] = K[(size_t)(r0 + (e >> 5)) * ld + k0 + (e & 31)];
    __syncthreads();
    const int r = t >> 2, c0 = (t & 3) * 8;
    double w[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int k = 0; k < kGjNb; ++k) {
        const double cv = Cr[r][k];
#pragma unroll
        for (int c = 0; c < 8; ++c) w[c] += cv * Ds[k][c0 + c];
    }
    const int rg = r0 + r;
    const bool pivot_row = rg >= k0 && rg < k0 + kGjNb;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const size_t s = (size_t)rg * kGjNb + c0 + c;
        if (pivot_row) {
            Wneg[s] = 0.0;
            Cs[s] = 0.0;
            K[(size_t)rg * ld + k0 + c0 + c] = -Ds[rg - k0][c0 + c];
        } else {
            Wneg[s] = -w[c];
            Cs[s] = Cr[r][c0 + c];
            K[(size_t)rg * ld + k0 + c0 + c] = w[c];
            K[(size_t)(k0 + c0 + c) * ld + rg] = w[c];
        }
    }
}

// Trailing update of one 64 x 64 tile (ti >= tj): K_ij += Wneg_i Cs_j' on the FP64 tensor pipe; the mirror tile K_ji is
// written through shared memory (coalesced).  8 warps: warp (wm, wn) owns rows 16 wm.., columns 32 wn.. of the tile =
// 2 x 4 DMMA accumulators; k = 32 in 8 steps of 4.
__global__ void __launch_bounds__(256) gj_update_kernel(double *K, int ld, const double *Wneg, const double *Cs) {
    const int ti = blockIdx.y, tj = blockIdx.x;
    if (tj > ti) return;
    __shared__ __align__(16) double sm[2 * kGjTile * kGjLds];    // Ws | Cb, later the 64 x 65 transpose buffer
    double *Ws = sm, *Cb = sm + kGjTile * kGjLds;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int e = t; e < kGjTile * kGjNb; e += 256) {
        const int r = e >> 5, c = e & 31;
        Ws[r * kGjLds + c] = Wneg[(size_t)(ti * kGjTile + r) * kGjNb + c];
        Cb[r * kGjLds + c] = Cs[(size_t)(tj * kGjTile + r) * kGjNb + c];
    }
    const int wm = warp >> 1, wn = warp & 1;
    const int fr = lane >> 2, fc = lane & 3;
    double acc[2][4][2];
    double *tile = K + (size_t)ti * kGjTile * ld + (size_t)tj * kGjTile;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const double2 v = *reinterpret_cast<const double2 *>(tile + (size_t)(16 * wm + 8 * mt + fr) * ld + 32 * wn + 8 * nt + 2 * fc);
            acc[mt][nt][0] = v.x;
            acc[mt][nt][1] = v.y;
        }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < kGjNb / 4; ++ks) {
        double a[2], b[4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) a[mt] = Ws[(16 * wm + 8 * mt + fr) * kGjLds + 4 * ks + fc];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) b[nt] = Cb[(32 * wn + 8 * nt + fr) * kGjLds + 4 * ks + fc];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) dmma_8x8x4(acc[mt][nt], a[mt], b[nt]);
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
            *reinterpret_cast<double2 *>(tile + (size_t)(16 * wm + 8 * mt + fr) * ld + 32 * wn + 8 * nt + 2 * fc) =
                make_double2(acc[mt][nt][0], acc[mt][nt][1]);
    if (ti == tj) return;
    __syncthreads();                                  // the panels in shared memory are dead: reuse as T[64][65]
    constexpr int kT = kGjTile + 1;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int r = 16 * wm + 8 * mt + fr, c = 32 * wn + 8 * nt + 2 * fc;
            sm[r * kT + c] = acc[mt][nt][0];
            sm[r * kT + c + 1] = acc[mt][nt][1];
        }
    __syncthreads();
    double *mirror = K + (size_t)tj * kGjTile * ld + (size_t)ti * kGjTile;
    for (int e = t; e < kGjTile * kGjTile; e += 256) {
        const int c = e >> 6, r = e & 63;             // mirror row c (a column of the tile), consecutive r: coalesced
        mirror[(size_t)c * ld + r] = sm[r * kT + c];
    }
}

// Dense K = P + sigma I + A' diag(rho_i) A and its in-place inversion (direct_kernels.cuh): the "factorisation" of
// the exact-solve plugins (LinearSystemSolvers.jl:18,49,81 at init; :30-32,61-63,93-95 on a rho change).
int SparseSolver::refactor(double rho, int64_t *launches) {
    QPB_CUDA(cudaMemsetAsync(d_gjStatus, 0, sizeof(int), stream));
    const int T = ldk / kGjTile;
    build_k_kernel<<<std::max(1, std::min(4 * num_sms, (ldk + 7) / 8)), 256, 0, stream>>>(prob.H, prob.A, n, ldk, settings.sigma, rho,
                                                                                          prob.rs, d_K);
    const int nblk = (n + kGjNb - 1) / kGjNb;       // the padding beyond n is the identity: nothing to sweep there
    for (int kb = 0; kb < nblk; ++kb) {
        gj_pivot_kernel<<<1, 256, 0, stream>>>(d_K, ldk, kb, d_gjD, d_gjStatus);
        gj_panel_kernel<<<T, 256, 0, stream>>>(d_K, ldk, kb, d_gjD, d_gjW, d_gjC);
        gj_update_kernel<<<dim3(T, T), 256, 0, stream>>>(d_K, ldk, d_gjW, d_gjC);
    }
    QPB_CUDA(cudaGetLastError());
    int status = 0;
    QPB_CUDA(cudaMemcpyAsync(&status, d_gjStatus, sizeof(int), cudaMemcpyDeviceToHost, stream));
    QPB_CUDA(cudaStreamSynchronize(stream));
    if (launches) *launches += 1 + 3LL * nblk;
    if (status) {
        k_valid = false;
        return fail(QPB200_ERR_FACTOR, "direct solve: P + sigma I + rho A'A is not positive definite (non-positive pivot), rho = %g", rho);
    }
    k_valid = true;
    k_rho = rho;
    return QPB200_OK;
}

// ProxQP front end (proxqp_kernels.cuh) on an exact-solve handle whose constraint matrix is [A; C].
int SparseSolver::solve_proxqp(int64_t m_eq, const qpb200_settings &ps, double *x, double *y, double *z, double *sl,
                               bool init_slack, qpb200_proxqp_report *report) {
    if (!direct) return fail(QPB200_ERR_ARG, "qpb200_proxqp_solve: the handle must be created with lin_solver = QPB200_LINSOLVE_CHOLESKY");
    if (scaled || prob.rs) return fail(QPB200_ERR_ARG, "qpb200_proxqp_solve: equilibration / per-constraint rho are not part of this solver");
    if (m_eq < 0 || m_eq > m) return fail(QPB200_ERR_ARG, "qpb200_proxqp_solve: m_eq = %lld outside [0, %d]", (long long)m_eq, m);
    if (!x || (m_eq > 0 && !y) || (m_eq < m && (!z || (!sl && !init_slack)))) return fail(QPB200_ERR_ARG, "qpb200_proxqp_solve: NULL vector");
    if (!(ps.rho > 0.0) || !(ps.sigma >= 0.0) || ps.max_iter < 0 || ps.check_every <= 0 || !(ps.rho_factor > 0.0))
        return fail(QPB200_ERR_ARG, "qpb200_proxqp_solve: need rho > 0, sigma >= 0, max_iter >= 0, check_every > 0, tau > 0");
    const int m_in = m - (int)m_eq;
    if (!all_finite(x, (size_t)n) || (m_eq && !all_finite(y, (size_t)m_eq)) || (m_in && !all_finite(z, (size_t)m_in)) ||
        (sl && !init_slack && m_in && !all_finite(sl, (size_t)m_in)))
        return fail(QPB200_ERR_NONFINITE, "qpb200_proxqp_solve: a start vector holds NaN or Inf");
    for (int i = 0; i < m; ++i)
        if (!std::isfinite(h_u[(size_t)i])) return fail(QPB200_ERR_ARG, "qpb200_proxqp_solve: b and d (the handle's u) must be finite (row %d)", i);
    QPB_CUDA(cudaSetDevice(device));
    static std::mutex mu;
    static bool prepped[64] = {false};
    {
        std::lock_guard<std::mutex> gl(mu);
        if (device >= 0 && device < 64 && !prepped[device]) {
            int per_sm = 0;
            if (int prc = prep_tile_kernel((const void *)proxqp_kernel<1>, &per_sm)) return prc;
            if (per_sm * num_sms < grid) return fail(QPB200_ERR_CUDA, "proxqp kernel cannot be co-resident at grid %d", grid);
            prepped[device] = true;
        }
    }
    if (ps.sigma != settings.sigma) k_valid = false;   // M = P + sigma I + rho Abar'Abar
    settings.sigma = ps.sigma;
    AdmmSettingsDev &dv = prob.s;
    dv.max_iter = ps.max_iter; dv.check_every = ps.check_every; dv.eps_abs = ps.eps_abs; dv.eps_rel = ps.eps_rel;
    dv.sigma = ps.sigma; dv.adaptive_rho = ps.adaptive_rho; dv.rho = ps.rho;
    int rc = reset_state(nullptr);
    if (rc) return rc;
    QPB_CUDA(cudaMemcpyAsync(prob.XG, x, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, stream));
    if (m_eq) QPB_CUDA(cudaMemcpyAsync(prob.zt, y, (size_t)m_eq * sizeof(double), cudaMemcpyHostToDevice, stream));
    if (m_in) QPB_CUDA(cudaMemcpyAsync(prob.zt + m_eq, z, (size_t)m_in * sizeof(double), cudaMemcpyHostToDevice, stream));
    if (m_in && sl && !init_slack) QPB_CUDA(cudaMemcpyAsync(prob.z + m_eq, sl, (size_t)m_in * sizeof(double), cudaMemcpyHostToDevice, stream));
    static_assert(sizeof(ProxInfoDev) <= sizeof(AdmmInfoDev), "the report block reuses the ADMM info block");
    ProxInfoDev *d_info = reinterpret_cast<ProxInfoDev *>(prob.info);
    ProxDev pd{};
    pd.m_eq = (int)m_eq;
    pd.init_s = init_slack ? 1 : 0;
    pd.tau = ps.rho_factor;
    pd.info = d_info;
    pd.carry = ProxInfoDev{};
    pd.carry.res_prim = pd.carry.res_dual = INFINITY;
    prob.iter0 = 0;
    prob.rho0 = prob.rhorho0 = ps.rho;
    prob.resume_changed = 0;
    int64_t launches = 0;
    QPB_CUDA(cudaEventRecord(ev0, stream));
    if (!k_valid || k_rho != ps.rho)
        if ((rc = refactor(ps.rho, &launches))) return rc;
    ProxInfoDev hi{};
    for (;;) {
        QPB_CUDA(cudaMemsetAsync(sync_words, 0, 64 * sizeof(unsigned long long), stream));
        void *args[] = {(void *)&prob, (void *)&pd};
        QPB_CUDA(cudaLaunchCooperativeKernel((const void *)proxqp_kernel<1>, dim3(grid), dim3(kThreads), args, sizeof(SpmvSmem), stream));
        ++launches;
        QPB_CUDA(cudaMemcpyAsync(&hi, d_info, sizeof(hi), cudaMemcpyDeviceToHost, stream));
        QPB_CUDA(cudaStreamSynchronize(stream));
        if (hi.status == 0) break;
        pd.carry = hi;                               // UpdateDecomposition! for the new rho (ProxQP.jl:159-165), then go on
        prob.iter0 = hi.iterations_done;
        prob.rho0 = prob.rhorho0 = hi.rho;
        prob.resume_changed = 1;
        if ((rc = refactor(hi.rho, &launches))) return rc;
    }
    QPB_CUDA(cudaEventRecord(ev1, stream));
    QPB_CUDA(cudaMemcpyAsync(x, prob.XG, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, stream));
    if (m_eq) QPB_CUDA(cudaMemcpyAsync(y, prob.zt, (size_t)m_eq * sizeof(double), cudaMemcpyDeviceToHost, stream));
    if (m_in) QPB_CUDA(cudaMemcpyAsync(z, prob.zt + m_eq, (size_t)m_in * sizeof(double), cudaMemcpyDeviceToHost, stream));
    if (m_in && sl) QPB_CUDA(cudaMemcpyAsync(sl, prob.z + m_eq, (size_t)m_in * sizeof(double), cudaMemcpyDeviceToHost, stream));
    QPB_CUDA(cudaStreamSynchronize(stream));
    float ms = 0.f;
    QPB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    if (report) {
        std::memset(report, 0, sizeof(*report));
        report->converged = hi.converged;
        report->iterations = hi.conv_iter > 0 ? hi.conv_iter : ps.max_iter;   // ProxQP.jl:125,156
        report->rho = hi.rho;
        report->sigma = ps.sigma;
        report->res_prim = hi.res_prim;
        report->res_dual = hi.res_dual;
        report->rho_updates = hi.rho_updates;
        report->solve_ms = ms;
        report->kernel_launches = launches;
    }
    return QPB200_OK;
}

}  // namespace qpb
