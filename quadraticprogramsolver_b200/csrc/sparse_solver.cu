// sparse_solver.cu -- host side of the single-GPU sparse path: qpb200_create / solve / apply / destroy.
// Replaces SolveQuadraticProgram! + LinOpCgInit/LinOpCg! (see admm_kernels.cuh for the line map).
#include <chrono>
#include <cmath>
#include <cstring>
#include <mutex>
#include <new>

#include "admm_kernels.cuh"
#include "host_common.h"
#include "sparse_solver.h"

namespace qpb {

// small element-wise helpers for qpb200_apply(which = 3)
__global__ void scale_kernel(double *v, double a, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] *= a;
}
__global__ void axpy_kernel(double *y, const double *x, double a, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] += a * x[i];
}
// column sums of rs[i] * A_ij^2 out of H = [P A']: row j of H carries column j of A behind rowmid[j]
__global__ void scaled_colsq_kernel(CsrTiled H, int n, const double *rs, double *out) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int k = H.rowmid[j]; k < H.rowptr[j + 1]; ++k) acc += rs[H.col[k] - n] * (H.val[k] * H.val[k]);
        out[j] = acc;
    }
}
// L2 flush by READING a buffer larger than L2 (a writing flush would leave dirty lines whose
// write-back is then charged to the kernel being timed)
__global__ void flush_kernel(const double *buf, size_t n, double *sink) {
    double s = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += buf[i];
    if (s == 123.456) *sink = s;
}

// Bulk arrays of a tiled CSR matrix (row pointers, split points, columns, values): asynchronous on `st`, so the copy
// of one matrix overlaps the host conversion of the next; the host arrays must stay alive until `st` is synchronised.
static int upload_matrix(DeviceArena &ar, const HostCsr &M, CsrTiled &d, cudaStream_t st) {
    int *rowptr = nullptr, *rowmid = nullptr, *col = nullptr;
    double *val = nullptr;
    const size_t nnz = (size_t)M.nnz();
    QPB_CUDA(ar.alloc(&rowptr, M.ptr.size() + 16));   // the row pointers of a tile are staged in whole 16-byte groups too
    QPB_CUDA(ar.alloc(&col, nnz + 16));
    QPB_CUDA(ar.alloc(&val, nnz + 16));
    QPB_CUDA(cudaMemcpyAsync(rowptr, M.ptr.data(), M.ptr.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    QPB_CUDA(cudaMemsetAsync(rowptr + M.ptr.size(), 0, 16 * sizeof(int), st));
    if (nnz) {
        QPB_CUDA(cudaMemcpyAsync(col, M.idx.data(), nnz * sizeof(int), cudaMemcpyHostToDevice, st));
        QPB_CUDA(cudaMemcpyAsync(val, M.val.data(), nnz * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    QPB_CUDA(cudaMemsetAsync(col + nnz, 0, 16 * sizeof(int), st));       // the tile loaders read whole 16-byte groups
    QPB_CUDA(cudaMemsetAsync(val + nnz, 0, 16 * sizeof(double), st));
    if (!M.mid.empty()) {
        QPB_CUDA(ar.alloc(&rowmid, M.mid.size()));
        QPB_CUDA(cudaMemcpyAsync(rowmid, M.mid.data(), M.mid.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    }
    d.rows = M.rows;
    d.cols = M.cols;
    d.rowptr = rowptr;
    d.rowmid = rowmid;
    d.col = col;
    d.val = val;
    return QPB200_OK;
}

// The tile plan (depends on the grid size, which is known only when both matrices have been tiled).
static int upload_plan(DeviceArena &ar, const HostTiles &T, CsrTiled &d, cudaStream_t st) {
    int4 *tiles = nullptr;
    int *cta = nullptr;
    QPB_CUDA(ar.alloc(&tiles, T.tiles.size()));
    QPB_CUDA(ar.alloc(&cta, T.cta_begin.size()));
    if (!T.tiles.empty())
        QPB_CUDA(cudaMemcpyAsync(tiles, T.tiles.data(), T.tiles.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
    QPB_CUDA(cudaMemcpyAsync(cta, T.cta_begin.data(), T.cta_begin.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    d.tiles = tiles;
    d.cta_begin = cta;
    d.ntiles = (int)T.tiles.size();
    d.lpr = T.lpr;
    return QPB200_OK;
}

namespace {
struct StreamSyncGuard {   // pending async uploads must drain before the host staging arrays are released
    cudaStream_t &st;
    ~StreamSyncGuard() { if (st) cudaStreamSynchronize(st); }
};
}  // namespace

// Shared-memory opt-in plus an explicit L1 / shared-memory split: kMinCtas CTAs of the tile engine and not a byte
// more.  Left to itself the driver sizes the carve-out for as many CTAs as the registers allow -- a 48-register kernel
// then gets 5 x 41 KB of shared memory and 28 KB of L1, and the x-gathers collapse (stand-alone H pass on cfg5:
// 0.19 ms -> 0.34 ms, profiles/r2c_spmv_friendly.jsonl).
int prep_tile_kernel(const void *kernel, int *blocks_per_sm) {
    QPB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SpmvSmem)));
    int dev = 0, max_sm_smem = 0;
    QPB_CUDA(cudaGetDevice(&dev));
    QPB_CUDA(cudaDeviceGetAttribute(&max_sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
    const int need = kMinCtas * ((int)sizeof(SpmvSmem) + 1024 + 256);   // + driver-reserved KB + static shared memory
    static const int kConfigsKiB[] = {0, 8, 16, 32, 64, 100, 132, 164, 196, 228};
    int cfg = max_sm_smem;
    for (int c : kConfigsKiB)
        if (c * 1024 >= need) { cfg = std::min(cfg, c * 1024); break; }
    int pct = (int)(100.0 * cfg / max_sm_smem);   // rounded down: the driver picks the smallest split >= the request
    if (const char *e = getenv("QPB200_CARVEOUT")) pct = atoi(e);   // A/B experiments only
    QPB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    QPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kernel, kThreads, sizeof(SpmvSmem)));
    return QPB200_OK;
}
static int prep_kernel(const void *kernel, int *blocks_per_sm) { return prep_tile_kernel(kernel, blocks_per_sm); }

int SparseSolver::settings_to_dev(const qpb200_settings &s) {
    if (!(s.rho > 0.0) || !(s.sigma >= 0.0) || s.max_iter < 0 || s.check_every <= 0 || s.pcg_max_iter < 0)
        return fail(QPB200_ERR_ARG, "settings: need rho > 0, sigma >= 0, max_iter >= 0, check_every > 0, pcg_max_iter >= 0");
    if (s.lin_solver != QPB200_LINSOLVE_PCG && s.lin_solver != QPB200_LINSOLVE_CHOLESKY)
        return fail(QPB200_ERR_ARG, "settings: lin_solver must be QPB200_LINSOLVE_PCG or QPB200_LINSOLVE_CHOLESKY");
    if (created && (s.lin_solver == QPB200_LINSOLVE_CHOLESKY) != direct)
        return fail(QPB200_ERR_ARG, "qpb200_update_settings: lin_solver is fixed at create (the dense factor is allocated there)");
    if (created && (s.device != settings.device || s.reserved_i[QPB200_RSV_SCALING_ITERS] != settings.reserved_i[QPB200_RSV_SCALING_ITERS] ||
                    s.reserved_i[QPB200_RSV_DIST_MODE] != settings.reserved_i[QPB200_RSV_DIST_MODE]))
        return fail(QPB200_ERR_ARG, "qpb200_update_settings: device, scaling iterations and the multi-GPU mode are fixed at create");
    if (created && s.sigma != settings.sigma) k_valid = false;   // the dense inverse was built for the old sigma
    settings = s;
    direct = s.lin_solver == QPB200_LINSOLVE_CHOLESKY;
    AdmmSettingsDev &d = prob.s;
    d.max_iter = s.max_iter;
    d.check_every = s.check_every;
    d.pcg_max_iter = s.pcg_max_iter;
    d.eps_abs = s.eps_abs;
    d.eps_rel = s.eps_rel;
    d.rho = s.rho;
    d.sigma = s.sigma;
    d.alpha = s.alpha;
    d.rho_factor = s.rho_factor;
    d.pcg_eps = s.pcg_eps;
    d.pcg_rel_eps = s.pcg_rel_eps < 0.0 ? std::sqrt(2.220446049250313e-16) : s.pcg_rel_eps;
    d.adaptive_rho = s.adaptive_rho;
    loader = s.spmv_loader == 1 ? 0 : (s.spmv_loader == 3 ? 2 : 1);   // settings: 0 auto (= 2, TMA), 1 LDG, 2 TMA, 3 TMA pipelined
    use_pre = s.precond != QPB200_PRECOND_NONE;
    return QPB200_OK;
}

int SparseSolver::init(int64_t n64, int64_t m64, const int64_t *Pp, const int64_t *Pi, const double *Pv,
                       const int64_t *Ap, const int64_t *Ai, const double *Av, const double *q, const double *l,
                       const double *u, const qpb200_settings &s, int32_t base) {
    const auto t0 = std::chrono::steady_clock::now();
    std::string laps;
    auto tlast = t0;
    auto lap = [&](const char *name) {
        const auto now = std::chrono::steady_clock::now();
        char b[96];
        snprintf(b, sizeof(b), " %s=%.1f", name, std::chrono::duration<double, std::milli>(now - tlast).count());
        laps += b;
        tlast = now;
    };
    if (n64 <= 0 || m64 < 0 || n64 + m64 >= (int64_t(1) << 31) - 64)
        return fail(QPB200_ERR_ARG, "qpb200_create: need 0 < n, 0 <= m, n + m < 2^31 (got n=%lld m=%lld)", (long long)n64,
                    (long long)m64);
    if (base != 0 && base != 1) return fail(QPB200_ERR_ARG, "index_base must be 0 or 1");
    if (!q || (m64 > 0 && (!l || !u))) return fail(QPB200_ERR_ARG, "q, l, u must not be NULL");
    int rc = validate_csc("P", n64, n64, Pp, Pi, Pv, base);
    if (rc) return rc;
    rc = validate_csc("A", m64, n64, Ap, Ai, Av, base);
    if (rc) return rc;
    for (int64_t j = 0; j < n64; ++j)
        if (!std::isfinite(q[j])) return fail(QPB200_ERR_NONFINITE, "q[%lld] is not finite", (long long)j);
    for (int64_t i = 0; i < m64; ++i)
        if (std::isnan(l[i]) || std::isnan(u[i]) || l[i] > u[i])
            return fail(QPB200_ERR_NONFINITE, "bounds: need l[i] <= u[i], not NaN (row %lld)", (long long)i);
    lap("validate");
    rc = check_device(s.device);
    if (rc) return rc;
    QPB_CUDA(cudaGetDevice(&device));
    rc = settings_to_dev(s);
    if (rc) return rc;

    n = (int)n64;
    m = (int)m64;
    if (direct && n > kDirectMaxN)
        return fail(QPB200_ERR_ARG, "qpb200_create: lin_solver = CHOLESKY keeps a dense n x n inverse: n = %d exceeds %d, use PCG", n,
                    kDirectMaxN);
    // ---- grid limit: co-resident CTAs of the persistent kernel
    QPB_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
    int per_sm = 1 << 30, tmp = 0;
    {
        // shared-memory opt-in + occupancy of every kernel variant: once per device and process
        static std::mutex mu;
        static int cached_per_sm[64];
        std::lock_guard<std::mutex> g(mu);
        const bool cacheable = device >= 0 && device < 64;
        if (cacheable && cached_per_sm[device] > 0) {
            per_sm = cached_per_sm[device];
        } else {
            for (const void *fn : {(const void *)admm_kernel<1, false, false>, (const void *)admm_kernel<1, true, false>,
                                   (const void *)admm_kernel<1, false, true>, (const void *)admm_kernel<1, true, true>,
                                   (const void *)admm_kernel<1, false, false, true>, (const void *)polish_kernel<1>}) {
                if ((rc = prep_kernel(fn, &tmp))) return rc;
                per_sm = std::min(per_sm, tmp);
            }
            for (const void *fn : {(const void *)spmv_kernel<1, false>, (const void *)spmv_kernel<1, true>})
                if ((rc = prep_kernel(fn, &tmp))) return rc;
            if (cacheable) cached_per_sm[device] = per_sm;
        }
    }
    if (per_sm < 1) return fail(QPB200_ERR_CUDA, "persistent kernel does not fit on an SM");
    {
        const char *e = getenv("QPB200_CTAS_PER_SM");   // A/B experiments only
        per_sm = std::min(per_sm, e ? std::max(1, atoi(e)) : kMinCtas);
    }
    const int grid_max = num_sms * per_sm;

    QPB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    QPB_CUDA(cudaEventCreate(&ev0));
    QPB_CUDA(cudaEventCreate(&ev1));
    lap("kernel_prep");
    // ---- host conversion: CSC(P), CSC(A) -> H = [P A'] and CSR(A)
    const int scaling_iters = s.reserved_i[QPB200_RSV_SCALING_ITERS];
    if (scaling_iters < 0 || scaling_iters > 1000) return fail(QPB200_ERR_ARG, "settings: scaling iterations must be in [0, 1000]");
    nnzP = Pp[n] - base;
    nnzA = Ap[n] - base;
    if (nnzP + nnzA >= (int64_t(1) << 31) - 64)
        return fail(QPB200_ERR_ARG, "qpb200_create: nnz(P) + nnz(A) = %lld exceeds the int32 index range", (long long)(nnzP + nnzA));
    HostCsr A, H;
    HostTiles TH, TA;
    StreamSyncGuard drain{stream};                   // declared after H and A: runs before their buffers are released
    H.idx.want_pinned = H.val.want_pinned = A.idx.want_pinned = A.val.want_pinned = true;
    std::vector<double> qs, ls, us;
    double nq_unscaled = 0.0;
    for (int64_t j = 0; j < n64; ++j) nq_unscaled = std::fmax(nq_unscaled, std::fabs(q[j]));
    std::vector<double> dP((size_t)n, 0.0), dAA((size_t)n, 0.0);
    scaled = scaling_iters > 0;
    if (!scaled) {
        assemble_h_direct(n, m, Pp, Pi, Pv, Ap, Ai, Av, base, H, dP, dAA);
        lap("assemble_H");
        if ((rc = upload_matrix(arena, H, prob.H, stream))) return rc;   // copies while A is being transposed
        build_tiles(H, kTileFill, kTileRows, TH);
        csc_to_csr(m, n, Ap, Ai, Av, base, A);
        lap("transpose_A");
    } else {
        // ---- optional Ruiz equilibration (not in the reference; off by default): from here on P, A, A', q, l, u are
        //      the scaled problem, the termination norms are brought back to the unscaled one inside the kernel
        HostCsr P, At;
        csc_to_csr(n, n, Pp, Pi, Pv, base, P);
        csc_to_csr(m, n, Ap, Ai, Av, base, A);
        csc_as_csr_of_transpose(m, n, Ap, Ai, Av, base, At);
        lap("transposes");
        qs.assign(q, q + n);
        ruiz_equilibrate(P, A, At, qs, scaling_iters, scaling);
        ls.resize((size_t)m);
        us.resize((size_t)m);
        for (int i = 0; i < m; ++i) {
            ls[(size_t)i] = scaling.E[(size_t)i] * l[i];
            us[(size_t)i] = scaling.E[(size_t)i] * u[i];
        }
        q = qs.data();
        l = ls.data();
        u = us.data();
        lap("ruiz");
        H.rows = n;
        H.cols = n + m;
        H.ptr.resize((size_t)n + 1);
        H.mid.resize((size_t)n);
        H.idx.resize((size_t)(nnzP + nnzA));
        H.val.resize((size_t)(nnzP + nnzA));
        for (int j = 0; j <= n; ++j) H.ptr[(size_t)j] = P.ptr[(size_t)j] + At.ptr[(size_t)j];
        const int nloc = n;
        parallel_chunks(n, [&](int, int64_t j0, int64_t j1) {
            for (int64_t j = j0; j < j1; ++j) {
                int pos = H.ptr[(size_t)j];
                for (int k = P.ptr[(size_t)j]; k < P.ptr[(size_t)j + 1]; ++k) {
                    H.idx[(size_t)pos] = P.idx[(size_t)k];
                    H.val[(size_t)pos] = P.val[(size_t)k];
                    if (P.idx[(size_t)k] == j) dP[(size_t)j] += P.val[(size_t)k];
                    ++pos;
                }
                H.mid[(size_t)j] = pos;
                for (int k = At.ptr[(size_t)j]; k < At.ptr[(size_t)j + 1]; ++k) {
                    H.idx[(size_t)pos] = At.idx[(size_t)k] + nloc;
                    H.val[(size_t)pos] = At.val[(size_t)k];
                    dAA[(size_t)j] += At.val[(size_t)k] * At.val[(size_t)k];
                    ++pos;
                }
            }
        }, 4096);
        lap("assemble_H");
        if ((rc = upload_matrix(arena, H, prob.H, stream))) return rc;
        build_tiles(H, kTileFill, kTileRows, TH);
    }
    if ((rc = upload_matrix(arena, A, prob.A, stream))) return rc;
    build_tiles(A, kTileFill, kTileRows, TA);
    prob.normQ = nq_unscaled;

    int64_t want = std::max<int64_t>((int64_t)std::max(TH.tiles.size(), TA.tiles.size()),
                                     ((int64_t)std::max(n, m) + kThreads * 4 - 1) / (kThreads * 4));
    grid = (int)std::max<int64_t>(1, std::min<int64_t>(grid_max, want));
    if (const char *e = getenv("QPB200_GRID")) grid = std::max(1, std::min(grid_max, atoi(e)));   // A/B experiments only
    assign_tiles(TH, grid);
    assign_tiles(TA, grid);

    lap("tiles");
    // ---- upload
    if ((rc = upload_plan(arena, TH, prob.H, stream))) return rc;
    if ((rc = upload_plan(arena, TA, prob.A, stream))) return rc;
    prob.n = n;
    prob.m = m;
    double *dq, *dl, *du, *ddP, *ddAA;
    QPB_CUDA(arena.alloc(&dq, (size_t)n));
    QPB_CUDA(arena.alloc(&dl, (size_t)m));
    QPB_CUDA(arena.alloc(&du, (size_t)m));
    QPB_CUDA(arena.alloc(&ddP, (size_t)n));
    QPB_CUDA(arena.alloc(&ddAA, (size_t)n));
    QPB_CUDA(cudaMemcpy(dq, q, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    if (m) {
        QPB_CUDA(cudaMemcpy(dl, l, (size_t)m * sizeof(double), cudaMemcpyHostToDevice));
        QPB_CUDA(cudaMemcpy(du, u, (size_t)m * sizeof(double), cudaMemcpyHostToDevice));
    }
    QPB_CUDA(cudaMemcpy(ddP, dP.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    QPB_CUDA(cudaMemcpy(ddAA, dAA.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    prob.q = dq; prob.l = dl; prob.u = du; prob.dP = ddP; prob.dAA = ddAA;
    d_q = dq; d_l = dl; d_u = du; d_dAA = ddAA;
    prob.rs = nullptr;
    prob.Dv = prob.Dinvc = prob.Einv = nullptr;
    if (scaled) {
        std::vector<double> dinvc((size_t)n), einv((size_t)m);
        for (int j = 0; j < n; ++j) dinvc[(size_t)j] = 1.0 / (scaling.c * scaling.D[(size_t)j]);
        for (int i = 0; i < m; ++i) einv[(size_t)i] = 1.0 / scaling.E[(size_t)i];
        double *dD, *dDi, *dEi;
        QPB_CUDA(arena.alloc(&dD, (size_t)n));
        QPB_CUDA(arena.alloc(&dDi, (size_t)n));
        QPB_CUDA(arena.alloc(&dEi, (size_t)std::max(m, 1)));
        QPB_CUDA(cudaMemcpy(dD, scaling.D.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
        QPB_CUDA(cudaMemcpy(dDi, dinvc.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
        if (m) QPB_CUDA(cudaMemcpy(dEi, einv.data(), (size_t)m * sizeof(double), cudaMemcpyHostToDevice));
        prob.Dv = dD; prob.Dinvc = dDi; prob.Einv = dEi;
    }
    const size_t nm = (size_t)n + (size_t)m;
    QPB_CUDA(arena.alloc(&prob.XY, nm + 8, true));
    QPB_CUDA(arena.alloc(&prob.XG, nm + 8, true));
    QPB_CUDA(arena.alloc(&prob.UT, nm + 8, true));
    QPB_CUDA(arena.alloc(&prob.z, (size_t)m + 8, true));
    QPB_CUDA(arena.alloc(&prob.zt, (size_t)m + 8, true));
    QPB_CUDA(arena.alloc(&prob.r, (size_t)n + 8, true));
    QPB_CUDA(arena.alloc(&prob.c, (size_t)n + 8, true));
    QPB_CUDA(arena.alloc(&prob.zp, (size_t)n + 8, true));
    QPB_CUDA(arena.alloc(&prob.dinv, (size_t)n + 8, true));
    QPB_CUDA(arena.alloc(&prob.wv, (size_t)n + 8, true));
    QPB_CUDA(arena.alloc(&prob.info, 1, true));
    QPB_CUDA(arena.alloc(&sync_words, 64, true));   // count @0, flag @32 (separate 128-B lines)
    prob.gs.count = sync_words;
    prob.gs.flag = sync_words + 32;
    QPB_CUDA(arena.alloc(&prob.gs.partials[0], (size_t)grid_max * kMaxRed, true));
    QPB_CUDA(arena.alloc(&prob.gs.partials[1], (size_t)grid_max * kMaxRed, true));
    QPB_CUDA(arena.alloc(&scratch, std::max(nm, (size_t)2 * n) + 8, true));
    prob.Kneg = nullptr;
    prob.ldk = 0;
    if (direct) {
        ldk = (n + kGjTile - 1) / kGjTile * kGjTile;
        QPB_CUDA(arena.alloc(&d_K, (size_t)ldk * ldk));
        QPB_CUDA(arena.alloc(&d_gjD, (size_t)kGjNb * kGjNb));
        QPB_CUDA(arena.alloc(&d_gjW, (size_t)ldk * kGjNb));
        QPB_CUDA(arena.alloc(&d_gjC, (size_t)ldk * kGjNb));
        QPB_CUDA(arena.alloc(&d_gjStatus, 1, true));
        prob.Kneg = d_K;
        prob.ldk = ldk;
        k_valid = false;
    }
    QPB_CUDA(cudaDeviceSynchronize());
    lap("upload+alloc");
    h_l.assign(l, l + m);   // host copies of the (scaled) bounds: qpb200_update_vectors re-checks l <= u against them
    h_u.assign(u, u + m);
    created = true;
    setup_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (getenv("QPB200_TIMING")) fprintf(stderr, "[qpb200_create] total %.1f ms:%s\n", setup_ms, laps.c_str());
    return QPB200_OK;
}

SparseSolver::~SparseSolver() {
    const bool timing = getenv("QPB200_TIMING") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    auto ms_since = [&](std::chrono::steady_clock::time_point a) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count();
    };
    if (device >= 0) cudaSetDevice(device);
    if (flush_buf) cudaFree(flush_buf);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
    const double t_obj = ms_since(t0);
    const auto t1 = std::chrono::steady_clock::now();
    const size_t nblocks = arena.ptrs.size();
    arena.release();
    if (timing)
        fprintf(stderr, "[qpb200 ~SparseSolver] stream/events/flush %.1f ms, arena release (%zu blocks) %.1f ms\n", t_obj, nblocks,
                ms_since(t1));
}

int SparseSolver::reset_state(const double *x0_host) {
    const size_t nm = (size_t)n + (size_t)m;
    QPB_CUDA(cudaMemsetAsync(prob.XY, 0, nm * sizeof(double), stream));
    QPB_CUDA(cudaMemsetAsync(prob.XG, 0, nm * sizeof(double), stream));
    QPB_CUDA(cudaMemsetAsync(prob.UT, 0, nm * sizeof(double), stream));
    QPB_CUDA(cudaMemsetAsync(prob.z, 0, (size_t)m * sizeof(double) + 8, stream));
    QPB_CUDA(cudaMemsetAsync(prob.zt, 0, (size_t)m * sizeof(double) + 8, stream));
    QPB_CUDA(cudaMemsetAsync(prob.r, 0, (size_t)n * sizeof(double), stream));
    QPB_CUDA(cudaMemsetAsync(prob.c, 0, (size_t)n * sizeof(double), stream));
    QPB_CUDA(cudaMemsetAsync(prob.zp, 0, (size_t)n * sizeof(double), stream));
    QPB_CUDA(cudaMemsetAsync(sync_words, 0, 64 * sizeof(unsigned long long), stream));
    if (x0_host)
        QPB_CUDA(cudaMemcpyAsync(prob.XY, x0_host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, stream));
    return QPB200_OK;
}

bool SparseSolver::one_reduction() const {
    const int r = settings.reserved_i[QPB200_RSV_CG_RECURRENCE];
    return r == 2 || (r == 0 && nnzP + 2 * nnzA <= (int64_t)16 << 20);
}

int SparseSolver::launch_admm() {
    void *args[] = {(void *)&prob};
    const void *fns[2][2] = {{(const void *)admm_kernel<1, false, true>, (const void *)admm_kernel<1, true, true>},
                             {(const void *)admm_kernel<1, false, false>, (const void *)admm_kernel<1, true, false>}};
    const void *fn = direct ? (const void *)admm_kernel<1, false, false, true> : fns[one_reduction() ? 0 : 1][use_pre ? 1 : 0];
    QPB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kThreads), args, sizeof(SpmvSmem), stream));
    return QPB200_OK;
}

int SparseSolver::solve(double *x_inout, double *z_out, double *y_out, qpb200_info *info) {
    if (!x_inout) return fail(QPB200_ERR_ARG, "qpb200_solve: x_inout is NULL");
    if (!all_finite(x_inout, (size_t)n)) return fail(QPB200_ERR_NONFINITE, "qpb200_solve: the start point holds NaN or Inf");
    const auto wall0 = std::chrono::steady_clock::now();
    QPB_CUDA(cudaSetDevice(device));
    std::vector<double> x0s;
    if (scaled) {                                   // start point in the scaled variables: x_s = D^-1 x
        x0s.resize((size_t)n);
        for (int j = 0; j < n; ++j) x0s[(size_t)j] = x_inout[j] / scaling.D[(size_t)j];
    }
    int rc = reset_state(scaled ? x0s.data() : x_inout);
    if (rc) return rc;
    prob.iter0 = 0;
    prob.rho0 = prob.rhorho0 = settings.rho;
    prob.resume_changed = 0;
    int64_t launches = 0;
    AdmmInfoDev seg{};                              // direct path: totals over the launches of this solve
    seg.res_prim = seg.res_dual = std::nan("");
    QPB_CUDA(cudaEventRecord(ev0, stream));
    if (direct && (!k_valid || k_rho != settings.rho)) {
        if ((rc = refactor(settings.rho, &launches))) return rc;
    }
    for (;;) {
        if ((rc = launch_admm())) return rc;
        ++launches;
        if (!direct) break;
        // the exact-solve path leaves the kernel when the rho trigger fires (SolveQuadraticProgram.jl:46-52): refactorise
        // K for the new rho (LinearSystemSolvers.jl:93-95) and re-enter at the same iteration
        AdmmInfoDev part;
        QPB_CUDA(cudaMemcpyAsync(&part, prob.info, sizeof(part), cudaMemcpyDeviceToHost, stream));
        QPB_CUDA(cudaStreamSynchronize(stream));
        seg.n_h_passes += part.n_h_passes;
        seg.n_a_passes += part.n_a_passes;
        if (!std::isnan(part.res_prim)) { seg.res_prim = part.res_prim; seg.res_dual = part.res_dual; }
        if (part.conv_flag != 0) break;
        seg.rho_updates += 1;
        prob.iter0 = part.iterations;
        prob.rho0 = prob.rhorho0 = part.rho_final;
        prob.resume_changed = 1;
        QPB_CUDA(cudaMemsetAsync(sync_words, 0, 64 * sizeof(unsigned long long), stream));   // the barrier epochs restart with the launch
        if ((rc = refactor(part.rho_final, &launches))) return rc;
    }
    const bool do_polish = settings.reserved_i[QPB200_RSV_POLISH] != 0;
    if (do_polish) {
        if ((rc = polish())) return rc;
        ++launches;
    }
    QPB_CUDA(cudaEventRecord(ev1, stream));
    QPB_CUDA(cudaMemcpyAsync(x_inout, prob.XY, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, stream));
    if (do_polish) QPB_CUDA(cudaMemcpyAsync(pol_out, pol.out, sizeof(pol_out), cudaMemcpyDeviceToHost, stream));
    if (z_out && m) QPB_CUDA(cudaMemcpyAsync(z_out, prob.z, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, stream));
    if (y_out && m) QPB_CUDA(cudaMemcpyAsync(y_out, prob.XY + n, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, stream));
    AdmmInfoDev hi;
    QPB_CUDA(cudaMemcpyAsync(&hi, prob.info, sizeof(hi), cudaMemcpyDeviceToHost, stream));
    QPB_CUDA(cudaStreamSynchronize(stream));
    if (scaled) {                                   // x = D x_s, z = z_s / E, y = E y_s / c
        for (int j = 0; j < n; ++j) x_inout[j] *= scaling.D[(size_t)j];
        if (z_out)
            for (int i = 0; i < m; ++i) z_out[i] /= scaling.E[(size_t)i];
        if (y_out)
            for (int i = 0; i < m; ++i) y_out[i] = scaling.E[(size_t)i] * y_out[i] / scaling.c;
    }
    float ms = 0.f;
    QPB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    if (direct) {
        hi.n_h_passes = seg.n_h_passes;
        hi.n_a_passes = seg.n_a_passes;
        hi.rho_updates = seg.rho_updates;
        if (std::isnan(hi.res_prim)) { hi.res_prim = seg.res_prim; hi.res_dual = seg.res_dual; }
    }
    last_info = hi;
    if (info) {
        std::memset(info, 0, sizeof(*info));
        info->conv_flag = hi.conv_flag;
        info->iterations = hi.iterations;
        info->rho_final = hi.rho_final;
        info->res_prim = hi.res_prim;
        info->res_dual = hi.res_dual;
        info->rho_updates = hi.rho_updates;
        info->pcg_iters_total = hi.pcg_iters_total;
        info->pcg_maxed = hi.pcg_maxed;
        info->solve_ms = ms;
        info->setup_ms = setup_ms;
        info->kernel_launches = launches;
        if (do_polish) {
            info->polish_status = (int32_t)pol_out[0];
            info->polish_minres_iters = pol_out[1];
            info->polish_active = pol_out[2];
        }
    }
    if (getenv("QPB200_TIMING"))
        fprintf(stderr, "[qpb200_solve] device %.1f ms, wall %.1f ms\n", ms,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count());
    return QPB200_OK;
}

// Polish (polish_kernels.cuh): active sets, numItrPolish refinement rounds, MINRES on the masked KKT operator -- one
// cooperative launch behind the ADMM kernel on the same stream; x (the n-part of XY) is replaced on success.
int SparseSolver::polish() {
    const size_t N = (size_t)n + (size_t)m;
    if (!pol_ready) {
        double **bufs[] = {&pol.T, &pol.TT, &pol.G, &pol.B, &pol.V, &pol.Y[0], &pol.Y[1], &pol.Y[2], &pol.W[0], &pol.W[1], &pol.W[2]};
        for (double **b : bufs) QPB_CUDA(arena.alloc(b, N + 8, true));
        QPB_CUDA(arena.alloc(&pol.mask, (size_t)m + 8, true));
        QPB_CUDA(arena.alloc(&pol.out, 4, true));
        pol_ready = true;
    }
    pol.delta = settings.delta;
    pol.tol = settings.minres_eps;
    pol.polish_iter = settings.polish_iter;
    pol.minres_iter = settings.minres_iter;
    QPB_CUDA(cudaMemsetAsync(sync_words, 0, 64 * sizeof(unsigned long long), stream));   // barrier epochs restart with the launch
    void *args[] = {(void *)&prob, (void *)&pol};
    QPB_CUDA(cudaLaunchCooperativeKernel((const void *)polish_kernel<1>, dim3(grid), dim3(kThreads), args, sizeof(SpmvSmem), stream));
    return QPB200_OK;
}

int SparseSolver::set_rho_scale(const double *rs) {
    QPB_CUDA(cudaSetDevice(device));
    k_valid = false;                                // K = P + sigma I + A' diag(rho_i) A changes with the scale
    if (!rs) {                                      // back to the scalar rho of the reference
        prob.rs = nullptr;
        prob.dAA = d_dAA;
        return QPB200_OK;
    }
    for (int i = 0; i < m; ++i)
        if (!(rs[i] > 0.0) || !std::isfinite(rs[i])) return fail(QPB200_ERR_ARG, "qpb200_set_rho_scale: entry %d must be positive and finite", i);
    if (!d_rs) {
        QPB_CUDA(arena.alloc(&d_rs, (size_t)std::max(m, 1)));
        QPB_CUDA(arena.alloc(&d_dAA_scaled, (size_t)n));
    }
    if (m) QPB_CUDA(cudaMemcpyAsync(d_rs, rs, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, stream));
    scaled_colsq_kernel<<<std::max(1, std::min(1024, (n + 255) / 256)), 256, 0, stream>>>(prob.H, n, d_rs, d_dAA_scaled);
    QPB_CUDA(cudaGetLastError());
    QPB_CUDA(cudaStreamSynchronize(stream));         // rs is a borrowed host array
    prob.rs = d_rs;
    prob.dAA = d_dAA_scaled;
    return QPB200_OK;
}

int64_t SparseSolver::spmv_bytes(int which) const {
    // 12 nnz + 4 (rows + 1) + 8 cols + 8 rows   
    const int64_t bH = 12 * (nnzP + nnzA) + 4 * ((int64_t)n + 1) + 8 * ((int64_t)n + m) + 8 * (int64_t)n;
    const int64_t bA = 12 * nnzA + 4 * ((int64_t)m + 1) + 8 * (int64_t)n + 8 * (int64_t)m;
    switch (which) {
        case 1: return bA;
        case 0: case 2: case 5: return bH + 8 * (int64_t)n;   // split pass writes two n-vectors
        case 4: return bH;
        case 3: return bA + bH;
        default: return 0;
    }
}

int64_t SparseSolver::solve_bytes() const {
    // algorithmic bytes moved by the last solve: matrix passes + the element-wise vector passes
    const AdmmInfoDev &i = last_info;
    const int64_t bH = spmv_bytes(4), bA = spmv_bytes(1);
    // standard recurrence: S4 x~,u,r,c,(dinv) reads + x~,r,(zp) writes; S1 zp,u reads + u write = 11 n-vectors;
    // one-reduction arrangement: z,w,p,s,x~,r,(dinv) reads + p,s,x~,r,z writes = 12 (12 n-vector passes)
    const int64_t pcg_vec = 8LL * (one_reduction() ? 12 : 11) * n;
    const int64_t upd_vec = 8LL * (3 * (int64_t)n + 8 * (int64_t)m);
    const int64_t dense = direct ? 8LL * n * (int64_t)ldk + 24LL * n : 0;   // one pass over -K^-1 per ADMM iteration
    return i.n_h_passes * bH + i.n_a_passes * bA + i.pcg_iters_total * pcg_vec + i.iterations * (upd_vec + dense);
}

template <bool SPLIT>
static void launch_spmv(int loader, const CsrTiled &M, int grid, const double *x, double *y0, double *y1, cudaStream_t st) {
    (void)loader;   // one tile loader since round 2 (TMA bulk copies); settings.spmv_loader is accepted and ignored
    spmv_kernel<1, SPLIT><<<grid, kThreads, sizeof(SpmvSmem), st>>>(M, x, y0, y1);
}

int SparseSolver::apply_device(int which, const double *x, double *y) {
    // x, y device pointers; uses scratch (n+m) as needed
    switch (which) {
        case 1: launch_spmv<false>(loader, prob.A, grid, x, y, nullptr, stream); break;
        case 4: launch_spmv<false>(loader, prob.H, grid, x, y, nullptr, stream); break;
        case 5: launch_spmv<true>(loader, prob.H, grid, x, y, y + n, stream); break;
        default: return fail(QPB200_ERR_ARG, "apply_device: which = %d", which);
    }
    QPB_CUDA(cudaGetLastError());
    return QPB200_OK;
}

int SparseSolver::apply(int which, const double *x_host, double *y_host) {
    if (!x_host || !y_host) return fail(QPB200_ERR_ARG, "qpb200_apply: NULL vector");
    if (scaled) return fail(QPB200_ERR_ARG, "qpb200_apply: the handle holds the equilibrated operators; create it with scaling off");
    QPB_CUDA(cudaSetDevice(device));
    const size_t nm = (size_t)n + (size_t)m;
    double *in = prob.UT;       // borrow the (u; t) pair and the scratch pair; a solve resets them anyway
    double *out = scratch;
    QPB_CUDA(cudaMemsetAsync(in, 0, nm * sizeof(double), stream));
    int rc = 0;
    switch (which) {
        case 0:   // y = P x  (first split sum of H [x; 0])
            QPB_CUDA(cudaMemcpyAsync(in, x_host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, stream));
            if ((rc = apply_device(5, in, out))) return rc;
            QPB_CUDA(cudaMemcpyAsync(y_host, out, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, stream));
            break;
        case 1:   // y = A x
            QPB_CUDA(cudaMemcpyAsync(in, x_host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, stream));
            if ((rc = apply_device(1, in, out))) return rc;
            QPB_CUDA(cudaMemcpyAsync(y_host, out, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, stream));
            break;
        case 2:   // y = A' x  (second split sum of H [0; x])
            QPB_CUDA(cudaMemcpyAsync(in + n, x_host, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, stream));
            if ((rc = apply_device(5, in, out))) return rc;
            QPB_CUDA(cudaMemcpyAsync(y_host, out + n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, stream));
            break;
        case 3: { // y = (P + sigma I + rho A'A) x, the operator of LinearSystemSolvers.jl:152-157
            QPB_CUDA(cudaMemcpyAsync(in, x_host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, stream));
            if ((rc = apply_device(1, in, in + n))) return rc;
            scale_kernel<<<256, 256, 0, stream>>>(in + n, settings.rho, m);
            if ((rc = apply_device(4, in, out))) return rc;
            axpy_kernel<<<256, 256, 0, stream>>>(out, in, settings.sigma, n);
            QPB_CUDA(cudaGetLastError());
            QPB_CUDA(cudaMemcpyAsync(y_host, out, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, stream));
            break;
        }
        default: return fail(QPB200_ERR_ARG, "qpb200_apply: which must be 0..3");
    }
    QPB_CUDA(cudaStreamSynchronize(stream));
    return QPB200_OK;
}

int SparseSolver::time_apply(int which, int reps, int flush_l2, double *ms_out) {
    if (reps <= 0 || !ms_out) return fail(QPB200_ERR_ARG, "qpb200_time_apply: reps > 0 and ms_out required");
    QPB_CUDA(cudaSetDevice(device));
    const size_t flush_n = (size_t)48 << 20;   // 48 Mi doubles = 384 MiB > 126 MB L2
    if (flush_l2 && !flush_buf) {
        QPB_CUDA(cudaMalloc(&flush_buf, flush_n * sizeof(double)));
        QPB_CUDA(cudaMemset(flush_buf, 0, flush_n * sizeof(double)));
    }
    const size_t nm = (size_t)n + (size_t)m;
    // deterministic non-trivial input
    std::vector<double> hx(nm);
    for (size_t i = 0; i < nm; ++i) hx[i] = 1.0 + 1e-3 * (double)(i % 1000);
    QPB_CUDA(cudaMemcpyAsync(prob.UT, hx.data(), nm * sizeof(double), cudaMemcpyHostToDevice, stream));
    double total = 0.0;
    for (int r = -2; r < reps; ++r) {     // 2 untimed warm-up launches
        if (flush_l2) flush_kernel<<<2048, 256, 0, stream>>>(flush_buf, flush_n, scratch);
        QPB_CUDA(cudaEventRecord(ev0, stream));
        int rc = 0;
        if (which == 3) {
            if ((rc = apply_device(1, prob.UT, prob.UT + n))) return rc;
            if ((rc = apply_device(4, prob.UT, scratch))) return rc;
        } else if (which == 0 || which == 2 || which == 5) {
            if ((rc = apply_device(5, prob.UT, scratch))) return rc;
        } else if (which == 1 || which == 4) {
            if ((rc = apply_device(which, prob.UT, scratch))) return rc;
        } else {
            return fail(QPB200_ERR_ARG, "qpb200_time_apply: which = %d", which);
        }
        QPB_CUDA(cudaEventRecord(ev1, stream));
        QPB_CUDA(cudaStreamSynchronize(stream));
        float ms = 0.f;
        QPB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
        if (r >= 0) total += ms;
    }
    *ms_out = total / reps;
    return QPB200_OK;
}

int SparseSolver::update_vectors(const double *q, const double *l, const double *u) {
    QPB_CUDA(cudaSetDevice(device));
    std::vector<double> tmp;
    if (q) {
        double nq = 0.0;
        for (int j = 0; j < n; ++j) {
            if (!std::isfinite(q[j])) return fail(QPB200_ERR_NONFINITE, "q[%d] is not finite", j);
            nq = std::fmax(nq, std::fabs(q[j]));
        }
        prob.normQ = nq;
        if (scaled) {                               // q_s = c D q (the equilibration itself is kept)
            tmp.resize((size_t)n);
            for (int j = 0; j < n; ++j) tmp[(size_t)j] = scaling.c * scaling.D[(size_t)j] * q[j];
            q = tmp.data();
        }
        QPB_CUDA(cudaMemcpy(d_q, q, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    }
    // bounds: validated as a pair against the stored counterpart (create checks l <= u, so must every update)
    if (m && (l || u)) {
        std::vector<double> nl(h_l), nu(h_u);
        for (int which = 0; which < 2; ++which) {
            const double *b = which ? u : l;
            if (!b) continue;
            std::vector<double> &dst = which ? nu : nl;
            for (int i = 0; i < m; ++i) {
                if (std::isnan(b[i])) return fail(QPB200_ERR_NONFINITE, "bound %d is NaN", i);
                dst[(size_t)i] = scaled ? scaling.E[(size_t)i] * b[i] : b[i];
            }
        }
        for (int i = 0; i < m; ++i)
            if (nl[(size_t)i] > nu[(size_t)i])
                return fail(QPB200_ERR_NONFINITE, "bounds: need l[i] <= u[i] (row %d) after the update", i);
        if (l) QPB_CUDA(cudaMemcpy(d_l, nl.data(), (size_t)m * sizeof(double), cudaMemcpyHostToDevice));
        if (u) QPB_CUDA(cudaMemcpy(d_u, nu.data(), (size_t)m * sizeof(double), cudaMemcpyHostToDevice));
        h_l.swap(nl);
        h_u.swap(nu);
    }
    return QPB200_OK;
}

}  // namespace qpb
