// dense_host.cu -- host side of the batched dense path (qpb200_batch_*); the kernel is in dense_kernel.cuh.
#include <chrono>
#include <cmath>
#include <cstring>
#include <new>

#include "dense_kernel.cuh"
#include "dense_shared_kernel.cuh"
#include "host_common.h"

namespace qpb {

// default kernel for the mp = 96 shape: flipped by measurement (profiles/), both stay selectable
constexpr int kDenseRegDefault = 2;   // A and K^-1 in registers; 16 384 QPs: 68.4 (smem) -> 50.5 (A) -> 48.1 ms (A, K^-1): profiles/r1d_dense_variant_ab2.jsonl

struct DenseBatch {
    int64_t batch = 0;
    int n = 0, m = 0, mp = 0, device = -1, grid = 0, grid_cap = 0;
    qpb200_settings settings{};
    DenseBatchParams prm{};
    DeviceArena arena;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double setup_ms = 0.0;
    size_t smem = 0;
    int reg = 0;                 // mp = 96: 0 = products out of shared memory, 1 = A in registers, 2 = A and K^-1 in registers
    bool shared = false;         // one (P, A) for the whole batch: dense_shared_kernel.cuh
    DenseSharedParams sprm{};
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

// ---- pieces shared by qpb200_batch_create (everything resident) and qpb200_batch_solve_once (chunk pipeline) ------
static int batch_check_args(int64_t batch, int64_t n, int64_t m, const double *P, const double *A, const double *q, const double *l,
                            const double *u, const qpb200_settings *settings, qpb200_settings &s) {
    if (batch <= 0 || n <= 0 || n > kDN || m <= 0 || m > 128)
        return fail(QPB200_ERR_ARG, "qpb200_batch_create: need batch > 0, 0 < n <= 64, 0 < m <= 128 (got %lld, %lld, %lld)",
                    (long long)batch, (long long)n, (long long)m);
    if (batch >= (int64_t(1) << 31)) return fail(QPB200_ERR_ARG, "qpb200_batch_create: batch too large");
    if (!P || !A || !q || !l || !u) return fail(QPB200_ERR_ARG, "qpb200_batch_create: NULL array");
    if (settings) s = *settings;
    else { qpb200_default_settings(&s); s.lin_solver = QPB200_LINSOLVE_CHOLESKY; }
    if (!(s.rho > 0.0) || !(s.sigma >= 0.0) || s.max_iter < 0 || s.check_every <= 0)
        return fail(QPB200_ERR_ARG, "settings: need rho > 0, sigma >= 0, max_iter >= 0, check_every > 0");
    if (s.lin_solver != QPB200_LINSOLVE_CHOLESKY)
        return fail(QPB200_ERR_ARG, "qpb200_batch_create: the dense batch path implements lin_solver = QPB200_LINSOLVE_CHOLESKY only");
    if (s.reserved_i[QPB200_RSV_SCALING_ITERS] != 0)
        return fail(QPB200_ERR_ARG, "qpb200_batch_create: equilibration is implemented for the sparse single-GPU path only");
    const int variant = s.reserved_i[QPB200_RSV_DENSE_VARIANT];
    if (variant < 0 || variant > 3 || (variant >= 2 && (((int)m + 3) & ~3) != 96))
        return fail(QPB200_ERR_ARG, "qpb200_batch_create: dense variant %d is not available for m = %lld", variant, (long long)m);
    // value checks on a strided sample would miss entries: everything is scanned -- P and A (GBs) in the same pass
    // that stages them for the upload (staged_upload), the small vectors here
    if (!all_finite(q, (size_t)batch * n)) return fail(QPB200_ERR_NONFINITE, "q has a non-finite entry");
    for (size_t i = 0; i < (size_t)batch * m; ++i)
        if (std::isnan(l[i]) || std::isnan(u[i]) || l[i] > u[i]) return fail(QPB200_ERR_NONFINITE, "bounds: need l <= u, not NaN (entry %zu)", i);
    return check_device(s.device);
}

// device buffers for `cap` problems, stream, events, kernel attributes, grid
static int batch_setup(DenseBatch &B, int64_t cap, int64_t n, int64_t m, const qpb200_settings &s) {
    QPB_CUDA(cudaGetDevice(&B.device));
    B.batch = cap; B.n = (int)n; B.m = (int)m; B.mp = ((int)m + 3) & ~3;
    B.settings = s;
    double *dP, *dA, *dq, *dl, *du;
    QPB_CUDA(B.arena.alloc(&dP, (size_t)cap * n * n));
    QPB_CUDA(B.arena.alloc(&dA, (size_t)cap * m * n));
    QPB_CUDA(B.arena.alloc(&dq, (size_t)cap * n));
    QPB_CUDA(B.arena.alloc(&dl, (size_t)cap * m));
    QPB_CUDA(B.arena.alloc(&du, (size_t)cap * m));
    QPB_CUDA(B.arena.alloc(&B.prm.X, (size_t)cap * n));
    QPB_CUDA(B.arena.alloc(&B.prm.flags, (size_t)cap));
    QPB_CUDA(B.arena.alloc(&B.prm.iters, (size_t)cap));
    QPB_CUDA(B.arena.alloc(&B.prm.factor_fail, 1, true));
    QPB_CUDA(B.arena.alloc(&B.prm.totals, 4, true));
    QPB_CUDA(B.arena.alloc(&B.prm.queue, 4, true));
    QPB_CUDA(cudaStreamCreateWithFlags(&B.stream, cudaStreamNonBlocking));
    QPB_CUDA(cudaEventCreate(&B.ev0));
    QPB_CUDA(cudaEventCreate(&B.ev1));
    B.prm.batch = (int)cap; B.prm.n = B.n; B.prm.m = B.m; B.prm.mp = B.mp;
    B.prm.P = dP; B.prm.A = dA; B.prm.q = dq; B.prm.l = dl; B.prm.u = du;
    B.prm.blocked_chol = s.reserved_i[0] == 1 ? 0 : 1;     // reserved_i[0] = 1 selects the unblocked factor (A/B testing)
    AdmmSettingsDev &d = B.prm.s;
    d.max_iter = s.max_iter; d.check_every = s.check_every; d.pcg_max_iter = 0;
    d.eps_abs = s.eps_abs; d.eps_rel = s.eps_rel; d.rho = s.rho; d.sigma = s.sigma; d.alpha = s.alpha;
    d.rho_factor = s.rho_factor; d.pcg_eps = 0; d.pcg_rel_eps = 0; d.adaptive_rho = s.adaptive_rho;
    B.smem = dense_smem_bytes(B.mp);
    // reserved_i[QPB200_RSV_DENSE_VARIANT]: 0 = default, 1 = shared-memory products, 2 = A in registers, 3 = A and K^-1
    const int variant = s.reserved_i[QPB200_RSV_DENSE_VARIANT];
    B.reg = B.mp != 96 ? 0 : (variant == 0 ? kDenseRegDefault : variant - 1);
    const void *kfn = B.mp != 96 ? (const void *)dense_batch_kernel<0>
                      : B.reg == 2 ? (const void *)dense_batch_kernel<96, 2>
                      : B.reg == 1 ? (const void *)dense_batch_kernel<96, 1> : (const void *)dense_batch_kernel<96, 0>;
    QPB_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B.smem));
    int per_sm = 0;
    QPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, kDThreads, B.smem));
    if (per_sm < 1) return fail(QPB200_ERR_CUDA, "dense batch kernel does not fit on an SM (smem %zu)", B.smem);
    int num_sms = 0;
    QPB_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, B.device));
    B.grid_cap = num_sms * per_sm;
    B.grid = (int)std::min<int64_t>(cap, (int64_t)B.grid_cap);
    return QPB200_OK;
}

// P and A of `count` problems through the page-locked staging ring (NaN/Inf scan in the same pass), queued on B.stream
static int batch_upload_matrices(DenseBatch &B, int64_t count, const double *P, const double *A) {
    bool finiteP = true, finiteA = true;
    QPB_CUDA(staged_upload(const_cast<double *>(B.prm.P), P, (size_t)count * B.n * B.n, B.stream, &finiteP));
    QPB_CUDA(staged_upload(const_cast<double *>(B.prm.A), A, (size_t)count * B.m * B.n, B.stream, &finiteA));
    if (!finiteP || !finiteA) {
        cudaStreamSynchronize(B.stream);
        return fail(QPB200_ERR_NONFINITE, "%s has a non-finite entry", finiteP ? "A" : "P");
    }
    return QPB200_OK;
}

static int batch_launch(DenseBatch &B, int64_t count) {
    if (B.shared) {
        QPB_CUDA(cudaMemsetAsync(B.prm.totals, 0, 4 * sizeof(unsigned long long), B.stream));
        QPB_CUDA(cudaMemsetAsync(B.prm.queue, 0, 4 * sizeof(unsigned int), B.stream));
        QPB_CUDA(cudaEventRecord(B.ev0, B.stream));
        dense_shared_kernel<<<B.grid, kShThreads, B.smem, B.stream>>>(B.sprm);
        QPB_CUDA(cudaGetLastError());
        QPB_CUDA(cudaEventRecord(B.ev1, B.stream));
        return QPB200_OK;
    }
    B.prm.batch = (int)count;
    B.grid = (int)std::min<int64_t>(count, (int64_t)B.grid_cap);
    QPB_CUDA(cudaMemsetAsync(B.prm.totals, 0, 4 * sizeof(unsigned long long), B.stream));
    QPB_CUDA(cudaMemsetAsync(B.prm.factor_fail, 0, sizeof(int), B.stream));
    QPB_CUDA(cudaMemsetAsync(B.prm.queue, 0, 4 * sizeof(unsigned int), B.stream));
    QPB_CUDA(cudaEventRecord(B.ev0, B.stream));
    if (B.mp == 96 && B.reg == 2) dense_batch_kernel<96, 2><<<B.grid, kDThreads, B.smem, B.stream>>>(B.prm);
    else if (B.mp == 96 && B.reg == 1) dense_batch_kernel<96, 1><<<B.grid, kDThreads, B.smem, B.stream>>>(B.prm);
    else if (B.mp == 96) dense_batch_kernel<96, 0><<<B.grid, kDThreads, B.smem, B.stream>>>(B.prm);   // configs[2] shape: compile-time loops
    else dense_batch_kernel<0><<<B.grid, kDThreads, B.smem, B.stream>>>(B.prm);
    QPB_CUDA(cudaGetLastError());
    QPB_CUDA(cudaEventRecord(B.ev1, B.stream));
    return QPB200_OK;
}

int qpb200_batch_create(qpb200_batch **out, int64_t batch, int64_t n, int64_t m, const double *P, const double *A,
                        const double *q, const double *l, const double *u, const qpb200_settings *settings) {
    if (!out) return fail(QPB200_ERR_ARG, "qpb200_batch_create: out is NULL");
    *out = nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    qpb200_settings s;
    int rc = batch_check_args(batch, n, m, P, A, q, l, u, settings, s);
    if (rc) return rc;
    qpb200_batch *h = new (std::nothrow) qpb200_batch();
    if (!h) return fail(QPB200_ERR_ARG, "out of host memory");
    DenseBatch &B = h->b;
    rc = batch_setup(B, batch, n, m, s);
    if (rc == QPB200_OK) rc = batch_upload_matrices(B, batch, P, A);
    if (rc == QPB200_OK) {
        cudaError_t e = cudaMemcpyAsync(const_cast<double *>(B.prm.q), q, (size_t)batch * n * sizeof(double), cudaMemcpyHostToDevice, B.stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(const_cast<double *>(B.prm.l), l, (size_t)batch * m * sizeof(double), cudaMemcpyHostToDevice, B.stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(const_cast<double *>(B.prm.u), u, (size_t)batch * m * sizeof(double), cudaMemcpyHostToDevice, B.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(B.stream);
        if (e != cudaSuccess) rc = fail(QPB200_ERR_CUDA, "qpb200_batch_create: vector upload failed: %s", cudaGetErrorString(e));
    }
    if (rc != QPB200_OK) {
        delete h;
        return rc;
    }
    B.setup_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    *out = h;
    return QPB200_OK;
}

// A batch whose problems share P and A (MPC-style: q, l, u differ): one K^-1 for the batch, 16 problems per CTA as the
// columns of FP64 tensor-pipe GEMMs (dense_shared_kernel.cuh).  The handle works with qpb200_batch_solve /
// qpb200_batch_update_vectors / qpb200_batch_destroy like any other.
int qpb200_batch_create_shared(qpb200_batch **out, int64_t batch, int64_t n, int64_t m, const double *P, const double *A,
                               const double *q, const double *l, const double *u, const qpb200_settings *settings) {
    if (!out) return fail(QPB200_ERR_ARG, "qpb200_batch_create_shared: out is NULL");
    *out = nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    if (batch <= 0 || batch >= (int64_t(1) << 31) || n <= 0 || n > kDN || m <= 0 || m > 96)
        return fail(QPB200_ERR_ARG, "qpb200_batch_create_shared: need batch > 0, 0 < n <= 64, 0 < m <= 96 (got %lld, %lld, %lld)",
                    (long long)batch, (long long)n, (long long)m);
    if (!P || !A || !q || !l || !u) return fail(QPB200_ERR_ARG, "qpb200_batch_create_shared: NULL array");
    qpb200_settings s;
    if (settings) s = *settings;
    else { qpb200_default_settings(&s); s.lin_solver = QPB200_LINSOLVE_CHOLESKY; }
    if (!(s.rho > 0.0) || !(s.sigma >= 0.0) || s.max_iter < 0 || s.check_every <= 0)
        return fail(QPB200_ERR_ARG, "settings: need rho > 0, sigma >= 0, max_iter >= 0, check_every > 0");
    if (s.lin_solver != QPB200_LINSOLVE_CHOLESKY || s.adaptive_rho || s.reserved_i[QPB200_RSV_SCALING_ITERS] != 0)
        return fail(QPB200_ERR_ARG, "qpb200_batch_create_shared: needs lin_solver = CHOLESKY, adaptive_rho = 0 (one factor serves the "
                                    "whole batch only while every problem keeps the same rho) and no equilibration");
    if (s.max_iter >= (int64_t(1) << 30)) return fail(QPB200_ERR_ARG, "qpb200_batch_create_shared: max_iter too large");
    if (!all_finite(P, (size_t)n * n) || !all_finite(A, (size_t)m * n)) return fail(QPB200_ERR_NONFINITE, "P or A has a non-finite entry");
    if (!all_finite(q, (size_t)batch * n)) return fail(QPB200_ERR_NONFINITE, "q has a non-finite entry");
    for (size_t i = 0; i < (size_t)batch * m; ++i)
        if (std::isnan(l[i]) || std::isnan(u[i]) || l[i] > u[i]) return fail(QPB200_ERR_NONFINITE, "bounds: need l <= u, not NaN (entry %zu)", i);
    int rc = check_device(s.device);
    if (rc) return rc;
    qpb200_batch *h = new (std::nothrow) qpb200_batch();
    if (!h) return fail(QPB200_ERR_ARG, "out of host memory");
    DenseBatch &B = h->b;
    auto body = [&]() -> int {
        QPB_CUDA(cudaGetDevice(&B.device));
        B.batch = batch; B.n = (int)n; B.m = (int)m; B.mp = ((int)m + 7) & ~7;
        B.settings = s;
        B.shared = true;
        double *dP, *dA, *dK, *dq, *dl, *du;
        QPB_CUDA(B.arena.alloc(&dP, (size_t)n * n));
        QPB_CUDA(B.arena.alloc(&dA, (size_t)m * n));
        QPB_CUDA(B.arena.alloc(&dK, (size_t)kDN * kDN));
        QPB_CUDA(B.arena.alloc(&dq, (size_t)batch * n));
        QPB_CUDA(B.arena.alloc(&dl, (size_t)batch * m));
        QPB_CUDA(B.arena.alloc(&du, (size_t)batch * m));
        QPB_CUDA(B.arena.alloc(&B.prm.X, (size_t)batch * n));
        QPB_CUDA(B.arena.alloc(&B.prm.flags, (size_t)batch));
        QPB_CUDA(B.arena.alloc(&B.prm.iters, (size_t)batch));
        QPB_CUDA(B.arena.alloc(&B.prm.factor_fail, 1, true));
        QPB_CUDA(B.arena.alloc(&B.prm.totals, 4, true));
        QPB_CUDA(B.arena.alloc(&B.prm.queue, 4, true));
        QPB_CUDA(cudaStreamCreateWithFlags(&B.stream, cudaStreamNonBlocking));
        QPB_CUDA(cudaEventCreate(&B.ev0));
        QPB_CUDA(cudaEventCreate(&B.ev1));
        QPB_CUDA(cudaMemcpyAsync(dP, P, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice, B.stream));
        QPB_CUDA(cudaMemcpyAsync(dA, A, (size_t)m * n * sizeof(double), cudaMemcpyHostToDevice, B.stream));
        QPB_CUDA(cudaMemcpyAsync(dq, q, (size_t)batch * n * sizeof(double), cudaMemcpyHostToDevice, B.stream));
        QPB_CUDA(cudaMemcpyAsync(dl, l, (size_t)batch * m * sizeof(double), cudaMemcpyHostToDevice, B.stream));
        QPB_CUDA(cudaMemcpyAsync(du, u, (size_t)batch * m * sizeof(double), cudaMemcpyHostToDevice, B.stream));
        B.prm.batch = (int)batch; B.prm.n = B.n; B.prm.m = B.m; B.prm.mp = B.mp;
        B.prm.P = dP; B.prm.A = dA; B.prm.q = dq; B.prm.l = dl; B.prm.u = du;
        // ---- the one factorisation of the batch
        const int mp4 = ((int)m + 3) & ~3;
        const size_t fsm = dense_smem_bytes(mp4);
        QPB_CUDA(cudaFuncSetAttribute((const void *)dense_shared_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
        dense_shared_factor_kernel<<<1, kDThreads, fsm, B.stream>>>(dP, dA, B.n, B.m, mp4, s.rho, s.sigma, dK, B.prm.factor_fail);
        QPB_CUDA(cudaGetLastError());
        int ffail = 0;
        QPB_CUDA(cudaMemcpyAsync(&ffail, B.prm.factor_fail, sizeof(int), cudaMemcpyDeviceToHost, B.stream));
        QPB_CUDA(cudaStreamSynchronize(B.stream));
        if (ffail) return fail(QPB200_ERR_FACTOR, "Cholesky breakdown: a pivot of P + sigma I + rho A'A was not positive");
        DenseSharedParams &sp = B.sprm;
        sp.batch = (int)batch; sp.n = B.n; sp.m = B.m; sp.mp = B.mp;
        sp.P = dP; sp.A = dA; sp.Kinv = dK; sp.q = dq; sp.l = dl; sp.u = du;
        sp.X = B.prm.X; sp.flags = B.prm.flags; sp.iters = B.prm.iters; sp.totals = B.prm.totals; sp.queue = B.prm.queue;
        AdmmSettingsDev &d = sp.s;
        d.max_iter = s.max_iter; d.check_every = s.check_every; d.pcg_max_iter = 0;
        d.eps_abs = s.eps_abs; d.eps_rel = s.eps_rel; d.rho = s.rho; d.sigma = s.sigma; d.alpha = s.alpha;
        d.rho_factor = s.rho_factor; d.pcg_eps = 0; d.pcg_rel_eps = 0; d.adaptive_rho = 0;
        B.prm.s = d;
        B.smem = dense_shared_smem_bytes(B.mp);
        QPB_CUDA(cudaFuncSetAttribute((const void *)dense_shared_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B.smem));
        int per_sm = 0, num_sms = 0;
        QPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)dense_shared_kernel, kShThreads, B.smem));
        if (per_sm < 1) return fail(QPB200_ERR_CUDA, "shared-matrix batch kernel does not fit on an SM (smem %zu)", B.smem);
        QPB_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, B.device));
        B.grid_cap = num_sms * per_sm;
        B.grid = (int)std::min<int64_t>((batch + kShNc - 1) / kShNc, (int64_t)B.grid_cap);
        return QPB200_OK;
    };
    rc = body();
    if (rc != QPB200_OK) {
        delete h;
        return rc;
    }
    B.setup_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    *out = h;
    return QPB200_OK;
}

// One call for a batch that is solved once (SolveQuadraticProgramBatch): the batch is cut into chunks, chunk c + 1 is
// staged and copied while the kernel of chunk c runs (two sets of device buffers on two streams; results come back
// through page-locked buffers so that no copy blocks the host), and nothing but two chunks is ever resident.
// cfg3 (5.5 GB of P and A for 179 ms of solving): upload and solve used to be serial.
int qpb200_batch_solve_once(int64_t batch, int64_t n, int64_t m, const double *P, const double *A, const double *q,
                            const double *l, const double *u, const qpb200_settings *settings, double *X_inout,
                            int32_t *flags, int64_t *iters, qpb200_info *info) {
    const auto t0 = std::chrono::steady_clock::now();
    if (!X_inout) return fail(QPB200_ERR_ARG, "qpb200_batch_solve_once: X_inout is NULL");
    qpb200_settings s;
    int rc = batch_check_args(batch, n, m, P, A, q, l, u, settings, s);
    if (rc) return rc;
    if (!all_finite(X_inout, (size_t)batch * n)) return fail(QPB200_ERR_NONFINITE, "qpb200_batch_solve_once: a start point holds NaN or Inf");
    const int64_t chunk = std::min<int64_t>(batch, s.reserved_i[QPB200_RSV_BATCH_CHUNK] > 0 ? s.reserved_i[QPB200_RSV_BATCH_CHUNK] : 4096);
    const int64_t nchunks = (batch + chunk - 1) / chunk;
    struct Lane {
        qpb200_batch *h = nullptr;
        double *pin = nullptr;           // page-locked: [q | l | u | X0] up, [X | iters | flags | totals | fail] down
        size_t cap = 0;
        bool pinned = false;
        int64_t first = -1, count = 0;   // the chunk whose results are pending in `pin`
        cudaEvent_t up = nullptr;        // the chunk's vectors have left `pin`
    } lane[2];
    const size_t up_doubles = (size_t)chunk * (2 * n + 2 * m);
    const size_t down_doubles = (size_t)chunk * (n + 2) + 8;             // X, iters (8 B), flags (4 B, padded to 8), totals, fail
    const size_t pin_bytes = (up_doubles + down_doubles) * sizeof(double);
    unsigned long long tot_it = 0, tot_rho = 0;
    int any_fail = 0;
    float dev_ms = 0.f;
    auto cleanup = [&]() {
        for (Lane &L : lane) {
            if (L.h) cudaStreamSynchronize(L.h->b.stream);
            if (L.up) cudaEventDestroy(L.up);
            if (L.pin) cached_pinned_free(L.pin, L.cap, L.pinned);
            delete L.h;
        }
    };
    // results of the chunk a lane holds: wait for its stream, copy out of the page-locked buffer
    auto collect = [&](Lane &L) -> int {
        if (L.first < 0) return QPB200_OK;
        DenseBatch &B = L.h->b;
        QPB_CUDA(cudaStreamSynchronize(B.stream));
        const double *down = L.pin + up_doubles;
        std::memcpy(X_inout + (size_t)L.first * n, down, (size_t)L.count * n * sizeof(double));
        const long long *it = reinterpret_cast<const long long *>(down + (size_t)chunk * n);
        const int *fl = reinterpret_cast<const int *>(down + (size_t)chunk * (n + 1));
        const unsigned long long *tt = reinterpret_cast<const unsigned long long *>(down + (size_t)chunk * (n + 2));
        if (iters) for (int64_t i = 0; i < L.count; ++i) iters[L.first + i] = it[i];
        if (flags) std::memcpy(flags + L.first, fl, (size_t)L.count * sizeof(int));
        tot_it += tt[0];
        tot_rho += tt[1];
        any_fail |= *reinterpret_cast<const int *>(tt + 4);
        float ms = 0.f;
        QPB_CUDA(cudaEventElapsedTime(&ms, B.ev0, B.ev1));
        dev_ms += ms;
        L.first = -1;
        return QPB200_OK;
    };
    for (Lane &L : lane) {
        L.h = new (std::nothrow) qpb200_batch();
        if (!L.h) { cleanup(); return fail(QPB200_ERR_ARG, "out of host memory"); }
        if ((rc = batch_setup(L.h->b, chunk, n, m, s))) { cleanup(); return rc; }
        L.pin = static_cast<double *>(cached_pinned_alloc(pin_bytes, &L.cap, &L.pinned));
        if (!L.pin || cudaEventCreateWithFlags(&L.up, cudaEventDisableTiming) != cudaSuccess) {
            cleanup();
            return fail(QPB200_ERR_CUDA, "qpb200_batch_solve_once: staging buffers unavailable");
        }
        if (nchunks == 1) break;
    }
#define QPB_ONCE(call)                                                                                       \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) {                                                                             \
            cleanup();                                                                                       \
            return fail(QPB200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                                    \
    } while (0)
    for (int64_t c = 0; c < nchunks; ++c) {
        Lane &L = lane[c & 1];
        DenseBatch &B = L.h->b;
        const int64_t first = c * chunk, count = std::min(chunk, batch - first);
        if ((rc = collect(L))) { cleanup(); return rc; }           // chunk c - 2: its buffers are free after this
        // vectors of the chunk -> page-locked buffer -> device (asynchronous: the host goes on to stage P and A)
        double *pq = L.pin, *pl = pq + (size_t)chunk * n, *pu = pl + (size_t)chunk * m, *px = pu + (size_t)chunk * m;
        std::memcpy(pq, q + (size_t)first * n, (size_t)count * n * sizeof(double));
        std::memcpy(pl, l + (size_t)first * m, (size_t)count * m * sizeof(double));
        std::memcpy(pu, u + (size_t)first * m, (size_t)count * m * sizeof(double));
        std::memcpy(px, X_inout + (size_t)first * n, (size_t)count * n * sizeof(double));
        QPB_ONCE(cudaMemcpyAsync(const_cast<double *>(B.prm.q), pq, (size_t)count * n * sizeof(double), cudaMemcpyHostToDevice, B.stream));
        QPB_ONCE(cudaMemcpyAsync(const_cast<double *>(B.prm.l), pl, (size_t)count * m * sizeof(double), cudaMemcpyHostToDevice, B.stream));
        QPB_ONCE(cudaMemcpyAsync(const_cast<double *>(B.prm.u), pu, (size_t)count * m * sizeof(double), cudaMemcpyHostToDevice, B.stream));
        QPB_ONCE(cudaMemcpyAsync(B.prm.X, px, (size_t)count * n * sizeof(double), cudaMemcpyHostToDevice, B.stream));
        if ((rc = batch_upload_matrices(B, count, P + (size_t)first * n * n, A + (size_t)first * m * n))) { cleanup(); return rc; }
        if ((rc = batch_launch(B, count))) { cleanup(); return rc; }
        double *down = L.pin + up_doubles;
        QPB_ONCE(cudaMemcpyAsync(down, B.prm.X, (size_t)count * n * sizeof(double), cudaMemcpyDeviceToHost, B.stream));
        QPB_ONCE(cudaMemcpyAsync(down + (size_t)chunk * n, B.prm.iters, (size_t)count * sizeof(long long), cudaMemcpyDeviceToHost, B.stream));
        QPB_ONCE(cudaMemcpyAsync(down + (size_t)chunk * (n + 1), B.prm.flags, (size_t)count * sizeof(int), cudaMemcpyDeviceToHost, B.stream));
        QPB_ONCE(cudaMemcpyAsync(down + (size_t)chunk * (n + 2), B.prm.totals, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, B.stream));
        QPB_ONCE(cudaMemcpyAsync(down + (size_t)chunk * (n + 2) + 4, B.prm.factor_fail, sizeof(int), cudaMemcpyDeviceToHost, B.stream));
        L.first = first;
        L.count = count;
    }
#undef QPB_ONCE
    for (Lane &L : lane)
        if (L.h && (rc = collect(L))) { cleanup(); return rc; }
    cleanup();
    if (info) {
        std::memset(info, 0, sizeof(*info));
        info->iterations = (int64_t)tot_it;
        info->rho_updates = (int64_t)tot_rho;
        info->rho_final = s.rho;
        info->res_prim = NAN;
        info->res_dual = NAN;
        info->solve_ms = dev_ms;                 // sum of the chunks' kernel times (they overlap the uploads)
        info->setup_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();   // whole call, wall
        info->kernel_launches = nchunks;
    }
    if (any_fail) return fail(QPB200_ERR_FACTOR, "Cholesky breakdown: a pivot of P + sigma I + rho A'A was not positive");
    return QPB200_OK;
}

int qpb200_batch_solve(qpb200_batch *h, double *X_inout, int32_t *flags, int64_t *iters, qpb200_info *info) {
    if (!h || !X_inout) return fail(QPB200_ERR_ARG, "qpb200_batch_solve: NULL argument");
    DenseBatch &B = h->b;
    QPB_CUDA(cudaSetDevice(B.device));
    const size_t nx = (size_t)B.batch * B.n;
    QPB_CUDA(cudaMemcpyAsync(B.prm.X, X_inout, nx * sizeof(double), cudaMemcpyHostToDevice, B.stream));
    if (int rc = batch_launch(B, B.batch)) return rc;
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

int qpb200_batch_update_vectors(qpb200_batch *h, const double *q, const double *l, const double *u) {
    if (!h) return fail(QPB200_ERR_ARG, "qpb200_batch_update_vectors: handle is NULL");
    DenseBatch &B = h->b;
    QPB_CUDA(cudaSetDevice(B.device));
    const size_t nq = (size_t)B.batch * B.n, nc = (size_t)B.batch * B.m;
    if (q && !all_finite(q, nq)) return fail(QPB200_ERR_NONFINITE, "q has a non-finite entry");
    for (int which = 0; which < 2; ++which) {
        const double *b = which ? u : l;
        if (!b) continue;
        for (size_t i = 0; i < nc; ++i)
            if (std::isnan(b[i])) return fail(QPB200_ERR_NONFINITE, "bound %zu is NaN", i);
    }
    if (l && u)
        for (size_t i = 0; i < nc; ++i)
            if (l[i] > u[i]) return fail(QPB200_ERR_NONFINITE, "bounds: need l <= u (entry %zu)", i);
    if (q) QPB_CUDA(cudaMemcpyAsync(const_cast<double *>(B.prm.q), q, nq * sizeof(double), cudaMemcpyHostToDevice, B.stream));
    if (l) QPB_CUDA(cudaMemcpyAsync(const_cast<double *>(B.prm.l), l, nc * sizeof(double), cudaMemcpyHostToDevice, B.stream));
    if (u) QPB_CUDA(cudaMemcpyAsync(const_cast<double *>(B.prm.u), u, nc * sizeof(double), cudaMemcpyHostToDevice, B.stream));
    QPB_CUDA(cudaStreamSynchronize(B.stream));       // the host arrays are borrowed for the duration of the call only
    return QPB200_OK;
}

void qpb200_batch_destroy(qpb200_batch *h) { delete h; }

}  // extern "C"
